"""The evaluation loop of the reference's ``report.py`` (lines 63-146, 188-217) on the GPU, in one place.

``report.py`` runs the frozen model over the prediction set, brings predictions and targets to the host, computes
per-image SSIM (+ full maps) / PSNR / MSE in chunks of 64 with torchmetrics, the SSIM over 16 depth bands, the global
RMSE, and writes PNG / CSV files.  Here everything up to the file writing stays on the device: batched inference through
the drop-in model, ``denormalize``, ONE pass of the metric kernel per pair (per-image SSIM, depth bands, squared error,
optional full maps), the reductions of ``metrics.finalize_report`` and the uint8 conversion of the maps (``to_int``).
PNG / CSV encoding is left to the caller (``depth_csv`` reproduces ``report.py:102-105``'s text).

With ``distributed=True`` every rank evaluates its contiguous shard of the pairs (``metrics.shard_bounds``) and the
returned statistics are those of the WHOLE set; predictions and maps stay local to the rank that produced them.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch

from . import metrics, ops


def _denormalize(x: torch.Tensor) -> torch.Tensor:          # models/utils.py:11
    return torch.clamp(x * 0.5 + 0.5, 0, 1)


def _to_int(x: torch.Tensor) -> torch.Tensor:
    """models/utils.py:12 (torchvision ConvertImageDtype(torch.uint8): ``x * 255.999`` truncated) on the device.  The
    reference converts the RAW SSIM map (report.py:134-135), whose values may be negative: like the host conversion
    it runs, out-of-range values wrap modulo 256 instead of being clamped."""
    if not x.is_cuda:
        raise RuntimeError("pai_b200.report: the uint8 conversion runs on the GPU (pai_to_uint8); there is no CPU fallback")
    return ops.to_uint8(x)


def hot_images(preds: torch.Tensor) -> torch.Tensor:
    """``output_hot_image`` (report.py:220-233) without the PNG encoder: matplotlib's "afmhot" colormap of every
    denormalised prediction ``[n, 1, h, w]`` -> uint8 ``[n, 3, h, w]``, one kernel on the device."""
    return ops.afmhot_uint8(preds)


def depth_csv(depth_ssim: torch.Tensor) -> str:
    """``depth,mean,std`` text of report.py:102-105 from the ``[16, 2]`` depth-SSIM table."""
    text = "depth,mean,std\n"
    for depth, (mean, std) in enumerate(depth_ssim.tolist(), 1):
        text += f"{depth},{mean},{std}\n"
    return text


@torch.no_grad()
def evaluate(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], want_maps: bool = True,
             distributed: bool = False, device=None):
    """``batches`` yields ``(input, target)`` pairs in [-1, 1] (host or device tensors), like the prediction dataloader
    of ``report.py:58-70``.  ``model`` is a frozen drop-in module (or any callable; ``report.py``'s "identity").

    Returns a dict: ``preds`` / ``targets`` (denormalised, on the device), ``ssim`` / ``psnr`` / ``mse`` per image,
    ``ssim_stat`` / ``psnr_stat`` / ``rmse_stat`` (report.py:143-146), ``depth_ssim`` ``[16, 2]``, ``ssim_maps`` and
    ``ssim_maps_uint8`` (when ``want_maps``), ``parameter_count``."""
    if device is None:
        device = next(model.parameters()).device if isinstance(model, torch.nn.Module) else torch.device("cuda")
    if isinstance(model, torch.nn.Module):
        model.eval()
    preds, targets = [], []
    for x, t in batches:
        y = model(x.to(device, non_blocking=True))
        preds.append(_denormalize(y.float()))
        targets.append(_denormalize(t.to(device, non_blocking=True).float()))
    preds, targets = torch.cat(preds), torch.cat(targets)
    res = metrics.report_metrics(preds, targets, want_maps=want_maps, want_depth=True, distributed=distributed)
    out = {
        "preds": preds, "targets": targets,
        "ssim": res["ssim"], "psnr": res["psnr"], "mse": res["mse"],
        "ssim_stat": res["ssim_mean"], "psnr_stat": res["psnr_mean"], "rmse_stat": res["rmse"],
        "depth_ssim": res["depth_ssim"],
        "ssim_maps": res["ssim_maps"],
        "ssim_maps_uint8": _to_int(res["ssim_maps"]) if res["ssim_maps"] is not None else None,
        "hot_images_uint8": hot_images(preds) if want_maps else None,
        "parameter_count": sum(p.numel() for p in model.parameters()) if isinstance(model, torch.nn.Module) else 0,
    }
    return out
