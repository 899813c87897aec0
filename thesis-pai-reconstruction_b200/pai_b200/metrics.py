"""SSIM / PSNR / RMSE on the B200 metric kernel (csrc/ssim.cu) with autograd.

Mirrors what the reference obtains from ``torchmetrics==0.11.4`` through ``models/utils.py:38-47``
(``data_range=1.0``) and ``report.py:78-96,146,188-217``:

* ``ssim(pred, target)``  -> 0-d tensor, batch mean (``reduction="elementwise_mean"``)
* ``psnr(pred, target)``  -> 0-d tensor from the batch-global MSE
* ``rmse(pred, target)``  -> 0-d tensor
* ``report_metrics(preds, targets)`` -> everything report.py's metric loop + ``depth_ssim`` produce,
  from ONE pass over each image pair.

All of them are differentiable w.r.t. ``pred`` (fused backward kernel).  Host (CPU) tensors are staged
to the current CUDA device and the result is returned on the caller's device; without a CUDA device a
RuntimeError is raised -- there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import lib

_F32, _BF16 = 0, 1


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_device(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("pai_b200.metrics needs a CUDA device (B200); there is no CPU fallback")
    return t.to("cuda", non_blocking=True)


def _canon(t: torch.Tensor) -> torch.Tensor:
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t.contiguous()


def _check(pred, target):
    if pred.shape != target.shape:
        raise RuntimeError(
            f"Predictions and targets are expected to have the same shape, but got {pred.shape} and {target.shape}.")
    if pred.dim() != 4:
        raise ValueError(f"Expected `preds` and `target` to have BxCxHxW shape. Got preds: {pred.shape} and target: {target.shape}.")


def _launch_fwd(pred, target, denormalize, want_bands, want_map):
    b, c, h, w = pred.shape
    n = b * c
    dev = pred.device
    ssim_sum = torch.empty(n, dtype=torch.float32, device=dev)
    sse = torch.empty(n, dtype=torch.float32, device=dev)
    bands = torch.empty(n, 16, dtype=torch.float32, device=dev) if want_bands else None
    fmap = torch.empty(b, c, h, w, dtype=torch.float32, device=dev) if want_map else None
    lib.call("pai_ssim_psnr_fwd", _ptr(pred), _ptr(target), _BF16 if pred.dtype == torch.bfloat16 else _F32, n, h, w,
             int(denormalize), _ptr(ssim_sum), _ptr(bands), _ptr(sse), _ptr(fmap), _stream())
    return ssim_sum, sse, bands, fmap


class _SsimSse(torch.autograd.Function):
    """(pred, target) -> (ssim_sum[B*C], sse[B*C]); backward = pai_ssim_psnr_bwd."""

    @staticmethod
    def forward(ctx, pred, target, denormalize):
        pred_c, target_c = _canon(pred), _canon(target).to(_canon(pred).dtype)
        ssim_sum, sse, _, _ = _launch_fwd(pred_c, target_c, denormalize, False, False)
        ctx.save_for_backward(pred_c, target_c)
        ctx.denormalize = denormalize
        ctx.in_dtype = pred.dtype
        return ssim_sum, sse

    @staticmethod
    def backward(ctx, g_ssim, g_sse):
        pred, target = ctx.saved_tensors
        b, c, h, w = pred.shape
        n = b * c
        g_ssim = torch.zeros(n, device=pred.device) if g_ssim is None else g_ssim.float().contiguous()
        g_sse = None if g_sse is None else g_sse.float().contiguous()
        work = torch.empty(n * h * w * 4, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        lib.call("pai_ssim_psnr_bwd", _ptr(pred), _ptr(target), _BF16 if pred.dtype == torch.bfloat16 else _F32, n, h,
                 w, int(ctx.denormalize), _ptr(g_ssim), _ptr(g_sse), _ptr(work), _ptr(grad), _stream(), kernels=2)
        return grad.to(ctx.in_dtype), None, None


def ssim_sse(pred: torch.Tensor, target: torch.Tensor, denormalize: bool = False):
    """Differentiable per-plane ``(ssim_sum[B*C], sse[B*C])`` -- the two sufficient statistics."""
    _check(pred, target)
    return _SsimSse.apply(_to_device(pred), _to_device(target), bool(denormalize))


def _psnr_from_sse(sse_total: torch.Tensor, numel: int) -> torch.Tensor:
    # torchmetrics 0.11.4 _psnr_compute with data_range=1.0, base=10
    return (2 * math.log(1.0) - torch.log(sse_total / numel)) * (10.0 / math.log(10.0))


def ssim(pred: torch.Tensor, target: torch.Tensor, denormalize: bool = False) -> torch.Tensor:
    """models/utils.py:38-39."""
    out_dev = pred.device
    s, _ = ssim_sse(pred, target, denormalize)
    b, c, h, w = pred.shape
    return (s.sum() / (b * c * (h - 10) * (w - 10))).to(out_dev)


def psnr(pred: torch.Tensor, target: torch.Tensor, denormalize: bool = False) -> torch.Tensor:
    """models/utils.py:42-43 (MSE over the whole batch tensor)."""
    out_dev = pred.device
    _, e = ssim_sse(pred, target, denormalize)
    return _psnr_from_sse(e.sum(), pred.numel()).to(out_dev)


def rmse(pred: torch.Tensor, target: torch.Tensor, denormalize: bool = False) -> torch.Tensor:
    """models/utils.py:46-47."""
    out_dev = pred.device
    _, e = ssim_sse(pred, target, denormalize)
    return torch.sqrt(e.sum() / pred.numel()).to(out_dev)


def train_metrics(pred: torch.Tensor, target: torch.Tensor, denormalize: bool = True):
    """One kernel launch -> ``(ssim, psnr, rmse)`` as logged every step by models/wrapper.py:150-156;
    all three carry gradients w.r.t. ``pred``."""
    s, e = ssim_sse(pred, target, denormalize)
    b, c, h, w = pred.shape
    ssim_v = s.sum() / (b * c * (h - 10) * (w - 10))
    et = e.sum()
    return ssim_v, _psnr_from_sse(et, pred.numel()), torch.sqrt(et / pred.numel())


@torch.no_grad()
def report_metrics(preds: torch.Tensor, targets: torch.Tensor, chunk: int = 4096, want_maps: bool = False,
                   want_depth: bool = True):
    """report.py:72-104,144-146,188-217 in one pass per pair.

    Returns a dict with per-image ``ssim``/``psnr``/``mse`` ``[N]``, ``depth_ssim`` ``[16, 2]`` (mean and
    unbiased std over images per depth band), ``ssim_mean``, ``psnr_mean``, global ``rmse`` and optionally
    the full SSIM maps.  ``preds``/``targets`` may live on the host; they are streamed to the GPU in
    ``chunk``-sized slices."""
    _check(preds, targets)
    n, c, h, w = preds.shape
    out_dev = preds.device
    ssims, sses, bands, maps = [], [], [], []
    for p, t in zip(preds.split(chunk), targets.split(chunk)):
        p, t = _canon(_to_device(p)), _canon(_to_device(t))
        t = t.to(p.dtype)
        s, e, bd, fm = _launch_fwd(p, t, False, want_depth, want_maps)
        m = p.shape[0]
        ssims.append(s.view(m, c).sum(1) / (c * (h - 10) * (w - 10)))
        sses.append(e.view(m, c).sum(1))
        if want_depth:
            bands.append(bd.view(m, c, 16).sum(1) / (c * (h // 16 - 10) * (w - 10)))
        if want_maps:
            maps.append(fm.to(out_dev))
    ssim_i, sse_i = torch.cat(ssims), torch.cat(sses)
    per = c * h * w
    mse_i = sse_i / per
    psnr_i = _psnr_from_sse(sse_i, per)
    res = {
        "ssim": ssim_i.to(out_dev),
        "psnr": psnr_i.to(out_dev),
        "mse": mse_i.to(out_dev),
        "ssim_mean": ssim_i.mean().to(out_dev),
        "psnr_mean": psnr_i.mean().to(out_dev),
        "rmse": torch.sqrt(sse_i.double().sum() / (n * per)).float().to(out_dev),
        "ssim_maps": torch.cat(maps) if want_maps else None,
        "depth_ssim": None,
    }
    if want_depth:
        bd = torch.cat(bands)                       # [N, 16]
        res["depth_ssim"] = torch.stack([bd.mean(0), bd.std(0)], dim=1).to(out_dev)
    return res
