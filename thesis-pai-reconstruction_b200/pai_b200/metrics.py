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
    with lib.on_device(pred):
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
        with lib.on_device(pred):
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
def per_image_stats(preds: torch.Tensor, targets: torch.Tensor, chunk: int = 4096, want_maps: bool = False,
                    want_depth: bool = True):
    """One pass of the metric kernel per pair -> per-image sufficient statistics on the GPU:
    ``ssim [N]``, ``sse [N]`` (squared error over C*H*W), ``bands [N, 16]`` (depth-band SSIMs, or None), ``maps``."""
    _check(preds, targets)
    n, c, h, w = preds.shape
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
            maps.append(fm)
    dev = ssims[0].device if ssims else torch.device("cuda")
    cat = lambda xs, shape: torch.cat(xs) if xs else torch.zeros(shape, device=dev)      # noqa: E731
    return (cat(ssims, (0,)), cat(sses, (0,)), cat(bands, (0, 16)) if want_depth else None,
            torch.cat(maps) if (want_maps and maps) else None)


def finalize_report(ssim_i: torch.Tensor, sse_i: torch.Tensor, bands, per: int):
    """The reductions report.py applies to the per-image values (report.py:88-96,144-146,213-214): per-image PSNR /
    MSE, their means, the global RMSE from the summed squared error, depth-band mean and unbiased std over images.
    ``per`` = C*H*W elements per image.  Plain tensor arithmetic on [N]-vectors (runs on any device)."""
    n = ssim_i.shape[0]
    mse_i = sse_i / per
    psnr_i = _psnr_from_sse(sse_i, per)
    res = {
        "ssim": ssim_i, "psnr": psnr_i, "mse": mse_i,
        "ssim_mean": ssim_i.mean(), "psnr_mean": psnr_i.mean(),
        "rmse": torch.sqrt(sse_i.double().sum() / (n * per)).float(),
        "depth_ssim": None,
    }
    if bands is not None:
        res["depth_ssim"] = torch.stack([bands.mean(0), bands.std(0)], dim=1)          # [16, 2]
    return res


def gather_stats(ssim_i, sse_i, bands):
    """Data-parallel evaluation sweep: every rank evaluated a contiguous shard of the pairs; the
    per-image vectors (19 floats per image) are all-gathered in rank order so every rank can apply
    ``finalize_report`` to the full set exactly as the single-process sweep does.  No collective touches the images.
    Shards may have different lengths.  Without an initialised process group this is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ssim_i, sse_i, bands
    world = dist.get_world_size()
    cols = [ssim_i.reshape(-1, 1).float(), sse_i.reshape(-1, 1).float()]
    if bands is not None:
        cols.append(bands.float())
    local = torch.cat(cols, dim=1).contiguous()                       # [n_local, 2 (+16)]
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    width, most = local.shape[1], max(counts)
    padded = torch.zeros(most, width, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    full = torch.cat([p[:k] for p, k in zip(parts, counts)])
    return full[:, 0].contiguous(), full[:, 1].contiguous(), (full[:, 2:].contiguous() if bands is not None else None)


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of ``n`` pairs for ``rank`` (the first ``n % world`` ranks get one extra pair)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@torch.no_grad()
def report_metrics(preds: torch.Tensor, targets: torch.Tensor, chunk: int = 4096, want_maps: bool = False,
                   want_depth: bool = True, distributed: bool = False):
    """report.py:72-104,144-146,188-217 in one pass per pair.

    Returns a dict with per-image ``ssim``/``psnr``/``mse`` ``[N]``, ``depth_ssim`` ``[16, 2]`` (mean and
    unbiased std over images per depth band), ``ssim_mean``, ``psnr_mean``, global ``rmse`` and optionally
    the full SSIM maps.  ``preds``/``targets`` may live on the host; they are streamed to the GPU in
    ``chunk``-sized slices.

    ``distributed=True``: ``preds``/``targets`` are this rank's shard (``shard_bounds``); the per-image vectors are
    all-gathered and every rank returns the metrics of the WHOLE sweep (``ssim_maps`` stays local)."""
    n, c, h, w = preds.shape
    out_dev = preds.device
    ssim_i, sse_i, bands, maps = per_image_stats(preds, targets, chunk, want_maps, want_depth)
    if distributed:
        ssim_i, sse_i, bands = gather_stats(ssim_i, sse_i, bands)
    res = finalize_report(ssim_i, sse_i, bands, c * h * w)
    res = {k: (v.to(out_dev) if isinstance(v, torch.Tensor) else v) for k, v in res.items()}
    res["ssim_maps"] = maps.to(out_dev) if maps is not None else None
    return res
