"""ORACLE (test infrastructure, not product code).

CPU restatement of the three ``torchmetrics==0.11.4`` functionals the reference
binds in ``models/utils.py:38-47`` and ``report.py:3-7,78-96,146,207-212``:

* ``structural_similarity_index_measure``  (upstream ``functional/image/ssim.py``)
* ``peak_signal_noise_ratio``              (upstream ``functional/image/psnr.py``)
* ``mean_squared_error``                   (upstream ``functional/regression/mse.py``)

``torchmetrics`` is a third-party dependency pinned in the reference's
``requirements.txt:6`` and is NOT vendored under /root/reference nor installed in
this image, so this file restates its published 0.11.4 algorithm (SURVEY.md
Appendix A).  PARITY UNPINNED against torchmetrics itself: the reference holds
no golden vectors for these calls; the restatement is pinned instead by
closed-form known answers and an independent fp64 separable implementation
(``oracle/ssim_ref.c``), see ``tests/test_oracle_metrics.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- helpers
def _gaussian_1d(size: int, sigma: float, dtype, device) -> Tensor:
    """upstream ``functional/image/helper.py::_gaussian`` -- taps at integer offsets
    centred on zero, normalised to unit sum, shape [1, size]."""
    half = (size - 1) / 2.0
    pos = torch.arange(-half, half + 1.0, 1.0, dtype=dtype, device=device)
    g = torch.exp(-0.5 * (pos / sigma) ** 2)
    return (g / g.sum()).unsqueeze(0)


def _gaussian_window(channels: int, size: Sequence[int], sigma: Sequence[float], dtype, device) -> Tensor:
    """upstream ``_gaussian_kernel_2d`` -- outer product, one copy per channel."""
    gx = _gaussian_1d(size[0], sigma[0], dtype, device)
    gy = _gaussian_1d(size[1], sigma[1], dtype, device)
    win = gx.t() @ gy
    return win.expand(channels, 1, size[0], size[1])


def _reduce(x: Tensor, reduction: Optional[str]) -> Tensor:
    if reduction == "elementwise_mean":
        return x.mean()
    if reduction == "sum":
        return x.sum()
    if reduction is None or reduction == "none":
        return x
    raise ValueError("Expected reduction to be one of 'elementwise_mean', 'sum', 'none' or None")


def _check_same_shape(preds: Tensor, target: Tensor) -> None:
    if preds.shape != target.shape:
        raise RuntimeError(
            f"Predictions and targets are expected to have the same shape, but got {preds.shape} and {target.shape}."
        )


# --------------------------------------------------------------------------- SSIM
def structural_similarity_index_measure(
    preds: Tensor,
    target: Tensor,
    gaussian_kernel: bool = True,
    sigma: Union[float, Sequence[float]] = 1.5,
    kernel_size: Union[int, Sequence[int]] = 11,
    reduction: Optional[str] = "elementwise_mean",
    data_range: Optional[float] = None,
    k1: float = 0.01,
    k2: float = 0.03,
    return_full_image: bool = False,
    return_contrast_sensitivity: bool = False,
):
    """SURVEY.md Appendix A steps 1-9.  Only the 2-D (4-D tensor) Gaussian path the
    reference uses is restated; the uniform-kernel and 3-D branches are not."""
    if preds.dtype != target.dtype:
        target = target.to(preds.dtype)
    _check_same_shape(preds, target)
    if preds.ndim != 4:
        raise ValueError(
            "Expected `preds` and `target` to have BxCxHxW shape."
            f" Got preds: {preds.shape} and target: {target.shape}."
        )
    if not gaussian_kernel:
        raise NotImplementedError("oracle restates the Gaussian-window path only")
    if return_contrast_sensitivity:
        raise NotImplementedError("oracle does not restate contrast sensitivity")

    if not isinstance(sigma, Sequence):
        sigma = 2 * [sigma]
    if not isinstance(kernel_size, Sequence):
        kernel_size = 2 * [kernel_size]
    if any(k % 2 == 0 or k <= 0 for k in kernel_size):
        raise ValueError(f"Expected `kernel_size` to have odd positive number. Got {kernel_size}.")
    if any(s <= 0 for s in sigma):
        raise ValueError(f"Expected `sigma` to have positive number. Got {sigma}.")

    if data_range is None:
        data_range = max(preds.max() - preds.min(), target.max() - target.min())

    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    dtype, device = preds.dtype, preds.device
    channels = preds.size(1)

    # 0.11.x: the Gaussian support comes from sigma, not from kernel_size.
    win_size = [int(3.5 * s + 0.5) * 2 + 1 for s in sigma]
    pad_h = (win_size[0] - 1) // 2
    pad_w = (win_size[1] - 1) // 2

    p = F.pad(preds, (pad_w, pad_w, pad_h, pad_h), mode="reflect")
    t = F.pad(target, (pad_w, pad_w, pad_h, pad_h), mode="reflect")
    window = _gaussian_window(channels, win_size, sigma, dtype, device)

    stacked = torch.cat((p, t, p * p, t * t, p * t))  # [5B, C, H+2p, W+2p]
    filt = F.conv2d(stacked, window, groups=channels)
    mu_p, mu_t, e_pp, e_tt, e_pt = filt.split(preds.shape[0])

    mu_pp = mu_p.pow(2)
    mu_tt = mu_t.pow(2)
    mu_pt = mu_p * mu_t
    var_p = e_pp - mu_pp  # no clamping in 0.11.4
    var_t = e_tt - mu_tt
    cov = e_pt - mu_pt

    upper = 2 * cov + c2
    lower = var_p + var_t + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)

    inner = full[..., pad_h:-pad_h, pad_w:-pad_w]
    per_image = inner.reshape(inner.shape[0], -1).mean(-1)
    out = _reduce(per_image, reduction)
    if return_full_image:
        return out, full
    return out


# --------------------------------------------------------------------------- PSNR
def peak_signal_noise_ratio(
    preds: Tensor,
    target: Tensor,
    data_range: Optional[float] = None,
    base: float = 10.0,
    reduction: Optional[str] = "elementwise_mean",
    dim=None,
) -> Tensor:
    """sse over the whole tensor, n = numel."""
    if dim is not None:
        raise NotImplementedError("oracle restates dim=None only (what the reference calls)")
    if data_range is None:
        data_range = target.max() - target.min()
    else:
        data_range = torch.tensor(float(data_range))
    _check_same_shape(preds, target)
    diff = preds - target
    sse = torch.sum(diff * diff)
    n = torch.tensor(target.numel(), device=target.device)
    base_e = 2 * torch.log(data_range) - torch.log(sse / n)
    val = base_e * (10 / torch.log(torch.tensor(base)))
    return _reduce(val, reduction)


# --------------------------------------------------------------------------- MSE
def mean_squared_error(preds: Tensor, target: Tensor, squared: bool = True) -> Tensor:
    _check_same_shape(preds, target)
    diff = preds - target
    sse = torch.sum(diff * diff)
    mse = sse / target.numel()
    return mse if squared else torch.sqrt(mse)
