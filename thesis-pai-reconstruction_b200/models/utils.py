"""Mirror of /root/reference/models/utils.py: ``denormalize, to_int, init_weights,
get_parameter_count, ssim, psnr, rmse`` with the same call signatures.  The three metrics run on the
fused B200 SSIM/PSNR/MSE kernel (pai_b200.metrics) instead of torchmetrics 0.11.4 and stay
differentiable w.r.t. ``pred``."""
import torch
import torch.nn as nn

from pai_b200 import metrics as _metrics


def denormalize(x: torch.Tensor) -> torch.Tensor:
    """[-1, 1] -> [0, 1], clamped (reference: models/utils.py:11)."""
    return torch.clamp(x * 0.5 + 0.5, 0, 1)


def to_int(x: torch.Tensor) -> torch.Tensor:
    """float [0, 1] -> uint8, torchvision ``ConvertImageDtype(torch.uint8)`` semantics (models/utils.py:12)."""
    if x.dtype == torch.uint8:
        return x
    return x.mul(255.0 + 1.0 - 1e-3).to(torch.uint8)


_WEIGHTED = (nn.Conv1d, nn.Conv2d, nn.ConvTranspose2d, nn.Linear)
_NORMS = (nn.BatchNorm1d, nn.BatchNorm2d, nn.GroupNorm, nn.LayerNorm)


def init_weights(module: nn.Module):
    """N(0, 0.02) conv / linear weights, unit-gain zero-shift norms (models/utils.py:15-28)."""
    if isinstance(module, _WEIGHTED):
        nn.init.normal_(module.weight, 0.0, 0.02)
    if isinstance(module, _NORMS):
        nn.init.constant_(module.weight, 1.0)
        nn.init.constant_(module.bias, 0.0)


def get_parameter_count(model) -> int:
    return sum(p.numel() for p in model.parameters()) if isinstance(model, nn.Module) else 0


def ssim(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return _metrics.ssim(pred, target)


def psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return _metrics.psnr(pred, target)


def rmse(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return _metrics.rmse(pred, target)
