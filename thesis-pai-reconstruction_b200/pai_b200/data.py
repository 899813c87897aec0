"""The reference's per-image input transform (dataset.py:51-61,126-134) as one device kernel over a batch.

``ImageDataset.__getitem__`` reads two grayscale PNGs and applies ``Resize((256, 256), antialias=True)`` ->
``ConvertImageDtype(float32)`` -> ``Normalize`` to each on the host, one image at a time; at the 7-8 k images/s a B200 trains
this model the host transform is the bottleneck.  ``preprocess`` takes the DECODED uint8 images (file
decoding stays on the host: ``torchvision.io.read_image``), stacked per size, and produces the normalised fp32 batch on the
GPU.  The reference's ``Normalize((0.5,)*3, (0.5,)*3)`` raises on its own ``ImageReadMode.GRAY`` tensors;
the single channel is normalised with the same constants here.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from . import lib


def preprocess(images_u8: torch.Tensor, size: Tuple[int, int] = (256, 256), normalize: bool = True,
               round_like_torchvision: bool = True) -> torch.Tensor:
    """``[n, 1, H, W]`` (or ``[n, H, W]``) uint8 -> fp32 ``[n, 1, size[0], size[1]]`` in [-1, 1] (``normalize``) or [0, 1].
    Host tensors are copied to the current CUDA device first (pinned memory makes that copy asynchronous)."""
    if images_u8.dtype != torch.uint8:
        raise TypeError("pai_b200.data.preprocess expects decoded uint8 images")
    if images_u8.dim() == 4:
        if images_u8.shape[1] != 1:
            raise RuntimeError("pai_b200.data.preprocess handles grayscale images (ImageReadMode.GRAY, dataset.py:126-131)")
        images_u8 = images_u8[:, 0]
    if not images_u8.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("pai_b200.data needs a CUDA device; there is no CPU fallback")
        images_u8 = images_u8.to("cuda", non_blocking=True)
    images_u8 = images_u8.contiguous()
    n, ih, iw = images_u8.shape
    out = torch.empty(n, 1, size[0], size[1], dtype=torch.float32, device=images_u8.device)
    with lib.on_device(images_u8):
        lib.call("pai_resize_aa_normalize_u8", ctypes.c_void_p(images_u8.data_ptr()), n, ih, iw, size[0], size[1],
                 int(normalize), int(round_like_torchvision), ctypes.c_void_p(out.data_ptr()),
                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def preprocess_pairs(inputs_u8: Sequence[torch.Tensor], targets_u8: Sequence[torch.Tensor], size=(256, 256), normalize=True):
    """A list of decoded ``(input, ground truth)`` images of possibly different sizes -> the ``(x, target)`` batch
    ``training_step`` takes (dataset.py:126-134 + the DataLoader's collate)."""
    def run(items):
        outs = [preprocess(t.reshape(1, *t.shape[-2:]), size, normalize) for t in items]
        return torch.cat(outs) if outs else torch.empty(0, 1, *size)
    return run(inputs_u8), run(targets_u8)
