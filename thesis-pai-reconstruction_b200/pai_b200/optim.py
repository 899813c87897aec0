"""Adam on the B200 kernels (csrc/optim.cu), a drop-in subclass of ``torch.optim.Adam``.

The reference builds ``torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999), eps=1e-7)``
(models/wrapper.py:97-115) and steps it from ``training_step`` (:136, :160).  ``FusedAdam`` keeps that
constructor, ``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter) and arithmetic,
but runs the whole step in a handful of launches:

* every 4x4 convolution weight ``[A, B, 4, 4]`` is updated by ONE kernel that also rewrites the two bf16
  GEMM-operand packs the implicit-GEMM kernels read (``engine`` pack cache) -- no separate repack pass;
* all remaining tensors (biases, BatchNorm affine, thin layers) share one multi-tensor launch.

Options the reference does not use (``amsgrad``, ``weight_decay``, ``maximize``, tensor ``lr``) and
non-CUDA parameters take ``torch.optim.Adam.step`` unchanged (optimizer plumbing, not the hot path).
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import engine, lib, ops


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class FusedAdam(torch.optim.Adam):
    def _plain(self, group) -> bool:
        return not (group.get("amsgrad") or group.get("weight_decay") or group.get("maximize")
                    or group.get("capturable") or group.get("differentiable")
                    or isinstance(group["lr"], torch.Tensor))

    @torch.no_grad()
    def step(self, closure=None):
        groups = self.param_groups
        ok = all(self._plain(g) for g in groups) and all(
            p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()
            for g in groups for p in g["params"])
        if not ok:
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stream = ops._stream()
        for group in groups:
            beta1, beta2 = group["betas"]
            lr, eps = float(group["lr"]), float(group["eps"])
            by_step = {}
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not g.is_contiguous():
                    g = g.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = int(st["step"].item()) if st["step"].device.type == "cpu" else int(st["step"])
                by_step.setdefault(t, []).append((p, g, st))
            for t, items in by_step.items():
                step_size = lr / (1.0 - beta1 ** t)
                inv_bc2 = 1.0 / math.sqrt(1.0 - beta2 ** t)
                small = []
                for p, g, st in items:
                    packs = engine.fused_pack_targets(p)
                    if packs is None:
                        p.__dict__.pop("_pai_packs", None)      # thin-layer packs are rebuilt lazily
                        small.append((p, g, st))
                        continue
                    p1, p2, b_pad = packs
                    lib.call("pai_adam_pack_conv4x4", _ptr(p), _ptr(g), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]),
                             p.shape[0], p.shape[1], beta1, beta2, step_size, inv_bc2, eps, _ptr(p1), _ptr(p2),
                             b_pad, stream)
                    engine.restamp_packs(p)
                if small:
                    n = len(small)
                    arr = ctypes.c_void_p * n
                    lib.call("pai_adam_multi", n,
                             arr(*[p.data_ptr() for p, _, _ in small]), arr(*[g.data_ptr() for _, g, _ in small]),
                             arr(*[s["exp_avg"].data_ptr() for _, _, s in small]),
                             arr(*[s["exp_avg_sq"].data_ptr() for _, _, s in small]),
                             (ctypes.c_int * n)(*[p.numel() for p, _, _ in small]),
                             beta1, beta2, step_size, inv_bc2, eps, stream, kernels=(n + 47) // 48)
        return loss
