"""Adam on the B200 kernels (csrc/optim.cu), a drop-in subclass of ``torch.optim.Adam``.

The reference builds ``torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999), eps=1e-7)``
(models/wrapper.py:97-115) and steps it from ``training_step`` (:136, :160).  ``FusedAdam`` keeps that
constructor, ``state_dict`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter) and arithmetic,
but runs the whole step in a handful of launches:

* all 4x4 convolution weights ``[A, B, 4, 4]`` are updated by ONE launch that also rewrites the two bf16
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
    """``torch.optim.Adam`` stepped by csrc/optim.cu.

    When every parameter of a group receives a gradient at every step (always true for the reference's two
    optimizers) the step count ``t`` of the group lives ON THE DEVICE (``pai_adam_prepare`` increments it and derives
    the two bias-correction scalars), so ``step()`` issues no host-dependent kernel argument and the whole training
    step can be captured in a CUDA graph (pai_b200.graph).  ``state_dict()`` / ``load_state_dict()`` translate between
    that counter and torch's per-parameter ``state[p]["step"]`` tensors, so checkpoints are interchangeable."""

    def _plain(self, group) -> bool:
        return not (group.get("amsgrad") or group.get("weight_decay") or group.get("maximize")
                    or group.get("capturable") or group.get("differentiable")
                    or isinstance(group["lr"], torch.Tensor))

    # ---- device-side step counter ---------------------------------------------------------------------------
    def _dev_state(self, gi: int, device):
        store = self.__dict__.setdefault("_pai_dev", {})
        ent = store.get(gi)
        if ent is None or ent["step"].device != device:
            ent = {"step": torch.zeros(1, dtype=torch.int32, device=device),
                   "dyn": torch.zeros(2, dtype=torch.float32, device=device), "uniform": None}
            store[gi] = ent
        return ent

    def _sync_host_steps(self):
        """Writes the device counters into torch's per-parameter ``step`` tensors."""
        for gi, ent in self.__dict__.get("_pai_dev", {}).items():
            if not ent["uniform"]:
                continue
            t = float(int(ent["step"].item()))
            for p in self.param_groups[gi]["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = torch.tensor(t, dtype=torch.float32)

    def state_dict(self):
        self._sync_host_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.__dict__.pop("_pai_dev", None)          # re-derived from the loaded per-parameter steps at the next step

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _launch(self, items, beta1, beta2, step_size, inv_bc2, eps, dyn, stream):
        small, fused = [], []
        for p, g, st in items:
            p.__dict__.pop("_pai_aux", None)            # packs the kernels below do not rewrite
            packs = engine.fused_pack_targets(p)
            if packs is None:
                p.__dict__.pop("_pai_packs", None)      # thin-layer packs are rebuilt lazily
                small.append((p, g, st))
                continue
            fused.append((p, g, st, packs))
        if fused:
            n = len(fused)
            arr, ints = ctypes.c_void_p * n, ctypes.c_int * n
            lib.call("pai_adam_pack_conv4x4_multi", n,
                     arr(*[p.data_ptr() for p, _, _, _ in fused]), arr(*[g.data_ptr() for _, g, _, _ in fused]),
                     arr(*[s["exp_avg"].data_ptr() for _, _, s, _ in fused]),
                     arr(*[s["exp_avg_sq"].data_ptr() for _, _, s, _ in fused]),
                     ints(*[p.shape[0] for p, _, _, _ in fused]), ints(*[p.shape[1] for p, _, _, _ in fused]),
                     arr(*[0 if k[0] is None else k[0].data_ptr() for _, _, _, k in fused]),
                     arr(*[0 if k[1] is None else k[1].data_ptr() for _, _, _, k in fused]),
                     ints(*[k[2] for _, _, _, k in fused]),
                     beta1, beta2, step_size, inv_bc2, eps, _ptr(dyn), stream, kernels=(n + 23) // 24)
            for p, _, _, _ in fused:
                engine.restamp_packs(p)
        if small:
            n = len(small)
            arr = ctypes.c_void_p * n
            lib.call("pai_adam_multi", n,
                     arr(*[p.data_ptr() for p, _, _ in small]), arr(*[g.data_ptr() for _, g, _ in small]),
                     arr(*[s["exp_avg"].data_ptr() for _, _, s in small]),
                     arr(*[s["exp_avg_sq"].data_ptr() for _, _, s in small]),
                     (ctypes.c_int * n)(*[p.numel() for p, _, _ in small]),
                     beta1, beta2, step_size, inv_bc2, eps, _ptr(dyn), stream, kernels=(n + 47) // 48)

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
        for gi, group in enumerate(groups):
            beta1, beta2 = group["betas"]
            lr, eps = float(group["lr"]), float(group["eps"])
            params = group["params"]
            if not params:
                continue
            items = []
            for p in params:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not g.is_contiguous():
                    g = g.contiguous()
                items.append((p, g, self._init_state(p)))
            ent = self._dev_state(gi, params[0].device)
            if ent["uniform"] is None:
                # first step of this group (or after load_state_dict): adopt the device counter when all parameters
                # share one step count and all of them are being stepped
                steps = {float(st["step"]) for _, _, st in items}
                ent["uniform"] = len(items) == len(params) and len(steps) == 1
                if ent["uniform"]:
                    ent["step"].fill_(int(steps.pop()))
            if ent["uniform"] and len(items) != len(params):
                self._sync_host_steps()                  # a parameter skipped this step: per-parameter counts from now on
                ent["uniform"] = False
            if ent["uniform"]:
                lib.call("pai_adam_prepare", _ptr(ent["step"]), lr, beta1, beta2, _ptr(ent["dyn"]), stream)
                self._launch(items, beta1, beta2, 0.0, 0.0, eps, ent["dyn"], stream)
                continue
            by_step = {}
            for p, g, st in items:
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append((p, g, st))
            for t, sub in by_step.items():
                self._launch(sub, beta1, beta2, lr / (1.0 - beta1 ** t), 1.0 / math.sqrt(1.0 - beta2 ** t), eps, None,
                             stream)
        return loss
