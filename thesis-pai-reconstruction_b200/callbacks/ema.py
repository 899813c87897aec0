"""Mirror of /root/reference/callbacks/ema.py: ``EMACallback`` keeps an exponential moving average of every parameter
(decay 0.9999, main.py:131), swaps it in for validation and restores the live weights afterwards.

The reference delegates the arithmetic to ``torch_ema.ExponentialMovingAverage`` (requirements.txt; not installed in this
image and not vendored: restated here from its documented behaviour, *unpinned* against the package):

    num_updates += 1;  d = min(decay, (1 + num_updates) / (10 + num_updates))
    shadow <- shadow - (1 - d) * (shadow - param)          for every parameter that requires grad

``update()`` is ONE launch of the library's multi-tensor kernel (``pai_ema_multi``, csrc/optim.cu: 12 bytes per parameter)
over all CUDA fp32 parameters -- the adjacent elementwise pass over the 54 M generator parameters of SURVEY.md 8(f)-3;
parameters on the host take ``torch._foreach`` ops (plumbing, like FusedAdam's CPU path).  The hooks have the reference's names and order, so ``pl.Trainer(callbacks=[EMACallback(0.9999)])`` works when
Lightning is installed; without it the class is a plain object with the same methods.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

try:  # pragma: no cover - exercised only where Lightning exists
    import pytorch_lightning as pl
    _Base = pl.callbacks.Callback
except ImportError:
    _Base = object


class ExponentialMovingAverage:
    """The subset of ``torch_ema.ExponentialMovingAverage`` that callbacks/ema.py uses."""

    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates: Optional[int] = 0 if use_num_updates else None
        self._params: List[torch.nn.Parameter] = [p for p in parameters if p.requires_grad]
        self.shadow_params = [p.detach().clone() for p in self._params]
        self.collected_params: Optional[List[torch.Tensor]] = None

    @torch.no_grad()
    def update(self) -> None:
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        live = [p.detach() for p in self._params]
        if live and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and s.is_contiguous()
                        for p, s in zip(live, self.shadow_params)):
            import ctypes
            from pai_b200 import lib
            n = len(live)
            arr = ctypes.c_void_p * n
            with lib.on_device(live[0]):
                lib.call("pai_ema_multi", n, arr(*[s.data_ptr() for s in self.shadow_params]),
                         arr(*[p.data_ptr() for p in live]), (ctypes.c_int * n)(*[p.numel() for p in live]),
                         one_minus_decay, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), kernels=(n + 47) // 48)
            return
        diff = torch._foreach_sub(self.shadow_params, live)             # shadow - param
        torch._foreach_add_(self.shadow_params, diff, alpha=-one_minus_decay)

    @torch.no_grad()
    def store(self) -> None:
        self.collected_params = [p.detach().clone() for p in self._params]

    @torch.no_grad()
    def copy_to(self) -> None:
        for s, p in zip(self.shadow_params, self._params):
            p.copy_(s)

    @torch.no_grad()
    def restore(self) -> None:
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, self._params):
            p.copy_(c)
        self.collected_params = None

    def state_dict(self) -> dict:
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state: dict) -> None:
        self.decay = state["decay"]
        self.num_updates = state["num_updates"]
        self.shadow_params = [s.to(p.device, p.dtype).clone() for s, p in zip(state["shadow_params"], self._params)]
        cp = state.get("collected_params")
        self.collected_params = None if cp is None else [c.to(p.device, p.dtype).clone() for c, p in zip(cp, self._params)]


class EMACallback(_Base):
    """Exponential Moving Average callback (callbacks/ema.py:5-72): same constructor and hooks."""

    def __init__(self, decay=0.9999):
        self.decay = decay
        self.ema = None

    def on_fit_start(self, trainer, pl_module):
        self.ema = ExponentialMovingAverage(pl_module.parameters(), decay=self.decay)

    def on_train_batch_end(self, trainer, pl_module, *args, **kwargs):
        self.ema.update()

    def on_validation_start(self, trainer, pl_module):
        self.ema.store()
        self.ema.copy_to()

    def on_validation_end(self, trainer, pl_module):
        self.ema.restore()

    def on_save_checkpoint(self, trainer, pl_module, checkpoint):
        return self.ema.state_dict()

    def on_load_checkpoint(self, trainer, pl_module, callback_state):
        self.ema.load_state_dict(callback_state)
