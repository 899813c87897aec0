"""``pl.LightningModule`` when pytorch_lightning is installed (the reference's requirements.txt:3),
otherwise a minimal base class implementing the manual-optimisation contract that
models/wrapper.py:117-173 relies on (Lightning 2.0: ``optimizers``, ``toggle_optimizer`` /
``untoggle_optimizer``, ``manual_backward``, ``log``, ``save_hyperparameters``, ``freeze``,
``load_from_checkpoint``), so the drop-in trains on a box without Lightning."""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

try:  # pragma: no cover - exercised only where Lightning exists
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except ImportError:
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.automatic_optimization = True
            self.logged = {}
            self._pai_optimizers = None
            self._pai_toggled = None
            self.hparams = {}

        # -- optimisers (manual optimisation)
        def optimizers(self):
            if self._pai_optimizers is None:
                cfg = self.configure_optimizers()
                self._pai_optimizers = list(cfg) if isinstance(cfg, (tuple, list)) else [cfg]
            opts = self._pai_optimizers
            return opts if len(opts) > 1 else opts[0]

        def toggle_optimizer(self, optimizer):
            """Freeze every parameter that `optimizer` does not own; remember the previous flags."""
            self.optimizers()
            own = {id(p) for grp in optimizer.param_groups for p in grp["params"]}
            state = {}
            for opt in self._pai_optimizers:
                for grp in opt.param_groups:
                    for p in grp["params"]:
                        if id(p) not in own and id(p) not in state:
                            state[id(p)] = (p, p.requires_grad)
                            p.requires_grad = False
            self._pai_toggled = state

        def untoggle_optimizer(self, optimizer):
            for p, flag in (self._pai_toggled or {}).values():
                p.requires_grad = flag
            self._pai_toggled = None

        def manual_backward(self, loss, *args, **kwargs):
            loss.backward(*args, **kwargs)

        def log(self, name, value, **_):
            v = value.detach() if isinstance(value, torch.Tensor) else value
            self.logged.setdefault(name, []).append(v)

        # -- hyper-parameters / checkpoints
        def save_hyperparameters(self, *_, **__):
            frame = inspect.currentframe().f_back
            loc = frame.f_locals
            sig = inspect.signature(type(loc["self"]).__init__)
            self.hparams = {k: loc[k] for k in sig.parameters if k != "self" and k in loc}

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            model = cls(**hp)
            model.load_state_dict(ckpt["state_dict"])
            return model

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")
