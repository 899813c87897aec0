"""ORACLE shim (test infrastructure, not product code).

A ~100-line stand-in for the slice of ``pytorch_lightning`` 2.0.x that the
reference touches, so that the UNMODIFIED reference modules under
/root/reference (``models/wrapper.py:4,9``, ``models/pix2pix.py:4``,
``dataset.py:11``, ``callbacks/ema.py:5``) import and step in this container,
where the real package is not installed.  Semantics restated from the Lightning
2.0 manual-optimisation contract:

* ``optimizers()``   -> the single optimizer, or the list when several were configured
* ``toggle_optimizer(opt)``   -> requires_grad=False on every parameter owned by the
  *other* optimizers, remembering the previous flags; ``untoggle_optimizer`` restores
* ``manual_backward(loss)``   -> ``loss.backward()``
* ``log(name, value)``        -> recorded into ``self.logged`` (a dict of lists)
"""
from __future__ import annotations

import inspect
from types import SimpleNamespace

import torch
import torch.nn as nn

__version__ = "2.0.2+oracle-shim"


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.automatic_optimization = True
        self.logged = {}
        self._opts = None
        self._saved_flags = {}
        self.hparams = SimpleNamespace()

    # ---- hyper-parameters / checkpoints
    def save_hyperparameters(self, *_, **__):
        frame = inspect.currentframe().f_back
        init_locals = frame.f_locals
        sig = inspect.signature(type(init_locals["self"]).__init__)
        hp = {k: init_locals[k] for k in sig.parameters if k != "self" and k in init_locals}
        self.hparams = SimpleNamespace(**hp)
        self._hparams_dict = hp

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu", **overrides):
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
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

    # ---- logging
    def log(self, name, value, **_):
        v = value.detach().float().item() if torch.is_tensor(value) else float(value)
        self.logged.setdefault(name, []).append(v)

    # ---- manual optimisation
    def _ensure_opts(self):
        if self._opts is None:
            o = self.configure_optimizers()
            self._opts = list(o) if isinstance(o, (tuple, list)) else [o]
        return self._opts

    def optimizers(self):
        o = self._ensure_opts()
        return o[0] if len(o) == 1 else o

    def toggle_optimizer(self, optimizer):
        flags = {}
        for opt in self._ensure_opts():
            for group in opt.param_groups:
                for p in group["params"]:
                    if p in flags:
                        continue
                    flags[p] = p.requires_grad
                    p.requires_grad = False
        for group in optimizer.param_groups:
            for p in group["params"]:
                p.requires_grad = flags[p]
        self._saved_flags = flags

    def untoggle_optimizer(self, optimizer):
        for opt in self._ensure_opts():
            if opt is optimizer:
                continue
            for group in opt.param_groups:
                for p in group["params"]:
                    if p in self._saved_flags:
                        p.requires_grad = self._saved_flags[p]
        self._saved_flags = {}

    def manual_backward(self, loss, *a, **k):
        loss.backward(*a, **k)


class LightningDataModule:
    def __init__(self):
        pass


class _Callback:
    pass


class _ModelCheckpoint(_Callback):
    def __init__(self, **kw):
        self.kw = kw


class _CSVLogger:
    def __init__(self, save_dir, name=None, **_):
        self.save_dir, self.name = save_dir, name


callbacks = SimpleNamespace(Callback=_Callback, ModelCheckpoint=_ModelCheckpoint)
loggers = SimpleNamespace(CSVLogger=_CSVLogger)


class Trainer:
    """Minimal fit loop: train batches through ``training_step`` only."""

    def __init__(self, max_epochs=1, max_steps=-1, **kw):
        self.max_epochs, self.max_steps, self.kw = max_epochs, max_steps, kw

    def fit(self, model, datamodule=None, train_dataloaders=None):
        loader = train_dataloaders if train_dataloaders is not None else datamodule.train_dataloader()
        step = 0
        model.train()
        for _ in range(self.max_epochs):
            for i, batch in enumerate(loader):
                model.training_step(batch, i)
                step += 1
                if 0 < self.max_steps <= step:
                    return
