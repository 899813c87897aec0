"""Mirror of /root/reference/models/wrapper.py (``UnetWrapper``, ``DiscriminatorBlock``,
``Discriminator``) on the B200 kernels.

* the module tree, ``state_dict`` keys and constructor signatures are the reference's, so checkpoints
  and ``main.py`` / ``report.py`` work unchanged;
* ``Discriminator.forward`` is one fused autograd node (pai_b200.engine.DiscFunction);
* ``training_step`` keeps the reference's order of operations (wrapper.py:117-162: D step on a
  graph-free generator forward, then the G step) but evaluates the loss and the three logged metrics
  from ONE pass of the SSIM/PSNR/MSE kernel instead of four torchmetrics calls.
"""
import os
from typing import Literal

import torch
import torch.nn as nn
import torch.nn.functional as F

from pai_b200 import dp, engine, lib, metrics, ops
from pai_b200.optim import FusedAdam

from ._lightning import LightningModule
from .utils import denormalize, init_weights, psnr, rmse, ssim  # noqa: F401  (re-exported like the reference)

BATCH_DISCRIMINATOR_PASSES = os.environ.get("PAI_NO_D_BATCHING", "0") != "1"

_ADAM = dict(lr=2e-4, betas=(0.5, 0.999), eps=1e-7)        # wrapper.py:98-111


class UnetWrapper(LightningModule):
    """U-net wrapper with the five reference loss types: "gan", "ssim", "psnr", "ssim+psnr", "mse"."""

    def __init__(self, unet: nn.Module, loss_type: Literal["gan", "ssim", "psnr", "ssim+psnr", "mse"] = "gan",
                 discriminator_in_channels: int = None):
        """``discriminator_in_channels`` is the one addition to the reference signature (wrapper.py:21-25): the
        reference always builds ``Discriminator()`` = 3 channels per image (wrapper.py:34), which cannot run on the
        1-channel models main.py constructs (its first conv wants 6 planes, cat([x, y]) has 2).  ``None``
        keeps that behaviour bit for bit (same RNG consumption, same state_dict); pass 1 to get a working grayscale GAN
        without swapping ``self.discriminator`` by hand."""
        super().__init__()
        self.automatic_optimization = False
        self.unet = unet
        self.loss_type = loss_type
        self.discriminator = None
        if loss_type == "gan":
            # same construction order as the reference so a shared seed gives identical weights
            self.discriminator = (Discriminator() if discriminator_in_channels is None
                                  else Discriminator(in_channels=discriminator_in_channels))
            self.discriminator.apply(init_weights)
        self.unet.apply(init_weights)

    def forward(self, x):
        return self.unet(x)

    def manual_backward(self, loss, *args, **kwargs):
        """Backward, then (data-parallel runs only) average the gradients of the parameters being
        trained over all ranks with one flat NCCL all-reduce -- the exchange Lightning's DDP strategy
        would perform for main.py:123-136.  BatchNorm statistics stay per replica."""
        super().manual_backward(loss, *args, **kwargs)
        dp.allreduce_gradients(p for p in self.parameters() if p.requires_grad)

    # ---- losses ---------------------------------------------------------------------------------
    def loss(self, x, pred, target):
        kind = self.loss_type
        if kind == "gan":
            pred_label = self.discriminator(x, pred)
            adversarial = F.binary_cross_entropy_with_logits(pred_label, torch.ones_like(pred_label))
            return adversarial + 50 * F.l1_loss(pred, target)
        if kind == "mse":
            return F.mse_loss(pred, target)
        if kind in ("ssim", "psnr", "ssim+psnr"):
            s, p, _ = metrics.train_metrics(pred, target, denormalize=True)
            return {"ssim": -s, "psnr": -p, "ssim+psnr": -(30 * s + p)}[kind]
        return None

    def discriminator_loss(self, pred_label: torch.Tensor, target_label: torch.Tensor) -> torch.Tensor:
        fake = F.binary_cross_entropy_with_logits(pred_label, torch.zeros_like(pred_label))
        real = F.binary_cross_entropy_with_logits(target_label, torch.ones_like(pred_label))
        return fake + real

    def configure_optimizers(self):
        # FusedAdam is torch.optim.Adam (same hyper-parameters, state_dict and arithmetic) stepped by the
        # fused update + weight-repack kernels; on CPU parameters it IS torch.optim.Adam
        opt_g = FusedAdam(self.unet.parameters(), **_ADAM)
        if self.discriminator is None:
            return opt_g
        return opt_g, FusedAdam(self.discriminator.parameters(), **_ADAM)

    # ---- steps ----------------------------------------------------------------------------------
    def enable_step_graph(self, warmup: int = 3):
        """Opt-in: run ``training_step`` as ONE replayed CUDA graph per input shape (pai_b200.graph.StepGraph) after
        ``warmup`` eager calls.  Same arithmetic and the same logged values; removes the host launch overhead of the
        ~520 kernels of a GAN step.  ``disable_step_graph()`` returns to eager launches."""
        from pai_b200.graph import StepGraph
        self.__dict__["_pai_step_graph"] = StepGraph(self, warmup=warmup)
        return self

    def disable_step_graph(self):
        self.__dict__["_pai_step_graph"] = None
        return self

    def training_step(self, batch, batch_idx):
        runner = self.__dict__.get("_pai_step_graph")
        with lib.on_device(batch[0]):
            if runner is not None:
                return runner(batch, batch_idx)
            return self._training_step_eager(batch, batch_idx)

    def _pai_log(self, name, value):
        """``self.log`` -- except while pai_b200.graph.StepGraph captures the step: then the (static) tensor is handed
        to the capture, which logs a clone after every replay.  Nothing here depends on the fallback shim: under real
        pytorch_lightning ``self.log`` would call into the Trainer, which must not happen on a capturing stream."""
        sink = self.__dict__.get("_pai_log_sink")
        if sink is not None:
            sink.append((name, value.detach()))
        else:
            self.log(name, value, prog_bar=True)

    def _training_step_eager(self, batch, batch_idx):
        with ops.zero_pool(batch[0].device):        # one fill for the step's small zeroed temporaries
            return self._training_step_body(batch, batch_idx)

    def _training_step_body(self, batch, batch_idx):
        x, target = batch
        if self.loss_type == "gan":
            opt_d = self.optimizers()[1]
            self.toggle_optimizer(opt_d)            # generator frozen: its forward builds no graph
            pred = self.unet(x)
            if BATCH_DISCRIMINATOR_PASSES and x.is_cuda:
                # D(x, target) and D(x, pred) as ONE pass over 2N samples: the PatchGAN has no BatchNorm, so every sample's
                # logits and every gradient are what the two separate calls of models/wrapper.py:121-122 give, with half
                # the kernel launches and no gradient-accumulation adds  [PAI_NO_D_BATCHING=1: two calls]
                n = x.shape[0]
                both = self.discriminator(torch.cat([x, x]), torch.cat([target, pred.detach()]))
                target_label, pred_label = both[:n], both[n:]
            else:
                target_label = self.discriminator(x, target)
                pred_label = self.discriminator(x, pred)
            d_loss = self.discriminator_loss(pred_label, target_label)
            self._pai_log("d_loss", d_loss)
            self.discriminator.zero_grad(set_to_none=True)
            self.manual_backward(d_loss)
            opt_d.step()
            self.untoggle_optimizer(opt_d)

        opt_g = self.optimizers()
        if isinstance(opt_g, list):
            opt_g = opt_g[0]
        self.toggle_optimizer(opt_g)                # discriminator frozen: its backward is dgrad-only
        pred = self.unet(x)
        s, p, r = metrics.train_metrics(pred, target, denormalize=True)   # one kernel: loss terms + metrics
        kind = self.loss_type
        if kind == "ssim":
            loss = -s
        elif kind == "psnr":
            loss = -p
        elif kind == "ssim+psnr":
            loss = -(30 * s + p)
        else:
            loss = self.loss(x, pred, target)
        self._pai_log("loss", loss)
        self._pai_log("train_ssim", s)
        self._pai_log("train_psnr", p)
        self._pai_log("train_rmse", r)
        self.unet.zero_grad(set_to_none=True)
        self.manual_backward(loss)
        opt_g.step()
        self.untoggle_optimizer(opt_g)

    def validation_step(self, batch, batch_idx):
        x, target = batch
        pred = self.forward(x)
        s, p, r = metrics.train_metrics(pred.detach(), target, denormalize=True)
        self.log("val_ssim", s, prog_bar=True)
        self.log("val_psnr", p, prog_bar=True)
        self.log("val_rmse", r, prog_bar=True)


class DiscriminatorBlock(nn.Module):
    """Conv4x4 s2 p1 (+ optional InstanceNorm) + LeakyReLU(0.2) -- parameter holder; the arithmetic of
    a whole ``Discriminator`` runs in pai_b200.engine."""

    def __init__(self, in_channels: int, out_channels: int, norm: bool = False):
        super().__init__()
        if norm:
            raise RuntimeError("pai_b200: DiscriminatorBlock(norm=True) is never instantiated by the reference "
                               "(wrapper.py:192,228-232) and has no B200 kernel")
        self.block = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1),
            nn.Identity(),
            nn.LeakyReLU(0.2),
        )

    def forward(self, x):
        raise RuntimeError("pai_b200: DiscriminatorBlock is executed as part of Discriminator.forward")


class Discriminator(nn.Module):
    """PatchGAN discriminator, ``forward(x, y) -> logits [N, 1, H/16-1, W/16-1]``."""

    def __init__(self, in_channels: int = 3):
        super().__init__()
        widths = (64, 128, 256, 512)
        layers, cin = [], in_channels * 2
        for wd in widths:
            layers.append(DiscriminatorBlock(cin, wd))
            cin = wd
        layers.append(nn.Conv2d(cin, 1, kernel_size=4, padding=1, bias=False))
        self.discriminator = nn.Sequential(*layers)
        self._spec = None

    def _engine_spec(self):
        if self._spec is None:
            convs = [blk.block[0] for blk in list(self.discriminator)[:-1]] + [self.discriminator[-1]]
            self._spec = engine.DiscSpec(convs)
        return self._spec

    def forward(self, x, y):
        with lib.on_device(x):
            return self._forward(x, y)

    def _forward(self, x, y):
        spec = self._engine_spec()
        if engine.check_path_enabled():
            return engine.disc_forward_check(spec, x, y)
        return engine.DiscFunction.apply(spec, x, y, *spec.params())
