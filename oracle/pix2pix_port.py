"""ORACLE (test infrastructure, not product code).

CPU restatement, in functional PyTorch, of the reference's Pix2Pix hot path:

* ``Unet.forward``            /root/reference/models/pix2pix.py:198-216 (ctor :130-196)
* ``EncoderBlock``            models/pix2pix.py:46-74   (LeakyReLU(0.2) -> Conv4x4 s2 p1 -> BN)
* ``DecoderBlock``            models/pix2pix.py:77-111  (ReLU -> ConvT4x4 s2 p1 -> BN -> Dropout2d)
* ``Discriminator``           models/wrapper.py:212-238 (cat -> 4x[Conv s2 + LeakyReLU] -> Conv s1 no bias)
* ``UnetWrapper.loss``        models/wrapper.py:42-66
* ``discriminator_loss``      models/wrapper.py:68-95
* ``configure_optimizers``    models/wrapper.py:97-115  (Adam 2e-4, betas (0.5,0.999), eps 1e-7)
* ``training_step``           models/wrapper.py:117-162 (D step on a graph-free G forward, then G step)
* ``validation_step``         models/wrapper.py:164-173
* ``report.py`` metric sweep  report.py:72-101,144-146,188-217

It is written against a flat ``state_dict`` that uses the reference's key names
and tensor layouts, so reference checkpoints load unchanged.  The reference is
pure Python and cannot travel to the GPU box (/root/reference does not exist
there); this port does.  It is pinned against the real reference modules by
``oracle/gen_golden.py`` -> ``tests/golden/*.npz`` and, in this container, by a
direct comparison in ``tests/test_oracle_models.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product never does.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torchmetrics_port as tm  # noqa: E402

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- utils (models/utils.py)
def denormalize(x: Tensor) -> Tensor:
    """models/utils.py:11"""
    return torch.clamp(x * 0.5 + 0.5, 0, 1)


def ssim(pred: Tensor, target: Tensor) -> Tensor:
    """models/utils.py:38-39"""
    return tm.structural_similarity_index_measure(pred, target, data_range=1.0)


def psnr(pred: Tensor, target: Tensor) -> Tensor:
    """models/utils.py:42-43"""
    return tm.peak_signal_noise_ratio(pred, target, data_range=1.0)


def rmse(pred: Tensor, target: Tensor) -> Tensor:
    """models/utils.py:46-47"""
    return tm.mean_squared_error(pred, target, squared=False)


# --------------------------------------------------------------------------- parameter construction
def init_state(
    seed: int,
    in_channels: int = 1,
    out_channels: int = 1,
    channel_mults: Sequence[int] = (1, 2, 4, 8, 8, 8, 8, 8),
    loss_type: str = "gan",
    disc_in_channels: Optional[int] = None,
) -> Dict[str, Tensor]:
    """Replays the reference's construction order so that, under the same
    ``torch.manual_seed``, every tensor equals the reference's bit for bit:
    ``Unet.__init__`` (pix2pix.py:140-196) creates the layers in order with
    PyTorch default init; ``UnetWrapper.__init__`` (wrapper.py:32-37) then builds
    ``Discriminator()`` with its default ``in_channels=3`` and applies
    ``init_weights`` (utils.py:15-28) to it and then to the U-Net.
    ``disc_in_channels`` mirrors what the benchmark harness must do for
    1-channel data: construct ``Discriminator(in_channels=1)``
    afterwards, apply ``init_weights``, and assign it."""
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def put(prefix: str, mod: nn.Module):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    mods: List[Tuple[str, nn.Module]] = []
    ch = channel_mults[0] * 64
    mods.append(("unet.encoders.0", nn.Conv2d(in_channels, ch, 4, 2, 1)))
    cin = ch
    last = len(channel_mults) - 1
    for level, mult in enumerate(channel_mults[1:], 1):
        c = mult * 64
        mods.append((f"unet.encoders.{level}.encode.1", nn.Conv2d(cin, c, 4, 2, 1)))
        if level != last:
            mods.append((f"unet.encoders.{level}.encode.2", nn.BatchNorm2d(c)))
        cin = c
    idx = 0
    for level, mult in reversed(list(enumerate(channel_mults[:-1]))):
        c = mult * 64
        mods.append((f"unet.decoders.{idx}.decode.1", nn.ConvTranspose2d(cin, c, 4, 2, 1)))
        mods.append((f"unet.decoders.{idx}.decode.2", nn.BatchNorm2d(c)))
        cin = c * 2
        idx += 1
    mods.append((f"unet.decoders.{idx}", nn.ConvTranspose2d(cin, out_channels, 4, 2, 1)))

    def disc(nin: int) -> List[Tuple[str, nn.Module]]:
        d = []
        chans = [nin * 2, 64, 128, 256, 512]
        for i in range(4):
            d.append((f"discriminator.discriminator.{i}.block.0", nn.Conv2d(chans[i], chans[i + 1], 4, 2, 1)))
        d.append(("discriminator.discriminator.4", nn.Conv2d(512, 1, 4, padding=1, bias=False)))
        return d

    def reinit(ms):
        # models/utils.py:15-28, module.apply() visits in registration order
        for _, m in ms:
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, 0.0, 0.02)
            if isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.0)

    dmods: List[Tuple[str, nn.Module]] = []
    if loss_type == "gan":
        dmods = disc(3)
        reinit(dmods)
    reinit(mods)
    if loss_type == "gan" and disc_in_channels is not None and disc_in_channels != 3:
        dmods = disc(disc_in_channels)
        reinit(dmods)
    for name, m in mods + dmods:
        put(name, m)
    return sd


def split_state(sd: Dict[str, Tensor]):
    g = {k: v for k, v in sd.items() if k.startswith("unet.")}
    d = {k: v for k, v in sd.items() if k.startswith("discriminator.")}
    return g, d


def _is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


# --------------------------------------------------------------------------- forward passes
def _bn(sd, prefix: str, h: Tensor, training: bool) -> Tensor:
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    out = F.batch_norm(h, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def unet_forward(sd: Dict[str, Tensor], x: Tensor, training: bool, n_levels: int = 8) -> Tensor:
    """models/pix2pix.py:198-216.  ``sd`` BN buffers are updated in place when training."""
    h = x.to(torch.float32)
    feats = []
    for i in range(n_levels):
        if i == 0:
            h = F.conv2d(h, sd["unet.encoders.0.weight"], sd["unet.encoders.0.bias"], 2, 1)
        else:
            p = f"unet.encoders.{i}.encode"
            h = F.leaky_relu(h, 0.2)
            h = F.conv2d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], 2, 1)
            if p + ".2.weight" in sd:
                h = _bn(sd, p + ".2", h, training)
        feats.append(h)
    feats.pop()
    for i in range(n_levels):
        if i:
            h = torch.cat([h, feats.pop()], 1)
        if i < n_levels - 1:
            p = f"unet.decoders.{i}.decode"
            h = F.relu(h)
            h = F.conv_transpose2d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], 2, 1)
            h = _bn(sd, p + ".2", h, training)
        else:
            p = f"unet.decoders.{i}"
            h = F.conv_transpose2d(h, sd[p + ".weight"], sd[p + ".bias"], 2, 1)
    return torch.tanh(h)


def disc_forward(sd: Dict[str, Tensor], x: Tensor, y: Tensor) -> Tensor:
    """models/wrapper.py:236-238 with blocks :196-206,228-234 (norm=False -> Identity)."""
    h = torch.cat([x, y], 1)
    for i in range(4):
        p = f"discriminator.discriminator.{i}.block.0"
        h = F.leaky_relu(F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], 2, 1), 0.2)
    return F.conv2d(h, sd["discriminator.discriminator.4.weight"], None, 1, 1)


# --------------------------------------------------------------------------- losses
def generator_loss(sd, loss_type: str, x: Tensor, pred: Tensor, target: Tensor) -> Tensor:
    """models/wrapper.py:42-66"""
    if loss_type == "gan":
        lab = disc_forward(sd, x, pred)
        bce = F.binary_cross_entropy_with_logits(lab, torch.ones_like(lab))
        return bce + 50 * F.l1_loss(pred, target)
    if loss_type == "ssim":
        return -ssim(denormalize(pred), denormalize(target))
    if loss_type == "psnr":
        return -psnr(denormalize(pred), denormalize(target))
    if loss_type == "ssim+psnr":
        return -(30 * ssim(denormalize(pred), denormalize(target)) + psnr(denormalize(pred), denormalize(target)))
    if loss_type == "mse":
        return F.mse_loss(pred, target)
    raise ValueError(loss_type)


def discriminator_loss(pred_label: Tensor, target_label: Tensor) -> Tensor:
    """models/wrapper.py:68-95"""
    return F.binary_cross_entropy_with_logits(pred_label, torch.zeros_like(pred_label)) + \
        F.binary_cross_entropy_with_logits(target_label, torch.ones_like(pred_label))


# --------------------------------------------------------------------------- trainer
class OracleTrainer:
    """State + ``training_step`` of ``UnetWrapper`` (models/wrapper.py:97-173) on CPU fp32."""

    def __init__(self, sd: Dict[str, Tensor], loss_type: str = "gan", n_levels: int = 8):
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.loss_type = loss_type
        self.n_levels = n_levels
        self.g_keys = [k for k in self.sd if k.startswith("unet.") and _is_param(k)]
        self.d_keys = [k for k in self.sd if k.startswith("discriminator.") and _is_param(k)]
        for k in self.g_keys + self.d_keys:
            self.sd[k].requires_grad_(True)
        adam = dict(lr=2e-4, betas=(0.5, 0.999), eps=1e-7)
        self.opt_g = torch.optim.Adam([self.sd[k] for k in self.g_keys], **adam)
        self.opt_d = torch.optim.Adam([self.sd[k] for k in self.d_keys], **adam) if self.d_keys else None
        self.logged: Dict[str, List[float]] = {}

    def _log(self, k, v):
        self.logged.setdefault(k, []).append(float(v.detach()))

    def _freeze(self, keys, flag):
        for k in keys:
            self.sd[k].requires_grad_(flag)

    def forward(self, x: Tensor, training: bool = False) -> Tensor:
        return unet_forward(self.sd, x, training, self.n_levels)

    def training_step(self, x: Tensor, target: Tensor) -> Tensor:
        sd = self.sd
        if self.loss_type == "gan":
            self._freeze(self.g_keys, False)          # toggle_optimizer(opt_d), wrapper.py:124
            pred = unet_forward(sd, x, True, self.n_levels)   # :126 (BN running stats advance)
            target_label = disc_forward(sd, x, target)
            pred_label = disc_forward(sd, x, pred)
            d_loss = discriminator_loss(pred_label, target_label)
            self._log("d_loss", d_loss)
            self.opt_d.zero_grad(set_to_none=True)
            d_loss.backward()
            self.opt_d.step()
            self._freeze(self.g_keys, True)           # untoggle
            self._freeze(self.d_keys, False)          # toggle_optimizer(opt_g), :145
        pred = unet_forward(sd, x, True, self.n_levels)       # :147
        loss = generator_loss(sd, self.loss_type, x, pred, target)
        dp, dt = denormalize(pred), denormalize(target)
        self._log("loss", loss)
        self._log("train_ssim", ssim(dp, dt))
        self._log("train_psnr", psnr(dp, dt))
        self._log("train_rmse", rmse(dp, dt))
        self.opt_g.zero_grad(set_to_none=True)
        loss.backward()
        self.opt_g.step()
        if self.loss_type == "gan":
            self._freeze(self.d_keys, True)
        return loss.detach()

    @torch.no_grad()
    def validation_step(self, x: Tensor, target: Tensor) -> Dict[str, float]:
        pred = unet_forward(self.sd, x, False, self.n_levels)
        dp, dt = denormalize(pred), denormalize(target)
        return {"val_ssim": float(ssim(dp, dt)), "val_psnr": float(psnr(dp, dt)), "val_rmse": float(rmse(dp, dt))}


# --------------------------------------------------------------------------- report.py metric sweep
def depth_ssim(preds: Tensor, targets: Tensor, num_depths: int = 16) -> Tensor:
    """report.py:188-217 -- SSIM per row band, (mean, unbiased std) over images."""
    out = []
    for xp, xt in zip(preds.chunk(num_depths, dim=2), targets.chunk(num_depths, dim=2)):
        s = tm.structural_similarity_index_measure(xp, xt, data_range=1.0, reduction="none")
        out.append((s.mean(), s.std()))
    return torch.tensor(out)


def report_metrics(preds: Tensor, targets: Tensor, chunk: int = 64, want_maps: bool = True):
    """report.py:72-104,144-146 -- per-image SSIM (+full maps), PSNR, MSE in chunks of 64,
    SSIM over depth, dataset means and the global RMSE."""
    ssims, maps, psnrs, mses = [], [], [], []
    for p, t in zip(preds.split(chunk), targets.split(chunk)):
        s, m = tm.structural_similarity_index_measure(
            p, t, data_range=1.0, return_full_image=True, reduction="none")
        ssims.append(s)
        if want_maps:
            maps.append(m)
        psnrs.append(torch.tensor([tm.peak_signal_noise_ratio(a, b, data_range=1.0) for a, b in zip(p, t)]))
        mses.append(torch.tensor([tm.mean_squared_error(a, b) for a, b in zip(p, t)]))
    ssims, psnrs, mses = torch.cat(ssims), torch.cat(psnrs), torch.cat(mses)
    return {
        "ssim": ssims,
        "psnr": psnrs,
        "mse": mses,
        "ssim_maps": torch.cat(maps) if want_maps else None,
        "depth_ssim": depth_ssim(preds, targets),
        "ssim_mean": ssims.mean(),
        "psnr_mean": psnrs.mean(),
        "rmse": tm.mean_squared_error(preds, targets, squared=False),
    }


# --------------------------------------------------------------------------- synthetic data
def synthetic_pairs(n: int, seed: int = 1234, size: int = 256):
    """Normalised ([-1,1]) training pairs: smooth target + noisy input."""
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(n, 1, 32, 32, generator=g), size=size, mode="bilinear")
    target = 2 * base - 1
    x = (target + 0.5 * torch.randn(target.shape, generator=g)).clamp(-1, 1)
    return x, target


def synthetic_eval_pairs(n: int, seed: int = 4321, size: int = 256):
    """De-normalised ([0,1]) evaluation pairs for the report.py sweep."""
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(n, 1, 32, 32, generator=g), size=size, mode="bilinear")
    pred = (base + 0.05 * torch.randn(base.shape, generator=g)).clamp(0, 1)
    return pred, base
