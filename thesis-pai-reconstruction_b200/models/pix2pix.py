"""Mirror of /root/reference/models/pix2pix.py (``Pix2Pix``, ``EncoderBlock``, ``DecoderBlock``,
``Unet``) on the B200 kernels.  The sub-modules only hold parameters under the reference's
``state_dict`` names (``encoders.{i}.encode.{1,2}``, ``decoders.{j}.decode.{1,2}``, ...); the whole
generator forward/backward is one fused autograd node (pai_b200.engine.UnetFunction)."""
from typing import Literal, Sequence

import torch
import torch.nn as nn

from pai_b200 import engine, lib

from .wrapper import UnetWrapper

_K = dict(kernel_size=4, stride=2, padding=1)


class Pix2Pix(UnetWrapper):
    """pix2pix (Isola et al. 2018) generator + PatchGAN, constructor as models/pix2pix.py:25-32."""

    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        channel_mults: Sequence[int] = (1, 2, 4, 8, 8, 8, 8, 8),
        dropout: float = 0.5,
        loss_type: Literal["gan", "ssim", "psnr", "ssim+psnr", "mse"] = "gan",
    ):
        unet = Unet(in_channels, out_channels, channel_mults=channel_mults, dropout=dropout)
        super().__init__(unet, loss_type=loss_type)
        self.example_input_array = torch.Tensor(2, in_channels, 256, 256)
        self.save_hyperparameters()


class EncoderBlock(nn.Module):
    """LeakyReLU(0.2) -> Conv4x4 s2 p1 -> BatchNorm (Identity at the bottleneck)."""

    def __init__(self, in_channels: int, out_channels: int, norm: bool = True):
        super().__init__()
        self.encode = nn.Sequential(
            nn.LeakyReLU(0.2),
            nn.Conv2d(in_channels, out_channels, **_K),
            nn.BatchNorm2d(out_channels) if norm else nn.Identity(),
        )

    def forward(self, x):
        raise RuntimeError("pai_b200: EncoderBlock is executed as part of Unet.forward")


class DecoderBlock(nn.Module):
    """ReLU -> ConvT4x4 s2 p1 -> BatchNorm -> Dropout2d."""

    def __init__(self, in_channels: int, out_channels: int, dropout: float = 0.5):
        super().__init__()
        self.decode = nn.Sequential(
            nn.ReLU(),
            nn.ConvTranspose2d(in_channels, out_channels, **_K),
            nn.BatchNorm2d(out_channels),
            nn.Dropout2d(dropout) if dropout > 0 else nn.Identity(),
        )
        self.dropout = dropout

    def forward(self, x):
        raise RuntimeError("pai_b200: DecoderBlock is executed as part of Unet.forward")


class Unet(nn.Module):
    """U-net generator, ``[N, in, H, W] -> [N, out, H, W]`` in (-1, 1)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 channel_mults: Sequence[int] = (1, 2, 4, 8, 8, 8, 8, 8), dropout: float = 0.5):
        super().__init__()
        widths = [64 * m for m in channel_mults]
        depth = len(widths)
        down = [nn.Conv2d(in_channels, widths[0], **_K)]
        for lvl in range(1, depth):
            down.append(EncoderBlock(widths[lvl - 1], widths[lvl], norm=lvl != depth - 1))
        self.encoders = nn.ModuleList(down)

        up, cin = [], widths[-1]
        widest = max(channel_mults)
        for lvl in range(depth - 2, -1, -1):
            drop = dropout if (channel_mults[lvl] == widest and lvl > depth - 5) else 0
            up.append(DecoderBlock(cin, widths[lvl], dropout=drop))
            cin = 2 * widths[lvl]
        up.append(nn.ConvTranspose2d(cin, out_channels, **_K))
        self.decoders = nn.ModuleList(up)
        self.out = nn.Tanh()
        self._spec = None

    def _engine_spec(self):
        if self._spec is None:
            enc_convs = [self.encoders[0]] + [b.encode[1] for b in list(self.encoders)[1:]]
            enc_bns = [None] + [b.encode[2] if isinstance(b.encode[2], nn.BatchNorm2d) else None
                                for b in list(self.encoders)[1:]]
            blocks = list(self.decoders)
            dec_convs = [b.decode[1] for b in blocks[:-1]] + [blocks[-1]]
            dec_bns = [b.decode[2] for b in blocks[:-1]] + [None]
            dec_dropout = [float(getattr(b, "dropout", 0) or 0) for b in blocks]
            self._spec = engine.UnetSpec(enc_convs, enc_bns, dec_convs, dec_bns, dec_dropout)
        return self._spec

    def forward(self, x):
        with lib.on_device(x):
            return self._forward(x)

    def _forward(self, x):
        spec = self._engine_spec()
        if engine.check_path_enabled():
            return engine.unet_forward_check(spec, x, self.training)
        return engine.UnetFunction.apply(spec, self.training, x, *spec.params())
