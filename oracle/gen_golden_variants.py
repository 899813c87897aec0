"""ORACLE (test infrastructure): golden vectors for the Residual / Attention / Trans U-Net variants, produced
by the UNMODIFIED reference classes of /root/reference/models/{res_unet,attention_unet,trans_unet}.py (through
the ``oracle/shim`` stand-ins for pytorch_lightning / torchmetrics).  Build container only:

    python oracle/gen_golden_variants.py      ->  tests/golden/variants_ref.npz

Per case: sorted ``state_dict`` keys + checksums, sub-sampled eval-mode and train-mode generator outputs, the
loss of one training step and the gradient norm of every parameter.  Inputs are regenerated from seeds by
``oracle/pix2pix_port.synthetic_pairs`` (cropped to the case's resolution).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pix2pix_port as port  # noqa: E402

# name -> (module, class, ctor kwargs, batch, resolution, loss_type)
CASES = {
    "res_next": ("models.res_unet", "ResUnetGAN", dict(res_type="next", channel_mults=(1, 2, 4, 8, 8, 8)), 2, 256, "ssim"),
    "res_18_small": ("models.res_unet", "ResUnetGAN", dict(res_type="18", channel_mults=(1, 2, 4, 8)), 4, 64, "ssim+psnr"),
    "res_v2_small": ("models.res_unet", "ResUnetGAN", dict(res_type="v2", channel_mults=(1, 2, 4, 8)), 4, 64, "mse"),
    "res_50_small": ("models.res_unet", "ResUnetGAN", dict(res_type="50", channel_mults=(1, 2, 4, 8)), 4, 64, "ssim+psnr"),
    "attention": ("models.attention_unet", "AttentionUnetGAN", dict(), 2, 256, "ssim"),
    "trans_small": ("models.trans_unet", "TransUnetGAN", dict(channel_mults=(1, 2, 2), patch_size=4), 2, 256, "ssim"),
    # the configurations BASELINE.json names (configs 3 and 4): the 8-level ResNeXt U-Net at batch 8, and the Trans U-Net
    # main.py:93-101 builds (mults 1,2,2,4,4, patch 4: d = 4096, 1.03 G parameters)
    "res_next_b8": ("models.res_unet", "ResUnetGAN", dict(res_type="next"), 8, 256, "ssim"),
    "trans_full": ("models.trans_unet", "TransUnetGAN", dict(channel_mults=(1, 2, 2, 4, 4), patch_size=4), 2, 256, "ssim"),
}


def case_inputs(n, res, seed=1234):
    x, t = port.synthetic_pairs(n, seed=seed)
    return x[:, :, :res, :res].contiguous(), t[:, :, :res, :res].contiguous()


def build(module, cls, kwargs, loss_type, seed=0):
    import importlib
    torch.manual_seed(seed)
    mod = importlib.import_module(module)
    return getattr(mod, cls)(in_channels=1, out_channels=1, dropout=0.0, loss_type=loss_type, **kwargs)


def main(only=None):
    torch.set_num_threads(os.cpu_count())
    path = os.path.join(ROOT, "tests", "golden", "variants_ref.npz")
    out = dict(np.load(path)) if (only and os.path.exists(path)) else {}
    for name, (module, cls, kwargs, n, res, loss_type) in CASES.items():
        if only and name not in only:
            continue
        x, target = case_inputs(n, res)
        m = build(module, cls, kwargs, loss_type)
        sd = m.state_dict()
        keys = sorted(sd.keys())
        out[f"{name}/state_keys"] = np.array(keys)
        out[f"{name}/state_checksums"] = np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])
        m.eval()
        with torch.no_grad():
            y = m(x)
        out[f"{name}/eval_sub"] = y[:, :, ::4, ::4].numpy()
        m.train()
        y = m(x)
        out[f"{name}/train_sub"] = y.detach()[:, :, ::4, ::4].numpy()
        # the reference against itself under bf16 autocast (fresh copy: BN running stats untouched): the
        # noise floor any bf16-operand implementation of this network has in train mode
        m16 = build(module, cls, kwargs, loss_type)
        m16.train()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            y16 = m16(x).float()
        gap = (y16 - y.detach()).abs()
        out[f"{name}/bf16_gap"] = np.array([float(gap.max()), float(gap.mean())])
        loss = m.loss(x, y, target)
        loss.backward()
        out[f"{name}/loss"] = np.array(float(loss))
        named = dict(m.named_parameters())
        gk = [k for k in sorted(named) if named[k].grad is not None]
        out[f"{name}/grad_keys"] = np.array(gk)
        out[f"{name}/grad_norms"] = np.array([float(named[k].grad.double().norm()) for k in gk])
        print(name, "loss", float(loss), "params", sum(p.numel() for p in m.parameters()), "bf16 gap", out[f"{name}/bf16_gap"])
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
