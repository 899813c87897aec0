"""ORACLE (test infrastructure): generates ``tests/golden/*.npz`` by running the
UNMODIFIED reference from /root/reference (through the ``oracle/shim`` stand-ins for
``pytorch_lightning`` / ``torchmetrics``, which are not installed here).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

Every array in the fixtures is an output of the reference's own classes/functions:
``models.pix2pix.Pix2Pix`` (+ ``models.wrapper.Discriminator(in_channels=1)``,
SURVEY.md Q1), ``models.utils.{ssim,psnr,rmse,denormalize}`` and
``report.depth_ssim``.  The inputs are regenerated from seeds by
``oracle/pix2pix_port.synthetic_pairs`` so only small outputs are committed.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# report.py imports plotting / flop-count packages that are not installed and are not on the
# metric path; stub them so ``report.depth_ssim`` can be imported unmodified.
for name in ("matplotlib", "fvcore", "fvcore.nn", "dataset"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].colormaps = {}
sys.modules["fvcore.nn"].FlopCountAnalysis = object
sys.modules["dataset"].ImageDataModule = object
for name in ("models.palette", "models.attention_unet", "models.res_unet", "models.trans_unet"):
    pass  # these import fine through the shim (einops is installed)

import pix2pix_port as port  # noqa: E402
from models.pix2pix import Pix2Pix  # noqa: E402
from models.wrapper import Discriminator  # noqa: E402
from models.utils import init_weights, ssim, psnr, rmse, denormalize  # noqa: E402


def build_reference(seed, loss_type):
    torch.manual_seed(seed)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type=loss_type)
    if loss_type == "gan":
        m.discriminator = Discriminator(in_channels=1)
        m.discriminator.apply(init_weights)
    return m


def checksums(sd):
    keys = sorted(sd.keys())
    return keys, np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    N = 2
    x, target = port.synthetic_pairs(N, seed=1234)

    # ---- model fixtures (gan)
    m = build_reference(0, "gan")
    keys, cs = checksums(m.state_dict())
    out["state_keys"] = np.array(keys)
    out["state_checksums"] = cs
    m.eval()
    with torch.no_grad():
        y_eval = m(x)
        out["gen_eval_sub"] = y_eval[:, :, ::4, ::4].numpy()
        out["gen_eval_stats"] = np.array([float(y_eval.mean()), float(y_eval.std()), float(y_eval.abs().max())])
        out["disc_logits"] = m.discriminator(x, target).numpy()
    m.train()
    # one train-mode forward on a throw-away copy (BN batch statistics)
    m2 = build_reference(0, "gan")
    m2.train()
    y_train = m2(x)
    out["gen_train_sub"] = y_train.detach()[:, :, ::4, ::4].numpy()
    loss = m2.loss(x, y_train, target)
    loss.backward()
    gkeys = [k for k, p in m2.named_parameters() if p.grad is not None]
    out["grad_keys"] = np.array(gkeys)
    out["grad_norms"] = np.array([float(dict(m2.named_parameters())[k].grad.double().norm()) for k in gkeys])
    out["gan_gloss0"] = np.array(float(loss))

    # three full GAN training steps
    for _ in range(3):
        m.training_step((x, target), 0)
    for k, v in m.logged.items():
        out[f"gan_log_{k}"] = np.array(v)
    _, cs_after = checksums(m.state_dict())
    out["state_checksums_after3"] = cs_after

    # ---- ssim+psnr loss type, three steps
    m3 = build_reference(0, "ssim+psnr")
    m3.train()
    for _ in range(3):
        m3.training_step((x, target), 0)
    for k, v in m3.logged.items():
        out[f"sp_log_{k}"] = np.array(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pix2pix_ref.npz"), **out)
    print("pix2pix_ref.npz", {k: getattr(v, "shape", None) for k, v in out.items()})

    # ---- metric fixtures
    import report  # noqa: E402  (reference report.py, unmodified)
    met = {}
    pred, tgt = port.synthetic_eval_pairs(8, seed=4321)
    met["ssim"] = np.array(float(ssim(pred, tgt)))
    met["psnr"] = np.array(float(psnr(pred, tgt)))
    met["rmse"] = np.array(float(rmse(pred, tgt)))
    s, full = report.ssim(pred, tgt, data_range=1.0, return_full_image=True, reduction="none")
    met["ssim_per_image"] = s.numpy()
    met["ssim_map_sub"] = full[:, :, ::8, ::8].numpy()
    met["ssim_map_img0"] = full[0, 0].numpy().astype(np.float16)
    met["psnr_per_image"] = np.array([float(report.psnr(p, t, data_range=1.0)) for p, t in zip(pred, tgt)])
    met["mse_per_image"] = np.array([float(report.mse(p, t)) for p, t in zip(pred, tgt)])
    met["depth_ssim"] = report.depth_ssim(pred, tgt).numpy()
    met["rmse_global"] = np.array(float(report.mse(pred, tgt, squared=False)))
    # normalised-domain variant as used in training (denormalize of [-1,1] tensors)
    xn, tn = port.synthetic_pairs(4, seed=77)
    met["train_ssim"] = np.array(float(ssim(denormalize(xn), denormalize(tn))))
    met["train_psnr"] = np.array(float(psnr(denormalize(xn), denormalize(tn))))
    met["train_rmse"] = np.array(float(rmse(denormalize(xn), denormalize(tn))))
    # gradient of the ssim+psnr loss wrt the normalised prediction (wrapper.py:59-63)
    xr = xn.clone().requires_grad_(True)
    l = -(30 * ssim(denormalize(xr), denormalize(tn)) + psnr(denormalize(xr), denormalize(tn)))
    l.backward()
    met["sp_loss"] = np.array(float(l))
    met["sp_grad_sub"] = xr.grad[:, :, ::8, ::8].numpy()
    met["sp_grad_norm"] = np.array(float(xr.grad.double().norm()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics_ref.npz"), **met)
    print("metrics_ref.npz", {k: getattr(v, "shape", None) for k, v in met.items()})


if __name__ == "__main__":
    main()
