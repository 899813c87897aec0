"""ORACLE (test infrastructure): ``tests/golden/pix2pix_rgb_ref.npz`` from the UNMODIFIED reference built with its
OWN default constructors -- ``Pix2Pix(in_channels=3, out_channels=3)`` and ``Discriminator()`` (in_channels=3: a
6-plane first convolution; models/pix2pix.py:25-27, models/wrapper.py:34,225).  Same shims and conventions as
oracle/gen_golden.py; run in the build container only:

    python oracle/gen_golden_rgb.py

Inputs: three independent grayscale synthetic pairs (oracle/pix2pix_port.synthetic_pairs, seeds 900 + channel) stacked
along the channel axis, regenerated from the seeds by the test.
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
from models.pix2pix import Pix2Pix  # noqa: E402
from models.utils import init_weights  # noqa: E402  (applied by the wrapper itself)


def rgb_pairs(n, seed=900):
    xs, ts = zip(*[port.synthetic_pairs(n, seed=seed + c) for c in range(3)])
    return torch.cat(xs, 1), torch.cat(ts, 1)


def build(seed, loss_type):
    torch.manual_seed(seed)
    return Pix2Pix(dropout=0.0, loss_type=loss_type)          # reference defaults: 3 -> 3 channels, Discriminator()


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    x, target = rgb_pairs(2)
    m = build(0, "gan")
    sd = m.state_dict()
    keys = sorted(sd.keys())
    out["state_keys"] = np.array(keys)
    out["state_checksums"] = np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])
    m.eval()
    with torch.no_grad():
        y = m(x)
        out["gen_eval_sub"] = y[:, :, ::4, ::4].numpy()
        out["disc_logits"] = m.discriminator(x, target).numpy()
    m2 = build(0, "gan")
    m2.train()
    yt = m2(x)
    out["gen_train_sub"] = yt.detach()[:, :, ::4, ::4].numpy()
    loss = m2.loss(x, yt, target)
    loss.backward()
    named = dict(m2.named_parameters())
    gkeys = [k for k, p in named.items() if p.grad is not None]
    out["grad_keys"] = np.array(gkeys)
    out["grad_norms"] = np.array([float(named[k].grad.double().norm()) for k in gkeys])
    out["gan_gloss0"] = np.array(float(loss))
    # discriminator gradients of one discriminator loss (fake = detached train-mode prediction)
    m2.zero_grad(set_to_none=True)
    dl = m2.discriminator_loss(m2.discriminator(x, yt.detach()), m2.discriminator(x, target))
    dl.backward()
    dnamed = dict(m2.discriminator.named_parameters())
    out["d_loss0"] = np.array(float(dl))
    out["d_grad_keys"] = np.array(list(dnamed.keys()))
    out["d_grad_norms"] = np.array([float(p.grad.double().norm()) for p in dnamed.values()])
    for _ in range(3):
        m.train()
        m.training_step((x, target), 0)
    for k, v in m.logged.items():
        out[f"gan_log_{k}"] = np.array(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pix2pix_rgb_ref.npz"), **out)
    print("pix2pix_rgb_ref.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
