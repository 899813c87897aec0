"""ORACLE (test infrastructure): 200-step training-loss curves of the UNMODIFIED reference
(models.pix2pix.Pix2Pix.training_step through the oracle/shim Lightning stand-in), for the north-star check
"loss curves over 200 synthetic steps within 2 %" (compared as 25-step window means on the bf16
path).  Build container only:

    python oracle/gen_golden_curves.py        ->  tests/golden/curves_ref.npz

Data: 8 synthetic batches of 4 pairs (oracle/pix2pix_port.synthetic_pairs, seeds 1000..1007) cycled in order.
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
from models.wrapper import Discriminator  # noqa: E402
from models.utils import init_weights  # noqa: E402

STEPS, BATCH, NBATCH = 200, 4, 8


def run(loss_type):
    torch.manual_seed(0)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type=loss_type)
    if loss_type == "gan":
        m.discriminator = Discriminator(in_channels=1)
        m.discriminator.apply(init_weights)
    m.train()
    data = [port.synthetic_pairs(BATCH, seed=1000 + i) for i in range(NBATCH)]
    for i in range(STEPS):
        m.training_step(data[i % NBATCH], i)
    return {k: np.array(v, dtype=np.float64) for k, v in m.logged.items()}


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    for lt in ("ssim+psnr", "gan"):
        for k, v in run(lt).items():
            out[f"{lt}/{k}"] = v
        print(lt, {k: (float(v[:25].mean()), float(v[-25:].mean())) for k, v in out.items() if k.startswith(lt)})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "curves_ref.npz"), **out)


if __name__ == "__main__":
    main()
