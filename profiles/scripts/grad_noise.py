"""Run-to-run spread of the gradient parity of tests/test_pix2pix_gpu.py::test_forward_backward_against_oracle_port:
cosine between this repo's parameter gradients and the CPU oracle's, per layer, over several runs (split-K / BatchNorm
atomics make the bf16 path non-deterministic).  usage: python grad_noise.py [batch] [runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "thesis-pai-reconstruction_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import pix2pix_port as port
from test_pix2pix_gpu import _build

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = _build("ssim+psnr", seed=3)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
x, target = port.synthetic_pairs(n, seed=50 + n)
tr = port.OracleTrainer(sd, "ssim+psnr")
lo = port.generator_loss(tr.sd, "ssim+psnr", x, port.unet_forward(tr.sd, x, training=True), target)
lo.backward()
named = dict(m.named_parameters())
cos = {}
for r in range(runs):
    m.load_state_dict(sd)
    m.train()
    m.zero_grad(set_to_none=True)
    y = m(x.cuda())
    m.loss(x.cuda(), y, target.cuda()).backward()
    for k in tr.g_keys:
        go, g = tr.sd[k].grad.double(), named[k].grad.cpu().double()
        if float(go.norm()) < 1e-4:
            continue
        cos.setdefault(k, []).append(float((g * go).sum() / (g.norm() * go.norm() + 1e-30)))
for k, v in cos.items():
    if k.endswith("weight") and ("code.1." in k or "encoders.0" in k or "decoders.7" in k):
        print(f"{k:40s} min {min(v):.4f} max {max(v):.4f}")
print("overall min", min(min(v) for v in cos.values()), [k for k, v in cos.items() if min(v) < 0.93])
