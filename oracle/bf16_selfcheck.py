"""Reference-vs-itself: fp32 vs bf16-autocast gradients of the oracle port on CPU (how much of the
deep-layer gradient disagreement is inherent to bf16 with batch-statistics BatchNorm)."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import pix2pix_port as port
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = port.init_state(3, 1, 1, loss_type="ssim+psnr")
x, target = port.synthetic_pairs(n, seed=50 + n)
def run(autocast):
    tr = port.OracleTrainer(sd, "ssim+psnr")
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        y = port.unet_forward(tr.sd, x, training=True)
    l = port.generator_loss(tr.sd, "ssim+psnr", x, y.float(), target)
    l.backward()
    return tr, y.float().detach()
a, ya = run(False); b, yb = run(True)
print("ydiff", (ya-yb).abs().max().item())
for k in a.g_keys:
    if not k.endswith("weight") or "code.2." in k: continue   # BatchNorm affine
    g, go = b.sd[k].grad.double(), a.sd[k].grad.double()
    print(f"{k:40s} cos {float((g*go).sum()/(g.norm()*go.norm())):.4f} ratio {float(g.norm()/go.norm()):.4f}")
