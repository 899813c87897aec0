"""ORACLE (test infrastructure): how far is the reference's own fp32 arithmetic from the exact (fp64) gradient of one
GAN generator-loss evaluation at small batch?  (Deep layers: 0.5-1.1e-2 at batch 2, 5e-3 at 6, 2e-3 at 16 -- train-mode
BatchNorm over N*4 values.)  This is the floor tests/test_check_path_gpu.py holds the fp32 check path to.
    python oracle/grad_conditioning.py [batch]"""
import sys
sys.path.insert(0,'oracle')
import torch, torch.nn.functional as F, pix2pix_port as port
from bf16_sites import fwd
torch.set_num_threads(8)
sd = port.init_state(11, loss_type="gan", disc_in_channels=1)
x, t = port.synthetic_pairs(int(sys.argv[1]) if len(sys.argv) > 1 else 2, seed=92)
def grads(dtype):
    s = {k: (v.to(dtype).requires_grad_(port._is_param(k)) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    xx, tt = x.to(dtype), t.to(dtype)
    y = fwd(s, xx, set())
    lab = port.disc_forward(s, xx, y)
    l = F.binary_cross_entropy_with_logits(lab, torch.ones_like(lab)) + 50 * F.l1_loss(y, tt)
    l.backward()
    return {k: v.grad.double() for k, v in s.items() if k.startswith("unet.") and v.is_floating_point() and v.grad is not None}, float(l.detach())
g32, l32 = grads(torch.float32)
g64, l64 = grads(torch.float64)
print('loss', l32, l64)
worst = []
for k in g32:
    n = float(g64[k].norm())
    if n < 1e-5: continue
    worst.append((float((g32[k]-g64[k]).norm())/n, k))
for r, k in sorted(worst, reverse=True)[:8]: print(f"{r:.3e} {k}")
