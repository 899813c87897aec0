import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
import bench
from pai_b200 import lib
lib.load()
dev = torch.device("cuda")
which = sys.argv[1]
from models.attention_unet import AttentionUnetGAN
from models.res_unet import ResUnetGAN
from models.trans_unet import TransUnetGAN
torch.manual_seed(0)
if which == "res":
    m, batch = ResUnetGAN(in_channels=1, out_channels=1, res_type="next", dropout=0.0, loss_type="ssim"), 32
elif which == "att":
    m, batch = AttentionUnetGAN(in_channels=1, out_channels=1, dropout=0.0, loss_type="ssim"), 64
else:
    m, batch = TransUnetGAN(in_channels=1, out_channels=1, channel_mults=(1, 2, 2, 4, 4), patch_size=4, dropout=0.0, loss_type="ssim"), 64
m = m.to(dev).train()
x, t = bench.synthetic_pairs(batch, seed=4242); x, t = x.to(dev), t.to(dev)
for i in range(3):
    m.training_step((x, t), i); m.logged.clear()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        m.training_step((x, t), i); m.logged.clear()
    torch.cuda.synchronize(); t1 = time.perf_counter()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev))
busy = sum(b - a for a, b, _ in ks); span = ks[-1][1] - ks[0][0]
print(f"{which}: wall {1e3*(t1-t0)/steps:.3f} ms/step; kernels {len(ks)/steps:.0f}/step; busy {busy/steps/1e3:.3f}; span {span/steps/1e3:.3f} ms/step")
agg = collections.Counter(); cnt = collections.Counter()
for a, b, n in ks: agg[n[:100]] += (b - a); cnt[n[:100]] += 1
for n, tt in agg.most_common(22): print(f"{tt/steps/1e3:8.3f} ms/step  x{cnt[n]/steps:6.1f}  {n}")
