import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
import bench
from pai_b200 import lib
lib.load()
dev = torch.device("cuda")
m = bench.build_model().to(dev).train()
data = [tuple(t.to(dev) for t in bench.synthetic_pairs(64, seed=i)) for i in range(4)]
for i in range(5):
    m.training_step(data[i % 4], i); m.logged.clear()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
steps = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        m.training_step(data[i % 4], i); m.logged.clear()
    torch.cuda.synchronize(); t1 = time.perf_counter()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev))
busy = sum(b - a for a, b, _ in ks)
span = ks[-1][1] - ks[0][0]
print(f"wall {1e3*(t1-t0)/steps:.3f} ms/step (under profiler); kernels {len(ks)/steps:.0f}/step; busy {busy/steps/1e3:.3f} ms/step; span {span/steps/1e3:.3f} ms/step; idle {(span-busy)/steps/1e3:.3f} ms/step")
# gaps histogram
gaps = [ks[i+1][0] - ks[i][1] for i in range(len(ks)-1)]
import collections
big = sorted(((g, ks[i][2][:50], ks[i+1][2][:50]) for i, g in enumerate(gaps) if g > 5), reverse=True)[:25]
for g, a, b in big: print(f"{g:8.1f} us  after {a}  before {b}")
print("gaps >2us:", sum(1 for g in gaps if g > 2)/steps, "per step, total", sum(g for g in gaps if g > 2)/steps/1e3, "ms/step")
agg = collections.Counter(); cnt = collections.Counter()
for a, b, n in ks: agg[n[:90]] += (b - a); cnt[n[:90]] += 1
for n, t in agg.most_common(60): print(f"{t/steps/1e3:8.3f} ms/step  x{cnt[n]/steps:6.1f}  {n}")
# per-kernel-name table (time per step, launches per step)
agg = collections.defaultdict(lambda: [0.0, 0])
for a, b, name in ks:
    agg[name][0] += b - a; agg[name][1] += 1
print("\nper kernel (us/step, launches/step):")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t/steps:9.1f} us  x{n/steps:5.1f}  {name[:150]}")
