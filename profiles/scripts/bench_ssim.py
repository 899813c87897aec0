import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from pai_b200 import metrics, lib
lib.load()
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
g = torch.Generator(device=dev).manual_seed(7)
base = torch.rand(n, 1, 256, 256, device=dev, generator=g)
pred = (base + 0.05 * torch.randn(n, 1, 256, 256, device=dev, generator=g)).clamp_(0, 1)

def run(tag):
    for _ in range(3):
        out = metrics._launch_fwd(pred, base, False, True, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = metrics._launch_fwd(pred, base, False, True, False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag}: {ms:.3f} ms  {n*524288/ms/1e6:.1f} GB/s  frac {n*524288/ms/1e6/6525.2:.3f}", flush=True)
    return out

os.environ["PAI_SSIM_NO_STREAM"] = "1"
a = run("rows  ")
del os.environ["PAI_SSIM_NO_STREAM"]
b = run("stream")
for name, u, v in zip(("ssim_sum", "sse", "bands"), a, b):
    d = (u - v).abs().max().item()
    print(name, "max abs diff", d, "max", u.abs().max().item(), "bit-equal frac", (u == v).float().mean().item())
# oracle on a few images
import torchmetrics_port as tm
idx = [0, 1, 147, 148, 149, n - 1]
want = tm.structural_similarity_index_measure(pred[idx].cpu(), base[idx].cpu(), data_range=1.0, reduction="none")
got = (b[0][idx] / (246 * 246)).cpu()
print("oracle diff", (want - got).abs().max().item())
