import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
from pai_b200 import metrics, lib
lib.load()
dev = torch.device("cuda")
n = 2000
g = torch.Generator(device=dev).manual_seed(7)
base = torch.rand(n, 1, 256, 256, device=dev, generator=g)
pred = (base + 0.05 * torch.randn(n, 1, 256, 256, device=dev, generator=g)).clamp_(0, 1)
for _ in range(3):
    metrics._launch_fwd(pred, base, False, True, False)
torch.cuda.synchronize()
