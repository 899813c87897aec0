import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch, bench
from pai_b200 import lib
lib.load()
dev = torch.device("cuda")
m = bench.build_model().to(dev).train()
data = [tuple(t.to(dev) for t in bench.synthetic_pairs(64, seed=i)) for i in range(2)]
for i in range(3):
    m.training_step(data[i % 2], i); m.logged.clear()
torch.cuda.synchronize()
