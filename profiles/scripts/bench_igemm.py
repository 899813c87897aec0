"""Per-layer microbenchmark of the implicit-GEMM fprop kernel on the BASELINE generator / PatchGAN shapes at batch 64
(CUDA events, L2 flushed between launches).  ONCE=1: a single launch per layer for an ncu capture.
LAYERS=enc1,dec6 restricts the list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
from pai_b200 import ops

dev = torch.device("cuda")
N = int(os.environ.get("N", 64))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = set(filter(None, os.environ.get("LAYERS", "").split(",")))


def timeit(fn, reps=5):
    if os.environ.get("ONCE"):
        fn(); torch.cuda.synchronize()
        return 1.0
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).bfloat16()


# name, kind, hin, cin, cout
LAYERS = [("enc1", "conv", 128, 64, 128), ("enc2", "conv", 64, 128, 256), ("enc3", "conv", 32, 256, 512),
          ("enc4", "conv", 16, 512, 512), ("dec3", "convT", 8, 1024, 512), ("dec4", "convT", 16, 1024, 256),
          ("dec5", "convT", 32, 512, 128), ("dec6", "convT", 64, 256, 64),
          ("D1dgrad", "dgrad_act", 64, 128, 64), ("D1dgrad-ns", "dgrad_act_nosum", 64, 128, 64), ("D2dgrad", "dgrad_act", 32, 256, 128), ("D3dgrad", "dgrad_act", 16, 512, 256),
          ("dec4dgrad", "conv", 32, 256, 1024), ("dec5dgrad", "conv", 64, 128, 512), ("dec6dgrad", "conv", 128, 64, 256)]
for name, kind, hin, cin, cout in LAYERS:
    if only and name not in only:
        continue
    x = rnd(N, hin, hin, cin)
    bias = torch.randn(cout, device=dev)
    if kind == "conv":
        wp = ops.pack_conv_weight(torch.randn(cout, cin, 4, 4, device=dev) * 0.05)
        fl = 2.0 * N * (hin // 2) ** 2 * cout * 16 * cin
        fn = lambda: ops.conv4x4_fprop_bnstats(x, wp, cout, bias=bias)
    elif kind == "convT":
        wp = ops.pack_convT_weight(torch.randn(cin, cout, 4, 4, device=dev) * 0.05)
        fl = 2.0 * N * hin ** 2 * cout * 16 * cin
        fn = lambda: ops.convT4x4s2_fprop_bnstats(x, wp, cout, bias=bias)
    else:
        wp = ops.pack_convT_weight(torch.randn(cin, cout, 4, 4, device=dev) * 0.05)
        saved = rnd(N, 2 * hin, 2 * hin, cout)
        fl = 2.0 * N * hin ** 2 * cout * 16 * cin
        fn = (lambda: ops.conv4x4_dgrad_act(x, wp, cout, saved)) if kind == "dgrad_act" else (
            lambda: ops.conv4x4_dgrad_act(x, wp, cout, saved, want_colsum=False))
    us = timeit(fn)
    print(f"{name:10s} {kind:9s} {cin:5d}->{cout:4d} @{hin:3d}  {fl / 1e9:7.1f} GFLOP {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
    del x, wp
