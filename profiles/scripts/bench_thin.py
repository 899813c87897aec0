"""Microbenchmark of the single-pass thin-layer kernels (csrc/thin.cu) at the BASELINE batch (64 x 256 x 256):
CUDA-event time per launch and achieved HBM GB/s on the algorithmic bytes; L2 is flushed between launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
from pai_b200 import engine, ops

dev = torch.device("cuda")
N = int(os.environ.get("N", 64))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    if os.environ.get("ONCE"):          # one launch per kernel (for an ncu capture of this script)
        fn(); torch.cuda.synchronize()
        return 1.0
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


def report(name, us, nbytes):
    print(f"{name:44s} {us:8.1f} us   {nbytes / us / 1e3:8.1f} GB/s   ({nbytes / 1e6:.0f} MB algorithmic)", flush=True)


g = torch.Generator(device="cuda").manual_seed(0)
plane = torch.randn(N, 256, 256, device=dev, generator=g)
plane2 = torch.randn(N, 256, 256, device=dev, generator=g)
w1 = torch.randn(64, 1, 4, 4, device=dev) * 0.1
w2 = torch.randn(64, 2, 4, 4, device=dev) * 0.1
wT = torch.randn(128, 1, 4, 4, device=dev) * 0.1
b64 = torch.randn(64, device=dev)
pk1 = engine._pad_cols(w1.permute(0, 2, 3, 1).reshape(64, -1))
pk2 = engine._pad_cols(w2.permute(0, 2, 3, 1).reshape(64, -1))
pkT = engine._pad_cols(wT[:, 0].reshape(128, 16))
a1 = torch.empty(N, 128, 128, 64, dtype=torch.bfloat16, device=dev)
cat = torch.randn(N, 128, 128, 128, device=dev).bfloat16()
pb = plane.numel() * 4
report("enc0 fprop (1 plane -> 64ch x2 outputs)", timeit(lambda: ops.thin_conv_fprop([plane], pk1, 64, b64, a1, 1, cat[..., 64:], 0)), pb + 2 * a1.numel() * 2)
report("D0 fprop (2 planes -> 64ch)", timeit(lambda: ops.thin_conv_fprop([plane, plane2], pk2, 64, b64, a1, 1)), 2 * pb + a1.numel() * 2)
dcat = torch.empty(N, 128, 128, 128, dtype=torch.bfloat16, device=dev)
report("dec7 dgrad (1 plane -> 128ch)", timeit(lambda: ops.thin_conv_fprop([plane], pkT, 128, None, dcat, 0)), pb + dcat.numel() * 2)
report("enc0 wgrad (64ch x 1 plane)", timeit(lambda: ops.thin_conv_wgrad(a1, [plane])), pb + a1.numel() * 2)
report("D0 wgrad (64ch x 2 planes)", timeit(lambda: ops.thin_conv_wgrad(a1, [plane, plane2])), 2 * pb + a1.numel() * 2)
report("dec7 wgrad (128ch x 1 plane)", timeit(lambda: ops.thin_conv_wgrad(cat, [plane])), pb + cat.numel() * 2)
taps128 = wT[:, 0].reshape(128, 16).t().contiguous().bfloat16()
taps64 = w2[:, 1].reshape(64, 16).t().contiguous().bfloat16()
bias1 = torch.zeros(1, device=dev)
report("dec7 fprop (128ch -> plane, tanh)", timeit(lambda: ops.thin_convT_plane(cat, taps128, bias1, 3)), pb + cat.numel() * 2)
report("D0 dgrad (64ch -> plane)", timeit(lambda: ops.thin_convT_plane(a1, taps64)), pb + a1.numel() * 2)
h = torch.randn(N, 16, 16, 512, device=dev).bfloat16()
hp = torch.randn(16, 512, device=dev).bfloat16()
report("head: pointwise gemm + gather", timeit(lambda: ops.col2im4x4s1(ops.pointwise_gemm(h, hp, 16, out_f32=True))), h.numel() * 2)
