"""Microbenchmark of the HBM-bound passes around the GEMMs (csrc/elementwise.cu, optim.cu) at the shapes of the batch-64
step: CUDA-event time per launch, achieved GB/s on the algorithmic bytes, L2 flushed between launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "thesis-pai-reconstruction_b200"))
import torch
from pai_b200 import ops

dev = torch.device("cuda")
N = int(os.environ.get("N", 64))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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


def report(name, us, nbytes):
    print(f"{name:52s} {us:8.1f} us   {nbytes / us / 1e3:8.1f} GB/s   ({nbytes / 1e6:.0f} MB algorithmic)", flush=True)


def bf(*shape):
    return torch.randn(*shape, device=dev).bfloat16()


for name, hw, c in (("enc1", 64, 128), ("enc2", 32, 256), ("enc3", 16, 512), ("dec6", 128, 64)):
    x = bf(N, hw, hw, c)
    e = x.numel()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sums = ops.bn_stats(x)
    report(f"{name} bn_stats [{N}x{hw}x{hw}x{c}]", timeit(lambda: ops.bn_stats(x)), 2 * e)
    ss = ops.bn_finalize(sums, x.numel() // c, c, gamma, beta, rm, rv)
    o1 = torch.empty_like(x)
    cat = torch.empty(N, hw, hw, 2 * c, dtype=torch.bfloat16, device=dev)
    report(f"{name} bn_apply_act -> 1 output", timeit(lambda: ops.bn_apply_act(x, ss, o1, 2)), 4 * e)
    report(f"{name} bn_apply_act -> 2 outputs (2nd in concat)", timeit(lambda: ops.bn_apply_act(x, ss, o1, 1, cat[..., c:], 2)), 6 * e)
    g1 = bf(N, hw, hw, c)
    gcat = bf(N, hw, hw, 2 * c)
    report(f"{name} bn_bwd_reduce (1 gradient)", timeit(lambda: ops.bn_bwd_reduce(x, ss, g1, 2)), 4 * e)
    report(f"{name} bn_bwd_reduce (2 gradients, 2nd from concat)", timeit(lambda: ops.bn_bwd_reduce(x, ss, g1, 1, gcat[..., c:], 2)), 6 * e)
    s2 = ops.bn_bwd_reduce(x, ss, g1, 2)
    dx = torch.empty_like(x)
    report(f"{name} bn_bwd_apply (1 gradient)", timeit(lambda: ops.bn_bwd_apply(x, ss, g1, 2, None, 0, s2, gamma, dx)), 6 * e)
    report(f"{name} bn_bwd_apply (2 gradients)", timeit(lambda: ops.bn_bwd_apply(x, ss, g1, 1, gcat[..., c:], 2, s2, gamma, dx)), 8 * e)
    del x, o1, cat, g1, gcat, dx

x = bf(N, 128, 128, 64); g1 = bf(N, 128, 128, 64); gcat = bf(N, 128, 128, 128); dx = torch.empty_like(x)
report("enc0 act_bwd (2 gradients) + bias sums", timeit(lambda: ops.act_bwd(x, g1, 1, gcat[..., 64:], 2, dx)), 8 * x.numel())
dw = torch.randn(16, 512, 1024, device=dev)
report("wgrad_finish [16,512,1024] fp32", timeit(lambda: ops.wgrad_finish(dw)), 2 * dw.numel() * 4)
report("torch zero fill 32 MB", timeit(lambda: dw.zero_()), dw.numel() * 4)
