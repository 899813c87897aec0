"""World-size-2 gloo test of the data-parallel exchange (pai_b200/dp.py): gradient averaging after
manual_backward and the initial parameter broadcast -- the N>1 host logic of bench.py, on CPU."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pai_b200 import dp
    r, _, w = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dp.world_size() == world
    torch.manual_seed(100 + rank)                      # different initial weights per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))
    dp.broadcast_parameters(net, src=0)
    ref = [p.detach().clone() for p in net.parameters()]
    x = torch.full((6, 8), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    n = dp.allreduce_gradients(net.parameters())
    dp.barrier()
    mx = dp.allreduce_max(float(rank), torch.device("cpu"))
    # overlapped exchange used by the fused backward nodes: a large gradient is averaged while "backward" continues,
    # and allreduce_gradients() afterwards must not average it a second time
    os.environ["PAI_DP_OVERLAP"] = "1"              # opt-in (see dp.overlap_enabled)
    big = torch.nn.Parameter(torch.zeros(1 << 18))
    small = torch.nn.Parameter(torch.zeros(5))
    big.grad = torch.full((1 << 18,), float(rank + 1))
    small.grad = torch.full((5,), float(10 * (rank + 1)))
    dp.allreduce_async(big.grad)
    n2 = dp.allreduce_gradients([big, small])
    assert float(big.grad[0]) == 1.5 and float(big.grad[-1]) == 1.5, big.grad[:3]
    assert float(small.grad[0]) == 15.0 and n2 == 5
    os.environ["PAI_DP_OVERLAP"] = "0"
    big.grad = torch.full((1 << 18,), float(rank + 1))
    dp.allreduce_async(big.grad)                    # switched off: a no-op, the grouped all-reduce averages it
    n3 = dp.allreduce_gradients([big])
    assert float(big.grad[7]) == 1.5 and n3 == 1 << 18
    q.put((rank, [t.numpy() for t in ref], [g.numpy() for g in local], [p.grad.numpy() for p in net.parameters()], n, mx))
    torch.distributed.destroy_process_group()


def test_gradient_allreduce_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, ref0, loc0, avg0, n0, mx0), (_, ref1, loc1, avg1, n1, mx1) = out
    for a, b in zip(ref0, ref1):                       # broadcast: both ranks start from rank 0's weights
        assert (a == b).all()
    for l0, l1, a0, a1 in zip(loc0, loc1, avg0, avg1):
        want = (l0 + l1) / 2
        assert abs(a0 - want).max() < 1e-6 and abs(a1 - want).max() < 1e-6
    assert n0 == n1 == sum(a.size for a in avg0) and mx0 == mx1 == 1.0


def test_single_process_is_a_no_op():
    from pai_b200 import dp
    lin = torch.nn.Linear(3, 3)
    lin(torch.ones(2, 3)).sum().backward()
    g = lin.weight.grad.clone()
    assert dp.allreduce_gradients(lin.parameters()) == 0 and torch.equal(g, lin.weight.grad)
    assert dp.world_size() == 1


def _sweep_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pai_b200 import dp, metrics
    dp.init_from_env(backend="gloo")
    g = torch.Generator().manual_seed(5)
    n = 11                                              # ragged: shards of 6 and 5 pairs
    ssim_all, sse_all, bands_all = torch.rand(n, generator=g), torch.rand(n, generator=g) * 50, torch.rand(n, 16, generator=g)
    lo, hi = metrics.shard_bounds(n, rank, world)
    s, e, b = metrics.gather_stats(ssim_all[lo:hi], sse_all[lo:hi], bands_all[lo:hi])
    got = metrics.finalize_report(s, e, b, 65536)
    want = metrics.finalize_report(ssim_all, sse_all, bands_all, 65536)
    ok = all(torch.equal(got[k], want[k]) for k in ("ssim", "psnr", "mse", "ssim_mean", "psnr_mean", "rmse", "depth_ssim"))
    s2, e2, b2 = metrics.gather_stats(ssim_all[lo:hi], sse_all[lo:hi], None)       # sweep without depth bands
    ok = ok and b2 is None and torch.equal(s2, ssim_all) and torch.equal(e2, sse_all)
    q.put((rank, ok, (lo, hi)))
    torch.distributed.destroy_process_group()


def test_sharded_evaluation_sweep_world2():
    """SURVEY.md 8(e), evaluation sweep: contiguous shards of the pairs per rank, all-gather of the per-image
    vectors, then exactly the single-process reductions (mean / unbiased std over images, global RMSE)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] and out[1][1]
    assert out[0][2] == (0, 6) and out[1][2] == (6, 11)


def test_shard_bounds_cover_everything():
    from pai_b200 import metrics
    for n in (0, 1, 7, 10000):
        for world in (1, 2, 3, 8):
            cuts = [metrics.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
