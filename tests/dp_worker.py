"""2-rank NCCL worker of tests/test_baseline_configs_gpu.py (run under torch.distributed.run, one rank per GPU).

Checks on real hardware what the gloo world-2 CPU tests cannot: the gradient all-reduce CAPTURED inside the
training-step CUDA graph (pai_b200/graph.py + dp.py; call sites models/wrapper.py:135,159 of the reference)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "thesis-pai-reconstruction_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pix2pix_port as port  # noqa: E402  (synthetic data only)
from pai_b200 import dp  # noqa: E402


def build(seed):
    from models.pix2pix import Pix2Pix
    from models.utils import init_weights
    from models.wrapper import Discriminator
    torch.manual_seed(seed)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type="gan")
    m.discriminator = Discriminator(in_channels=1)
    m.discriminator.apply(init_weights)
    return m.cuda().train()


def same_on_all_ranks(t: torch.Tensor) -> bool:
    ref = t.detach().clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([1.0 if torch.equal(ref, t.detach()) else 0.0], device=t.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())


def main():
    rank, local, world = dp.init_from_env()
    assert world == 2 and dist.get_backend() == "nccl"
    dev = torch.device("cuda", local)
    # ---- (1) replicas stay bit-identical through graph-replayed steps on different data
    m = build(seed=100 + rank)                       # different weights per rank on purpose
    dp.broadcast_parameters(m, src=0)
    for k, v in m.state_dict().items():
        assert same_on_all_ranks(v), f"broadcast: {k} differs"
    m.enable_step_graph(warmup=2)
    data = [tuple(t.to(dev) for t in port.synthetic_pairs(4, seed=1000 * (rank + 1) + i)) for i in range(3)]
    for i in range(7):                               # 2 eager + 5 replays
        m.training_step(data[i % 3], i)
    torch.cuda.synchronize()
    assert m.__dict__["_pai_step_graph"].replays == 5
    bad = [k for k, v in m.named_parameters() if not same_on_all_ranks(v)]
    assert not bad, f"replicas diverged after graph replays: {bad[:5]}"
    # BatchNorm running statistics are per replica (un-synced DDP): they MUST differ (different data)
    rm = m.state_dict()["unet.encoders.1.encode.2.running_mean"]
    assert not same_on_all_ranks(rm), "running statistics unexpectedly identical: is every rank seeing the same data?"
    # ---- (2) averaged gradients of replicated data == single-rank gradients
    m2 = build(seed=7)
    dp.broadcast_parameters(m2, src=0)
    x, t = (a.to(dev) for a in port.synthetic_pairs(4, seed=4242))       # the same batch on both ranks
    opt_g, opt_d = m2.optimizers()

    def g_backward(reduce):
        m2.toggle_optimizer(opt_g)
        m2.unet.zero_grad(set_to_none=True)
        # identical BatchNorm running stats are irrelevant in train mode; the forward is deterministic up to atomics
        loss = m2.loss(x, m2.unet(x), t)
        if reduce:
            m2.manual_backward(loss)                 # backward + NCCL average (models/wrapper.py:159)
        else:
            loss.backward()
        m2.untoggle_optimizer(opt_g)
        return {k: p.grad.detach().clone() for k, p in m2.unet.named_parameters()}

    # (a) the collective itself, exactly: local gradients of one backward, gathered from both ranks, against what
    # allreduce_gradients leaves in p.grad (NCCL average in fp32: equal to (g0 + g1) / 2 up to one rounding)
    local_g = g_backward(False)
    expect = {}
    for k, g in local_g.items():
        both = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(both, g)
        expect[k] = (both[0] + both[1]) / 2
    bf16_wire = dp.grad_exchange_dtype() == torch.bfloat16
    dp.allreduce_gradients(p for p in m2.unet.parameters())
    worst = 0.0
    for k, p in m2.unet.named_parameters():
        if bf16_wire:
            # the gradients travel as bf16: rounded once on the way in and once on the way out
            err = (p.grad - expect[k]).abs().max().item()
            assert err <= 2.0 ** -7 * expect[k].abs().max().item() + 1e-12, (k, err)
        else:
            assert torch.allclose(p.grad, expect[k], rtol=1e-5, atol=1e-8), f"all-reduced gradient {k} != mean of the local gradients"
        assert same_on_all_ranks(p.grad), f"averaged gradient {k} differs between ranks"
    # ... and bit-exact fp32 averaging on request
    os.environ["PAI_DP_GRAD_DTYPE"] = "fp32"
    for k, p in m2.unet.named_parameters():
        p.grad.copy_(local_g[k])
    dp.allreduce_gradients(p for p in m2.unet.parameters())
    for k, p in m2.unet.named_parameters():
        assert torch.allclose(p.grad, expect[k], rtol=1e-5, atol=1e-8), f"fp32 all-reduced gradient {k} != mean of the local gradients"
    del os.environ["PAI_DP_GRAD_DTYPE"]
    # (b) the path training uses: manual_backward = backward with the per-layer all-reduces started inside it
    # (dp.allreduce_async) + the grouped reduce of the rest.  Same data on both ranks: the result must be identical on
    # both ranks and equal a local backward up to the run-to-run noise of the bf16 network (two backward passes of the
    # same batch differ by up to ~10 % on the deep layers at batch 4: split-K atomics reorder sums, BatchNorm over N*4
    # values amplifies the flipped bf16 roundings)
    avg_g = g_backward(True)
    for k in local_g:
        # (every rank must issue the same collectives: nothing rank-local may decide whether same_on_all_ranks runs)
        assert same_on_all_ranks(avg_g[k]), f"averaged gradient {k} differs between ranks"
        a, b = local_g[k].double(), avg_g[k].double()
        if float(a.norm()) < 1e-6:
            continue
        rel = float((a - b).norm() / a.norm())
        worst = max(worst, rel)
        assert rel < 0.25, (k, rel)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"dp_worker ok (world {world}; worst averaged-vs-local gradient deviation {worst:.2e})", flush=True)
    # leave without tearing the communicator down: destroy_process_group() can block forever while captured CUDA graphs
    # still reference the NCCL communicator (seen on 2 B200s); the process is done
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
