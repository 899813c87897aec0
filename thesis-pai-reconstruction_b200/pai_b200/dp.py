"""Batch data parallelism over the GPUs of one box: one process per GPU, replicated parameters and
optimiser state, per-replica BatchNorm statistics, gradient averaging with NCCL over NVLink/NVSwitch.

The reference has no distributed code of its own; ``pl.Trainer`` (main.py:123-135) would run DDP over the
visible GPUs without ``sync_batchnorm`` -- this module is that exchange for the manual-optimisation
``training_step`` (models/wrapper.py:135,159 are the two ``manual_backward`` call sites)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

_enabled = False


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) -> (rank, local_rank, world)."""
    global _enabled
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        # collectives captured inside the training-step CUDA graph (pai_b200.graph): the NCCL watchdog must not query
        # events of a capturing stream
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            if overlap_enabled():
                # The all-reduce of layer k's weight gradient runs while the backward of layers k-1, k-2, ... continues.
                # Our tensor-core kernels are persistent with one 195 KB-shared-memory CTA per SM: a NCCL CTA cannot share
                # an SM with one, so without SMs of its own either side waits for the other for a whole kernel.  NCCL is
                # held to RESERVED_SMS CTAs and the library launches (#SMs - RESERVED_SMS) persistent CTAs.
                # (the reservation itself is switched on by the first all-reduce of a backward pass and off again by
                # finish_async(): forward passes keep every SM)
                os.environ.setdefault("NCCL_MAX_CTAS", str(RESERVED_SMS))
                os.environ.setdefault("NCCL_MIN_CTAS", str(min(4, RESERVED_SMS)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    _enabled = world > 1
    return rank, local, world


RESERVED_SMS = int(os.environ.get("PAI_DP_RESERVED_SMS", "8"))


def overlap_enabled() -> bool:
    """PAI_DP_OVERLAP=1: gradient all-reduces start inside the backward pass, on SMs reserved for NCCL.  Off by default:
    measured on 2 B200s (round 2, batch 64) the overlapped step takes 9.07 ms against 8.97 ms for the single grouped
    all-reduce after ``manual_backward`` (1 GPU: 8.35 ms) -- 20 separate collectives on 8 CTAs plus 5 % fewer SMs for the
    backward GEMMs cost more than the 0.6 ms they hide."""
    return os.environ.get("PAI_DP_OVERLAP", "0") == "1"


def world_size() -> int:
    return dist.get_world_size() if (_enabled and dist.is_initialized()) else 1


_BIG = 1 << 18      # elements: gradients at least this large are reduced in place, one collective each

_flat = {}


def grad_exchange_dtype():
    """Wire format of the large weight gradients: bf16 (default) or fp32 (PAI_DP_GRAD_DTYPE=fp32: bit-for-bit the fp32
    average Lightning DDP would compute for main.py:123-135, at twice the bytes)."""
    return torch.float32 if os.environ.get("PAI_DP_GRAD_DTYPE", "bf16").lower() in ("fp32", "float32") else torch.bfloat16


def _flat_buffer(numel: int, device) -> torch.Tensor:
    buf = _flat.get((numel, device))
    if buf is None:
        buf = _flat[(numel, device)] = torch.empty(numel, dtype=torch.bfloat16, device=device)
    return buf

_pending = []       # (work, tensor) of the all-reduces started while the backward pass is still running
_done = set()       # data_ptr of the gradients those collectives already average


def active() -> bool:
    return _enabled and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_async(t: torch.Tensor) -> None:
    """Called by the fused backward nodes (pai_b200.engine) the moment a large weight gradient exists: starts its
    in-place average over all ranks on NCCL's stream, so the exchange of layer k overlaps the dgrad / wgrad GEMMs of
    layers k-1, k-2, ... (the bucketed overlap DDP would give main.py:123-135).  ``finish_async`` joins them.

    Opt-in (PAI_DP_OVERLAP=1, see ``overlap_enabled``).  The persistent kernels leave RESERVED_SMS SMs to NCCL between the
    first of these calls and ``finish_async`` (``pai_reserve_sms``), so the collectives and the GEMMs run side by side."""
    if not active() or t.numel() < _BIG or not t.is_contiguous() or not overlap_enabled():
        return
    nccl = dist.get_backend() == "nccl"
    if nccl and not _pending:
        _reserve(RESERVED_SMS)         # from here to finish_async() the persistent kernels leave SMs to NCCL
    work = dist.all_reduce(t, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, async_op=True)
    _pending.append((work, t))
    _done.add(t.data_ptr())


_reserved = 0


def _reserve(n: int) -> None:
    global _reserved
    if n != _reserved:
        from . import lib
        lib.call("pai_reserve_sms", n, kernels=0)
        _reserved = n


def finish_async() -> None:
    """Makes the current stream wait for every collective started by ``allreduce_async``."""
    if not _pending:
        return
    nccl = dist.get_backend() == "nccl"
    world = dist.get_world_size()
    for work, t in _pending:
        work.wait()
        if not nccl:
            t.div_(world)
    _pending.clear()
    _reserve(0)


def allreduce_gradients(params) -> int:
    """In-place average of ``p.grad`` over all ranks (gradients already averaged during the backward pass by
    ``allreduce_async`` are skipped).  Large gradients (the convolution weights) are averaged
    in place by their own asynchronous all-reduce -- no flatten / copy-back passes over the 218 MB of generator
    gradients --, the many small ones (biases, BatchNorm affine) share one flat buffer per dtype.  Returns the
    number of elements exchanged (0 when not data-parallel)."""
    if not (_enabled and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    finish_async()
    grads = [p.grad for p in params if p.grad is not None and p.grad.data_ptr() not in _done]
    _done.clear()
    if not grads:
        return 0
    world = dist.get_world_size()
    avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else None
    total = 0
    big = [g for g in grads if g.numel() >= _BIG and g.is_contiguous()]
    small = [g for g in grads if not (g.numel() >= _BIG and g.is_contiguous())]
    # the large gradients (convolution weights, 229 MB fp32 for generator + PatchGAN)
    if big and avg is not None and grad_exchange_dtype() == torch.bfloat16 and all(g.dtype == torch.float32 for g in grads):
        # travel as bf16 (115 MB instead of 229 MB; at 8 GPUs the all-reduce is bandwidth bound:
        # 9.12 vs 9.35 ms/step): one multi-tensor cast into a flat buffer, ONE collective for large and small gradients
        # alike, one multi-tensor cast back.  Every element is rounded to bf16 once before and once after the sum (2^-9
        # relative); all ranks receive the same bits, so replicas stay identical.
        every = big + small
        n_all = sum(g.numel() for g in every)
        flat = _flat_buffer(n_all, every[0].device)
        views, off = [], 0
        for g in every:
            views.append(flat[off:off + g.numel()].view(g.shape))
            off += g.numel()
        torch._foreach_copy_(views, [g.contiguous() for g in every])
        dist.all_reduce(flat, op=avg)
        torch._foreach_copy_(every, views)
        return n_all
    elif big:
        # averaged in place by ONE grouped NCCL launch (ncclGroupStart/End around the per-tensor all-reduces) -- no
        # flatten / copy-back passes and no per-tensor launch latency
        op = avg if avg is not None else dist.ReduceOp.SUM
        if avg is not None and hasattr(dist, "_coalescing_manager"):
            with dist._coalescing_manager(device=big[0].device, async_ops=False):
                for g in big:
                    dist.all_reduce(g, op=op)
        else:
            works = [dist.all_reduce(g, op=op, async_op=True) for g in big]
            for w in works:
                w.wait()
            if avg is None:
                for g in big:
                    g.div_(world)
        total += sum(g.numel() for g in big)
    for dtype in {g.dtype for g in small}:
        group = [g for g in small if g.dtype == dtype]
        flat = torch.cat([g.reshape(-1) for g in group])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in group:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        total += flat.numel()
    return total


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Makes every replica start from rank ``src``'s parameters and buffers."""
    if not (_enabled and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def barrier() -> None:
    if _enabled and dist.is_initialized():
        dist.barrier()


def allreduce_max(value: float, device) -> float:
    if not (_enabled and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
