"""Batch sharding over the GPUs of one box: latents are independent, so rank r of W owns a
contiguous range of the batch and nothing but the final bit-match counters is exchanged
(one all-reduce of the GSWM_CTR_* int64 vector).  The uniform source is keyed by GLOBAL latent index, so the
latents a rank produces do not depend on W.

On GPUs the exchange is libgswm's own: a ``gswm.Comm`` (per-rank mailboxes mapped over NVLink, include/gswm.h) whose
all-reduce is one small kernel -- or no extra kernel at all when it is fused into the last extract launch
(``extract_batch(..., comm=comm)``).  ``torch.distributed`` carries the one-time handle exchange, and the reduction
itself only where there is no CUDA device (the gloo tests on CPU)."""
from __future__ import annotations


def shard_range(n_latents: int, rank: int, world: int):
    """[lo, hi) of the batch owned by `rank`: contiguous, sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_latents, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_counters(counters, group=None, comm=None):
    """Sum the GSWM_CTR_* counter tensor (int64, the buffer gswm_extract accumulated into) over ranks, in place, on the
    current stream: through ``comm`` (gswm.Comm) when given, else through torch.distributed (gloo in the CPU tests)."""
    if comm is not None:
        return comm.allreduce_counters(counters)
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters
