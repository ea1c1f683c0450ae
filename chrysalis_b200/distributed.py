"""Multi-GPU plumbing of the path: one process per GPU, ``torch.distributed``.

Moran's I shards by GENES - independent units, no data-path collective.  The only exchanges are
  * an all-reduce of the per-spot totals (``normalize_total`` sums over every kept gene, and the
    kept genes are spread over the ranks), and
  * an all-gather of the per-gene statistics (8 bytes per gene).
The PCA Gram shards by SPOTS with an all-reduce of the partial Grams and column sums.
Everything here works on tensors of any device, so the same code runs under NCCL on the GPUs and
under gloo in the CPU tests.  With no initialised process group every helper is the identity.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int | None = None, size: int | None = None):
    """Contiguous, balanced range [lo, hi) of ``n_items`` for ``rank`` (first ranks get the remainder)."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    base, rem = divmod(int(n_items), size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _host_backend(t: torch.Tensor) -> bool:
    """gloo exchanges device tensors through the host only for some collectives (no all_gather): stage them ourselves."""
    return t.is_cuda and dist.get_backend() == "gloo"


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (per-spot totals, partial Grams, column sums)."""
    if world()[1] > 1:
        if _host_backend(t):
            h = t.cpu()
            dist.all_reduce(h, op=dist.ReduceOp.SUM)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather_ragged(t: torch.Tensor, n_total: int | None = None) -> torch.Tensor:
    """Concatenate the 1-D / 2-D (rows) shards of every rank in rank order.

    Shards may differ in length (a gene count that does not divide by the world size, or a gene
    filter that keeps different numbers per shard): lengths are exchanged first, shards are padded
    to the longest one for the fixed-size all-gather."""
    rank, size = world()
    if size == 1:
        return t
    if _host_backend(t):
        return allgather_ragged(t.cpu(), n_total).to(t.device)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    lens = [torch.zeros_like(n_local) for _ in range(size)]
    dist.all_gather(lens, n_local)
    lens = [int(x.item()) for x in lens]
    m = max(lens)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:n] for p, n in zip(parts, lens)], dim=0)
    if n_total is not None and out.shape[0] != n_total:
        raise RuntimeError(f"all-gathered {out.shape[0]} items, expected {n_total}")
    return out


def allgather_objects(obj) -> list:
    """One picklable host object per rank, in rank order (gene names of a gene-sharded AnnData)."""
    rank, size = world()
    if size == 1:
        return [obj]
    out = [None] * size
    dist.all_gather_object(out, obj)
    return out
