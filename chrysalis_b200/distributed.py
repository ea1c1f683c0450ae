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


def allgather_ragged(t: torch.Tensor, n_total: int | None = None, return_lens: bool = False):
    """Concatenate the 1-D / 2-D (rows) shards of every rank in rank order.

    Shards may differ in length (a gene count that does not divide by the world size, or a gene
    filter that keeps different numbers per shard): lengths are exchanged first, shards are padded
    to the longest one for the fixed-size all-gather."""
    rank, size = world()
    if size == 1:
        return (t, [int(t.shape[0])]) if return_lens else t
    if _host_backend(t):
        out = allgather_ragged(t.cpu(), n_total, return_lens)
        return (out[0].to(t.device), out[1]) if return_lens else out.to(t.device)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    lens_t = torch.zeros(size, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(lens_t, n_local)
    lens = [int(x) for x in lens_t.cpu()]                    # one read-back for all ranks' lengths
    m = max(lens)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:n] for p, n in zip(parts, lens)], dim=0)
    if n_total is not None and out.shape[0] != n_total:
        raise RuntimeError(f"all-gathered {out.shape[0]} items, expected {n_total}")
    return (out, lens) if return_lens else out


def allgather_objects(obj) -> list:
    """One picklable host object per rank, in rank order (gene names of a gene-sharded AnnData)."""
    rank, size = world()
    if size == 1:
        return [obj]
    out = [None] * size
    dist.all_gather_object(out, obj)
    return out


class PeerGather:
    """All-gather of equally sized float64 shards through peer memory (``chr_p2p_allgather_f64``): every rank writes its
    shard into every peer's receive buffer over NVLink and the kernel itself waits for the peers' flags - one launch whose
    latency does not grow with the number of ranks (a ring all-gather of 18 KB pays one hop per rank).  Needs the NCCL backend
    with one GPU per rank and torch's symmetric-memory allocator; ``PeerGather.create`` returns None otherwise and the caller
    keeps the NCCL collective.  Two receive buffers alternate, so a rank that is one call ahead never overwrites data a peer
    still reads."""

    SIGNAL_SLOT0 = 384          # uint32 slots of the signal pad used here (torch's own barriers use the first channels)

    def __init__(self, handle, buf, max_elems):
        self.h, self.buf, self.max_elems, self.epoch = handle, buf, int(max_elems), 0
        self.rank, self.world = handle.rank, handle.world_size

    @staticmethod
    def create(max_elems: int):
        rank, size = world()
        if size == 1 or not torch.cuda.is_available() or dist.get_backend() != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm

            group = dist.group.WORLD
            buf = symm.empty(2 * size * int(max_elems), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
            handle = symm.rendezvous(buf, group)
            if handle.signal_pad_size < 4 * (PeerGather.SIGNAL_SLOT0 + 2 * size):
                return None
            buf.zero_()
            handle.barrier()
            return PeerGather(handle, buf, max_elems)
        except Exception:  # noqa: BLE001 - any failure of the optional fast path keeps the NCCL collective
            return None

    def gather(self, shard: torch.Tensor, count: int | None = None) -> torch.Tensor:
        """Returns a [world, max_elems] view of the receive buffer: row p holds rank p's shard (first ``count`` elements)."""
        from . import _lib

        count = shard.numel() if count is None else int(count)
        assert shard.dtype == torch.float64 and shard.is_contiguous() and count <= self.max_elems
        self.epoch += 1
        half = self.epoch & 1
        base = half * self.world * self.max_elems
        _lib.call("chr_p2p_allgather_f64", shard.data_ptr(), count, self.h.buffer_ptrs_dev, self.h.signal_pad_ptrs_dev, self.rank, self.world,
                  base + self.rank * self.max_elems, PeerGather.SIGNAL_SLOT0 + half * self.world, self.epoch,
                  torch.cuda.current_stream().cuda_stream)
        return self.buf[base:base + self.world * self.max_elems].view(self.world, self.max_elems)
