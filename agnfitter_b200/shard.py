"""
Catalogue sharding across GPUs (one process per GPU).

The reference's only parallelism is `multiprocessing.Pool.map` over catalogue rows
(RUN_AGNfitter_multi.py:246-270): sources are independent, every worker builds its own dictionary
and writes its own files.  The multi-GPU form keeps that shape: each rank owns a contiguous block of
the z-sorted catalogue (sources that share a dictionary z-slice stay on one rank, so a slice is
built and held once), there is NO collective on the evaluation path, and the only exchange is one
gather of fixed-size per-source summaries to rank 0 at the end.

Everything here is backend-agnostic `torch.distributed` (nccl on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_sources, rank, world):
    """[lo, hi) of rank's block: sizes differ by at most one, blocks are contiguous and ordered."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    base, rem = divmod(int(n_sources), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_catalog(z, rank, world):
    """Processing order and ownership for a catalogue with redshifts `z`:
    returns (order, lo, hi): `order` sorts the catalogue by z (stable), rank owns order[lo:hi].
    The block boundary is moved forward to the next change of z so that no dictionary z-slice is
    split between two ranks (unless one slice alone exceeds a block)."""
    z = np.asarray(z, dtype=np.float64)
    order = np.argsort(z, kind="stable")
    zs = z[order]
    n = len(zs)

    def snap(i):
        if i <= 0 or i >= n:
            return max(0, min(n, i))
        j = i
        while j < n and zs[j] == zs[j - 1]:
            j += 1
        return j

    cuts = [snap(shard_bounds(n, r, world)[0]) for r in range(world)] + [n]
    for r in range(1, world + 1):                      # keep the cuts monotone
        cuts[r] = max(cuts[r], cuts[r - 1])
    return order, cuts[rank], cuts[rank + 1]


def gather_summaries(local, counts=None, dst=0, group=None):
    """Gather per-source summary rows ([n_local, k] float64 tensor; n_local may differ per rank)
    to rank `dst`.  Returns the concatenated [n_total, k] tensor on dst, None elsewhere.  Without an
    initialised process group (single process) returns `local` unchanged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    if counts is not None and list(counts) != sizes:
        raise ValueError("ranks report %s rows, expected %s" % (sizes, list(counts)))
    n_max = max(sizes) if sizes else 0
    k = local.shape[1]
    padded = torch.zeros((n_max, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    # all_gather of equal-size blocks works on every backend (nccl has no ragged gather)
    blocks = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(blocks, sizes)], dim=0)
