"""Multi-GPU partitioning of the path (one process per GPU, torch.distributed).

The path shards in exactly two places; everything else is replicated:
  * REML probe / RHS columns: every rank assembles and factors V itself (inputs are replicated, recomputing
    is cheaper than broadcasting the factor), takes a contiguous slice of the sim_num probe columns through
    L*Z, the solves and the fused SpMM reductions, and the K partial trace sums are combined with ONE
    all-reduce of K doubles per evaluation.
  * HE row blocks: contiguous row ranges balanced by nonzeros; ONE all-reduce of 2K + 2K^2 doubles.
The helpers work on CPU tensors with the gloo backend as well, which is how the host logic is tested.
"""
import numpy as np


def active_group():
    """torch.distributed module if a process group with more than one rank is initialised, else None."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def rank_world():
    dist = active_group()
    return (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)


def column_block(ncols, rank, world):
    """Contiguous slice [lo, hi) of `ncols` probe columns owned by `rank` (sizes differ by at most one)."""
    return (ncols * rank) // world, (ncols * (rank + 1)) // world


def row_blocks_by_nnz(indptr, world):
    """world+1 row boundaries such that every block holds about nnz/world nonzeros (contiguous rows)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    targets = indptr[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate(([0], np.clip(cuts, 0, n), [n])).astype(np.int64)
    return np.maximum.accumulate(bounds)


def allreduce_sum_(t):
    """In-place sum over ranks (no-op without a process group).  Tiny payloads: latency, not bandwidth."""
    dist = active_group()
    if dist is not None:
        dist.all_reduce(t)
    return t
