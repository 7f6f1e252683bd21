"""One process per GPU: sharding helpers for the parts of the path that split with no
data-path communication -- squirmer (B1,B2) sweeps and tracer
ensembles -- plus the contiguous row-block partition used by the partitioned pressure CG.

``torch.distributed`` is plumbing only (rendezvous, the final count reductions, the
max-over-ranks timing); backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import os

import numpy as np


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world 1).
    Returns (rank, world, local_rank, dist_or_None)."""
    rank, world, local = env_rank()
    if world == 1:
        return rank, world, local, None
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group(backend, **kw)
    return rank, world, local, dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items for this rank; sizes differ by at most one."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def sweep_configs(n_b1=8, n_b2=8):
    """The 64 squirmer configurations of config 4: (B1,B2) in linspace(-4,-0.5,8) x linspace(-5,5,8)
    (B1, B2 enter the path only through makeDirBCU, code/StokesColor.py:419)."""
    b1 = np.linspace(-4.0, -0.5, n_b1)
    b2 = np.linspace(-5.0, 5.0, n_b2)
    return [(float(a), float(b)) for a in b1 for b in b2]


def shard_list(items, rank, world):
    """Round-robin shard: rank r takes items r, r+world, ... (equal counts when world | len)."""
    return list(items[rank::world])


def allreduce(value, op="sum", dist=None, device=None):
    """Reduce a python number over all ranks (identity for world 1)."""
    if dist is None:
        return value
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def row_block_partition(rowptr, colidx, rank, world, align=512):
    """Row-block partition of a square CSR matrix for the partitioned CG.

    Rank r owns rows [lo, hi) (block boundaries aligned to ``align`` rows, the CTA tile of the
    persistent kernel).  Returns a dict with the local CSR slice whose columns are renumbered
    to [0, n_own) for owned columns and [n_own, n_own + n_halo) for external ones, the sorted
    global ids of the halo columns, and, per neighbour rank, which halo slots it fills.
    """
    n = len(rowptr) - 1
    nblk = (n + align - 1) // align
    bounds = [min(n, ((nblk * r) // world) * align) for r in range(world)] + [n]
    lo, hi = bounds[rank], bounds[rank + 1]
    a, b = int(rowptr[lo]), int(rowptr[hi])
    cols = np.asarray(colidx[a:b], dtype=np.int64)
    local_ptr = (np.asarray(rowptr[lo:hi + 1], dtype=np.int64) - a).astype(np.int32)
    ext = (cols < lo) | (cols >= hi)
    halo = np.unique(cols[ext])
    local_cols = np.where(ext, (hi - lo) + np.searchsorted(halo, cols), cols - lo).astype(np.int32)
    owner = np.searchsorted(np.asarray(bounds[1:]), halo, side="right")
    recv = {int(r): np.where(owner == r)[0].astype(np.int32) for r in np.unique(owner)}
    return dict(lo=lo, hi=hi, bounds=bounds, rowptr=local_ptr, colidx=local_cols, nnz_range=(a, b),
                halo_global=halo.astype(np.int64), recv_slots=recv)
