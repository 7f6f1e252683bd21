"""Row-block partitioned pressure CG across GPUs (one rank per GPU, fs_dist_* in
libfluidsim): the halo exchange of the search direction and the dot-product reductions
run inside the persistent CG kernel over NVLink peer memory.  ``torch.distributed`` is
used for rendezvous only: exchanging the CUDA-IPC handles and halo index lists at setup,
and summing two scalars per solve (mean of b, initial dots).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, ptr
from .parallel import row_block_partition


class PartitionedCG:
    """Solve K x = b for a symmetric positive (semi-)definite CSR matrix whose rows are
    split into contiguous blocks, rank r holding block r.

    ``rowptr, colidx, vals`` is the GLOBAL matrix (host arrays, identical on every rank;
    each rank keeps only its slice).  ``dist`` is an initialised torch.distributed module
    (or None for a single rank)."""

    def __init__(self, rowptr, colidx, vals, rank=0, world=1, dist=None, align=512):
        self.rank, self.world, self.dist = rank, world, dist
        self.n_global = len(rowptr) - 1
        part = row_block_partition(rowptr, colidx, rank, world, align=align)
        self.part = part
        self.lo, self.hi = part["lo"], part["hi"]
        self.n_own = self.hi - self.lo
        self.halo_global = part["halo_global"]
        self.n_halo = len(self.halo_global)
        a, b = part["nnz_range"]
        lv = np.ascontiguousarray(vals[a:b], dtype=np.float64)
        h = C.c_void_p()
        call("fs_dist_create", rank, world, self.n_own, self.n_halo, b - a, ptr(part["rowptr"]), ptr(part["colidx"]),
             ptr(lv), C.byref(h))
        self._h = h
        # ---- rendezvous: IPC handles and who needs which of my rows
        mine = (C.c_char * 64)()
        call("fs_dist_ipc_handle", self._h, mine)
        bounds = part["bounds"]
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine))
            needs = [None] * world
            dist.all_gather_object(needs, (self.halo_global, self.n_own))
        else:
            handles, needs = [bytes(mine)], [(self.halo_global, self.n_own)]
        rows, peers, dsts = [], [], []
        for q in range(world):
            if q == rank:
                continue
            halo_q, n_own_q = needs[q]
            sel = np.where((halo_q >= self.lo) & (halo_q < self.hi))[0]
            rows.append(halo_q[sel] - self.lo)
            peers.append(np.full(len(sel), q))
            dsts.append(n_own_q + sel)
        if rows:
            rows, peers, dsts = np.concatenate(rows), np.concatenate(peers), np.concatenate(dsts)
            order = np.argsort(rows, kind="stable")
            rows, peers, dsts = rows[order], peers[order], dsts[order]
        else:
            rows = peers = dsts = np.zeros(0)
        rows, peers, dsts = (np.ascontiguousarray(v, dtype=np.int32) for v in (rows, peers, dsts))
        allh = b"".join(handles)
        call("fs_dist_connect", self._h, allh, ptr(rows), ptr(peers), ptr(dsts), len(rows))
        self.n_send = len(rows)
        if world > 1:
            dist.barrier()
        self.last_ns = np.zeros(3)

    def __del__(self):
        try:
            if self._h:
                _lib.lib.fs_dist_destroy(self._h)
        except Exception:
            pass

    def _sum(self, arr):
        if self.world == 1:
            return np.asarray(arr, dtype=np.float64)
        import torch
        t = torch.tensor(np.asarray(arr, dtype=np.float64), device="cuda")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def solve(self, b_own, rtol=1e-10, maxit=200000, precond=1, project_mean=False):
        """b_own: this rank's rows of b (numpy or torch cuda).  Returns (x_own, iters, relres)."""
        is_t = _lib._is_torch(b_own)
        if project_mean:
            s = float(b_own.sum()) if not is_t else float(b_own.sum().item())
            mean = self._sum([s])[0] / self.n_global
            b_own = b_own - mean
        if not is_t:
            b_own = np.ascontiguousarray(b_own, dtype=np.float64)
        loc = np.zeros(2)
        call("fs_dist_cg_begin", self._h, ptr(b_own, np.float64, (self.n_own,)), precond, ptr(loc))
        bb, rz = self._sum(loc)
        x = np.empty(self.n_own) if not is_t else b_own.new_empty(self.n_own)
        it, rr = C.c_int(0), C.c_double(0)
        call("fs_dist_cg_run", self._h, float(bb), float(rz), ptr(x), rtol, maxit, precond, C.byref(it), C.byref(rr),
             ptr(self.last_ns))
        if project_mean:
            s = float(x.sum()) if not is_t else float(x.sum().item())
            x -= self._sum([s])[0] / self.n_global
        return x, it.value, rr.value
