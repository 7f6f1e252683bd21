"""The Stokes step on a mesh cut into contiguous node blocks, one block (process, GPU) per rank:
``PartitionedStokes`` is ``StokesSolver`` for BASELINE config 5 (the 32M-triangle synthetic annulus on
1/2/4/8 GPUs).  Same reference sequence (code/StokesColor.py:537-575), same solvers (2-RHS Jacobi-CG,
AMG-preconditioned pressure CG with the single-GPU hierarchy), every kernel on this rank's rows only.

Halo values and dot products travel as peer stores over NVLink issued from inside the CUDA kernels
(csrc/dist.cuh); ``torch.distributed`` is used once, to exchange the 64-byte CUDA-IPC handles.
Setup is replicated (every rank builds the global operators and the AMG hierarchy, keeps its blocks
and frees the rest); the time loop is fully partitioned.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from . import hostmesh
from ._lib import StokesOpts, StokesStats, call, ptr
from .core import Mesh
from .stokes import StokesSolver


class PartitionedStokes:
    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, rank=0, world=1, dist=None, align=1,
                 node_split=None, gather_rows=100000, B1=-2.0, B2=0.0, DT=0.05, v=0.1, rtol_pressure=1e-10,
                 rtol_visc=1e-12, warm_start=True, maxit=200000, **kw):
        self.rank, self.world, self.dist = rank, world, dist
        self.B1, self.B2, self.DT, self.v = B1, B2, DT, v
        # replicated global setup: operators, index sets, initial velocity (code/StokesColor.py:442-483)
        glob = StokesSolver(nodes_coords, nodes_boundary_markers, triangles, B1=B1, B2=B2, DT=DT, v=v,
                            rtol_pressure=rtol_pressure, rtol_visc=rtol_visc, **kw)
        self.N = glob.N
        self.split = list(node_split) if node_split is not None else hostmesh.node_block_split(self.N, world, align)
        self.lo, self.hi = int(self.split[rank]), int(self.split[rank + 1])
        self.n_own = self.hi - self.lo
        # this rank's sub-mesh and boundary sets
        ln, lm, lt, l2g, _ = hostmesh.sub_mesh(glob.nodes_coords, glob.nodes_boundary_markers, glob.triangles, self.lo, self.hi)
        wall, inner, interior, pairs = hostmesh.local_index_sets(self.lo, self.hi, glob.wall_node_indices,
                                                                 glob.inner_boundary_indices, glob.interior, glob.pairs)
        self.l2g = l2g
        self.mesh = Mesh(ln, lt, lm)
        self.mesh.set_bc(wall, inner, pairs, interior)
        h = C.c_void_p()
        split = np.ascontiguousarray(self.split, dtype=np.int64)
        # FS_GATHER_ROWS: experiment switch for the level from which the hierarchy is replicated (global rows)
        gather_rows = int(os.environ.get("FS_GATHER_ROWS", gather_rows))
        call("fs_pstokes_create", glob._h, self.mesh._h, rank, world, ptr(split), ptr(l2g, np.int32), int(gather_rows), C.byref(h))
        self._h = h
        self.u = np.ascontiguousarray(glob.u[self.lo:self.hi])      # this rank's rows of the velocity
        del glob                                                    # global operators are no longer needed
        mine = (C.c_char * 64)()
        call("fs_pstokes_ipc_handle", self._h, mine)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine))
            call("fs_pstokes_connect", self._h, b"".join(handles))
            dist.barrier()
        else:
            call("fs_pstokes_connect", self._h, bytes(mine))
        self.opts = StokesOpts()
        call("fs_stokes_default_opts", C.byref(self.opts))
        self.opts.rtol_pressure, self.opts.rtol_visc = rtol_pressure, rtol_visc
        self.opts.precond = 2
        self.opts.warm_start = 1 if warm_start else 0
        self.opts.maxit = maxit
        self.stats = StokesStats()
        sz = [C.c_int64(0) for _ in range(4)]
        lv = C.c_int32(0)
        call("fs_pstokes_sizes", self._h, *[C.byref(s) for s in sz], C.byref(lv))
        self.n_halo_nodes, self.n_own_dofs, self.n_halo_dofs = sz[1].value, sz[2].value, sz[3].value
        self.levels_partitioned = lv.value

    def __del__(self):
        try:
            if self._h:
                _lib.lib.fs_pstokes_destroy(self._h)
        except Exception:
            pass

    def step(self, u=None):
        """Advance this rank's rows of the velocity in place ((n_own, 2): numpy or torch cuda tensor).
        Collective: every rank calls it."""
        if u is None:
            u = self.u
        call("fs_pstokes_step", self._h, ptr(u, np.float64, (self.n_own, 2), "u"), float(self.B1), float(self.B2),
             C.byref(self.opts), C.byref(self.stats))
        return self.stats

    def pressure(self):
        p, p2 = np.empty(self.n_own), np.empty(self.n_own)
        call("fs_pstokes_pressure", self._h, ptr(p), ptr(p2))
        return p, p2

    def get_state(self):
        """Warm-start state of this rank: the solution history of the two pressure solves, then (large systems) the
        bases the next guesses are projected onto."""
        base = 10 * self.n_own_dofs + 4
        need = C.c_int64(0)
        call("fs_pstokes_recycle_state", self._h, None, 0, 0, C.byref(need))
        q = np.empty(base + need.value)
        call("fs_pstokes_state", self._h, ptr(q), 0)
        call("fs_pstokes_recycle_state", self._h, C.c_void_p(q.ctypes.data + 8 * base), need.value, 0, C.byref(need))
        return q

    def set_state(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        base = 10 * self.n_own_dofs + 4
        if q.size < base:
            raise ValueError("state size mismatch")
        call("fs_pstokes_state", self._h, ptr(q), 1)
        call("fs_pstokes_recycle_state", self._h, C.c_void_p(q.ctypes.data + 8 * base) if q.size > base else None,
             q.size - base, 1, None)

    def profile_pcg(self, iters=40):
        """µs per AMG-PCG iteration on the last pressure right-hand side (fixed iteration count; collective)."""
        us = C.c_double(0)
        call("fs_pstokes_profile_pcg", self._h, int(iters), C.byref(us))
        return us.value

    def trace(self, cap=1 << 16):
        """Event log of the partitioned kernels (FS_DIST_TRACE=1): (n,2) uint64 array of {tag, ns}; resets the log."""
        buf = np.zeros((cap, 2), dtype=np.uint64)
        n = C.c_int64(0)
        call("fs_pstokes_trace", self._h, ptr(buf), cap, C.byref(n))
        return buf[:n.value]

    def gather(self, x_own):
        """All ranks' blocks of a nodal array, concatenated (host; for checks and output)."""
        x_own = np.ascontiguousarray(x_own.cpu().numpy() if _lib._is_torch(x_own) else x_own)
        if self.world == 1:
            return x_own
        parts = [None] * self.world
        self.dist.all_gather_object(parts, x_own)
        return np.concatenate(parts)
