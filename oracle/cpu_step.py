"""The operator-split Stokes step (code/StokesColor.py:537-575) as a multi-threaded CPU program at
sizes the dense reference cannot hold: the restated oracle's algorithm (oracle/restated.py:
RestatedStokes.flow_step) with every O(N) loop in oracle/_ref/libcgport.so (OpenMP) --
  * divergence / gradient as sparse operators Dx, Dy (the same element sums as
    restated.divergence / restated.gradient; summation order differs, values agree to rounding),
  * viscous solves: Jacobi-CG on A_visc, started from u (what libfluidsim does),
  * pressure solves on the periodic-merged SPD operator: Jacobi-CG (precond="jacobi", the sparse
    analogue of the reference's np.linalg.solve) or AMG-PCG (precond="amg", the algorithm of the GPU
    arm), warm-started with libfluidsim's policy (best of q, 2q - q1, 3q - 3q1 + q2 by residual).
TEST INFRASTRUCTURE / CPU BASELINE ONLY: used by tests/ and by bench.py's cpu_baseline and
--impl reference legs; never imported by the product.
"""
import numpy as np
import scipy.sparse as sp

from . import cgport
from . import restated as R


def grad_operators(nodes, tris):
    """Dx, Dy (N x N CSR) with  divergence(u) = Dx u_x + Dy u_y  and  gradient(p) = (Dx p, Dy p)
    (code/StokesColor.py:130-165, :224-263: both are the same area-weighted element sums)."""
    n = nodes.shape[0]
    tris = np.asarray(tris)
    det, yd, xd = R._geom(nodes, tris)
    keep, third, area_sum = R._area_sum(nodes, tris, det)
    with np.errstate(divide="ignore"):
        inv2a = 1.0 / det
    t = tris[keep]
    w = (third * inv2a)[keep]
    rows = np.repeat(t, 3, axis=1).ravel()                 # corner n of e, three times
    cols = np.tile(t, (1, 3)).ravel()                      # the three corners i of e
    scale = 1.0 / (area_sum + 1e-12)
    out = []
    for d in (yd[keep], xd[keep]):
        vals = (np.tile(d * w[:, None], (1, 3))).ravel() * scale[rows]
        M = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        M.sum_duplicates()
        M.sort_indices()
        out.append(M)
    return out[0], out[1]


class _Csr:
    def __init__(self, M):
        M = M.tocsr()
        M.sort_indices()
        self.rowptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(M.indices, dtype=np.int32)
        self.vals = np.ascontiguousarray(M.data, dtype=np.float64)

    def dot(self, x):
        return cgport.spmv(self.rowptr, self.colidx, self.vals, np.ascontiguousarray(x, dtype=np.float64))


class Recycler:
    """Solution-subspace projection for successive right-hand sides, as csrc/recycle.cu does it on the GPU: an
    A-orthonormal basis of previous solutions (Fischer 1998), compressed to the span of the last `keep` solutions
    when it reaches `kmax` vectors.  The basis is one (kmax + keep, n) array; the multi-dot and combination loops are
    the OpenMP ones of oracle/cg_port.c, so that the CPU arm of bench.py pays for the algorithm, not for numpy."""

    def __init__(self, K, kmax=12, keep=6):
        self.K, self.kmax, self.keep = K, kmax, keep
        self.X, self.k, self.C = None, 0, []
        self.alpha, self.x0 = None, None

    # the three operations a row-partitioned variant replaces (tests/test_parallel_gloo.py: sums all-reduced, K on row blocks)
    def _dots(self, v):
        return cgport.multidot(self.X, self.k, np.ascontiguousarray(v, dtype=np.float64))

    def _dot(self, a, b):
        return float(a @ b)

    def _Kdot(self, d):
        return self.K.dot(d)

    def guess(self, b):
        self.x0 = None
        if not self.k:
            return None
        self.alpha = self._dots(b)
        self.x0 = cgport.comb(self.X, self.k, self.alpha)
        return self.x0

    def update(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        if self.X is None:
            self.X = np.zeros((self.kmax + self.keep, q.size))
        if self.x0 is None and self.k:
            self.k, self.C = 0, []
        k = self.k
        if k:
            d = q - self.x0
            c = self._dots(self._Kdot(d))
            d = cgport.comb(self.X, k, -c, 1.0, d)
            coords = list(self.alpha + c)
        else:
            d, coords = q.copy(), []
        self.x0 = None
        nrm2 = self._dot(d, self._Kdot(d))
        have2 = float(np.dot(coords, coords)) if coords else 0.0
        if nrm2 > 0.0 and nrm2 > 1e-26 * have2:
            nrm = np.sqrt(nrm2)
            self.X[k] = d / nrm
            self.k = k = k + 1
            self.C = [cc + [0.0] for cc in self.C]
            coords = coords + [nrm]
        elif not k:
            return
        self.C = (self.C + [coords])[-self.keep:]
        if k < self.kmax:
            return
        Q, first = [], 0.0
        for v in (np.array(cc) for cc in self.C[::-1]):
            for _ in range(2):
                for qv in Q:
                    v = v - (qv @ v) * qv
            nv = np.linalg.norm(v)
            if not Q:
                first = nv
            if nv > 1e-10 * first and nv > 0.0:
                Q.append(v / nv)
        for j, qv in enumerate(Q):
            cgport.comb(self.X, k, qv, out=self.X[self.kmax + j])
        self.X[:len(Q)] = self.X[self.kmax:self.kmax + len(Q)]
        self.C = [[float(qv @ np.array(cc)) for qv in Q] for cc in self.C]
        self.k = len(Q)


RECYCLE_MIN_ROWS = 20000     # csrc/stokes.cu: kRecycleMinRows


class CpuStokes:
    def __init__(self, nodes, markers, tris, B1=-2.0, B2=0.0, DT=0.05, v=0.1, precond="amg",
                 rtol_pressure=1e-10, rtol_visc=1e-12, H=1.0, tol=1e-6, threads=None, exact_ops=False, recycle=None):
        # exact_ops: divergence / gradient by restated.divergence / restated.gradient (the element sums of the
        # reference in the reference's order) instead of the pre-assembled sparse operators
        self.exact_ops = exact_ops
        if threads:
            cgport.load().cgport_set_threads(int(threads))
        self.nodes, self.markers, self.tris = nodes, markers, np.asarray(tris)
        n = nodes.shape[0]
        self.N = n
        self.B1, self.B2, self.DT, self.v = B1, B2, DT, v
        self.precond, self.rtol_p, self.rtol_v = precond, rtol_pressure, rtol_visc
        self.pairs = R.filter_wall_pairs(nodes, R.find_boundary_pairs(nodes, 1.0, tol), H, tol)
        self.wall, self.inner_b, self.dirichlet, self.interior = R.index_sets(nodes, markers, H, tol)
        rowptr, colidx, scatter = R.csr_pattern(n, tris)
        kvals = R.assemble_stiffness(nodes, tris, rowptr, colidx, scatter)
        av = R.viscous_matrix(n, rowptr, colidx, kvals, self.dirichlet, DT, v)
        self.A_visc = _Csr(sp.csr_matrix((av, colidx, rowptr), shape=(n, n)))
        self.psys = R.PressureSystem(nodes, tris, self.pairs)
        self.K = _Csr(self.psys.K)
        self.amg = None
        if precond == "amg":
            from .amg_cpu import AmgPcg
            self.amg = AmgPcg(self.psys.K)
        Dx, Dy = grad_operators(nodes, tris)
        self.Dx, self.Dy = _Csr(Dx), _Csr(Dy)
        self.u = np.zeros((n, 2))
        R.make_dir_bcu(self.u, nodes, self.wall, self.inner_b, B1, B2)
        self.hist = [dict(q=None, q1=None, q2=None), dict(q=None, q1=None, q2=None)]
        if recycle is None:
            recycle = self.psys.nd >= RECYCLE_MIN_ROWS
        self.rec = [Recycler(self.K), Recycler(self.K)] if recycle else None
        self.iters = (0, 0, 0)

    def divergence(self, u):
        if self.exact_ops:
            return R.divergence(self.nodes, self.tris, u)
        return self.Dx.dot(u[:, 0]) + self.Dy.dot(u[:, 1])

    def gradient(self, p):
        if self.exact_ops:
            return R.gradient(self.nodes, self.tris, p)
        return self.Dx.dot(p), self.Dy.dot(p)

    def _visc(self, rhs):
        x, it, _ = cgport.cg(self.A_visc.rowptr, self.A_visc.colidx, self.A_visc.vals, rhs, x0=rhs,
                             rtol=self.rtol_v, jacobi=True)
        return x, it

    def _pressure(self, b_nodes, h):
        ps = self.psys
        r = np.bincount(ps.dof, weights=ps.M * b_nodes, minlength=ps.nd)
        x0 = None
        rec = self.rec[0 if h is self.hist[0] else 1] if self.rec else None
        if rec is not None:
            x0 = rec.guess(r - r.mean()) if h["q"] is not None else None
            if x0 is None and h["q"] is not None:
                x0 = h["q"]
        elif h["q"] is not None:
            rm = r - r.mean()
            cands = [h["q"]]
            if h["q1"] is not None:
                cands.append(2.0 * h["q"] - h["q1"])
            if h["q2"] is not None:
                cands.append(3.0 * h["q"] - 3.0 * h["q1"] + h["q2"])
            res = [np.linalg.norm(rm - self.K.dot(c)) for c in cands]
            x0 = cands[int(np.argmin(res))]
        if self.amg is not None:
            q, it, _ = self.amg.solve(r, x0=x0, rtol=self.rtol_p, project_mean=True)
        else:
            q, it, _ = cgport.cg(self.K.rowptr, self.K.colidx, self.K.vals, r, x0=x0, rtol=self.rtol_p, jacobi=True,
                                 project_mean=True)
        h["q2"], h["q1"], h["q"] = h["q1"], h["q"], q
        if rec is not None:
            rec.update(q)
        return q[ps.dof], it

    def step(self):
        DT, nodes, u = self.DT, self.nodes, self.u
        ux, i0 = self._visc(u[:, 0].copy())
        uy, i1 = self._visc(u[:, 1].copy())
        us = np.stack([ux, uy], axis=1)
        R.make_per_bcu(us, self.pairs)
        R.make_dir_bcu(us, nodes, self.wall, self.inner_b, self.B1, self.B2)
        p, itp = self._pressure(-(1.0 / DT) * self.divergence(us), self.hist[0])
        gx, gy = self.gradient(p)
        u[:, 0] = us[:, 0] - DT * gx
        u[:, 1] = us[:, 1] - DT * gy
        R.make_per_bcu(u, self.pairs)
        R.make_dir_bcu(u, nodes, self.wall, self.inner_b, self.B1, self.B2)
        p2, itp2 = self._pressure(-(1.0 / DT) * self.divergence(u), self.hist[1])
        g2x, g2y = self.gradient(p2)
        u[self.interior, 0] -= DT * g2x[self.interior]
        u[self.interior, 1] -= DT * g2y[self.interior]
        self.p, self.p2 = p, p2
        self.iters = (max(i0, i1), itp, itp2)
        return self.iters
