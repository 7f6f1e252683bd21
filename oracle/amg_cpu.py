"""CPU statement (scipy + the OpenMP port in cg_port.c) of the smoothed-aggregation AMG-PCG that
libfluidsim uses for large pressure systems.  TEST INFRASTRUCTURE / CPU BASELINE ONLY -- imported by
tests/, bench.py's cpu_baseline and --impl reference legs and scripts/amg_prototype.py, never by the
product.

The hierarchy follows csrc/amg.cu step by step (two passes of hashed handshake matching over strong
couplings, filtered smoothed prolongator, Galerkin product, dense (pseudo-)inverse on the last level)
and reproduces its level sizes (2 098 689 -> 391 885 -> 67 528 -> 8 762 -> 960 at 4M triangles) and
iteration counts.  The reference has no multigrid (every solve is a dense LU, code/StokesColor.py:555):
this is the CPU arm of the SAME algorithm the GPU runs, next to the Jacobi-CG port that is the
closest sparse analogue of the reference's solve.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import cgport


def hash_edges(i, j):
    a = np.minimum(i, j).astype(np.uint64)
    b = np.maximum(i, j).astype(np.uint64)
    z = (a << np.uint64(32)) | b
    z = z + np.uint64(0x9e3779b97f4a7c15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def pairwise(A, rounds=8):
    """one pairwise pass: handshake matching over strong couplings (>= 0.5 max), hashed priorities"""
    A = A.tocsr()
    n = A.shape[0]
    Cm = A.tocoo()
    off = Cm.row != Cm.col
    i, j, w = Cm.row[off], Cm.col[off], -Cm.data[off]
    wmax = np.zeros(n)
    np.maximum.at(wmax, i, w)
    strong = (w > 0) & (w >= 0.5 * wmax[i])
    i, j, w = i[strong], j[strong], w[strong]
    h = hash_edges(i, j)
    state = -np.ones(n, dtype=np.int64)
    partner = -np.ones(n, dtype=np.int64)
    for _ in range(rounds):
        ok = (state[i] < 0) & (state[j] < 0)
        ii, jj, hh = i[ok], j[ok], h[ok]
        if len(ii) == 0:
            break
        order = np.lexsort((hh, ii))           # per row: last entry = highest hash
        ii, jj = ii[order], jj[order]
        last = np.r_[ii[1:] != ii[:-1], True]
        best = -np.ones(n, dtype=np.int64)
        best[ii[last]] = jj[last]
        cand = np.nonzero(best >= 0)[0]
        mutual = cand[best[best[cand]] == cand]
        partner[mutual] = best[mutual]
        state[mutual] = 1
    leader = np.arange(n)
    m = state >= 0
    leader[m] = np.minimum(np.arange(n)[m], partner[m])
    # leftovers join their strongest matched neighbour
    i2, j2, w2 = Cm.row[off], Cm.col[off], -Cm.data[off]
    ok = (state[i2] < 0) & (state[j2] >= 0) & (w2 > 0)
    i2, j2, w2 = i2[ok], j2[ok], w2[ok]
    order = np.lexsort((w2, i2))
    i2, j2 = i2[order], j2[order]
    last = np.r_[i2[1:] != i2[:-1], True]
    leader[i2[last]] = leader[j2[last]]
    ids = np.unique(leader, return_inverse=True)[1]
    return ids, ids.max() + 1


def aggregates(A, passes):
    agg, nc = pairwise(A)
    for _ in range(1, passes):
        Pt = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)), shape=(len(agg), nc))
        Ak = (Pt.T @ A @ Pt).tocsr()
        a2, n2 = pairwise(Ak)
        agg, nc = a2[agg], n2
    return agg, nc


def prolongator(A, agg, nc, omega_p=2 / 3, theta=0.25, steps=1):
    n = A.shape[0]
    Pt = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
    Cm = A.tocoo()
    off = Cm.row != Cm.col
    w = -Cm.data
    wmax = np.zeros(n)
    np.maximum.at(wmax, Cm.row[off], w[off])
    weak = off & (w < theta * wmax[Cm.row])
    data = Cm.data.copy()
    lump = np.zeros(n)
    np.add.at(lump, Cm.row[weak], data[weak])
    data[weak] = 0
    AF = sp.csr_matrix((data, (Cm.row, Cm.col)), shape=A.shape) + sp.diags(lump)
    AF.eliminate_zeros()
    dF = AF.diagonal()
    P = Pt
    for _ in range(steps):
        P = P - sp.diags(omega_p / dF) @ (AF @ P)
    return P.tocsr()


class Hier:
    def __init__(self, A, passes0=2, passes=2, omega_p=2 / 3, theta=0.25, psteps=1, min_rows=2048):
        self.A = [A.tocsr()]
        self.P = []
        while self.A[-1].shape[0] > min_rows:
            Ak = self.A[-1]
            agg, nc = aggregates(Ak, passes0 if len(self.A) == 1 else passes)
            if nc >= 0.8 * Ak.shape[0]:
                break
            P = prolongator(Ak, agg, nc, omega_p, theta, psteps)
            self.P.append(P)
            self.A.append((P.T @ Ak @ P).tocsr())
        Ac = self.A[-1].toarray()
        n = Ac.shape[0]
        singular = abs(Ac.sum(1)).max() < 1e-9 * abs(Ac.diagonal()).max()
        self.Cinv = (np.linalg.pinv(Ac + np.ones((n, n)) * (np.abs(Ac.diagonal()).max() / n)) if singular
                     else np.linalg.inv(Ac))
        self.D = [a.diagonal() for a in self.A]
        self.rho = [None] * len(self.A)

    def sizes(self):
        return [a.shape[0] for a in self.A], [a.nnz for a in self.A]

    def smooth(self, l, x, b, kind, omega, sweeps):
        import scipy.sparse.linalg as spla
        A, D = self.A[l], self.D[l]
        if kind == "jacobi":
            for _ in range(sweeps):
                x = x + omega * (b - A @ x) / D
            return x
        if kind == "cheby":
            if self.rho[l] is None:
                Dinv = sp.diags(1.0 / D)
                self.rho[l] = abs(spla.eigs(Dinv @ A, k=1, which="LM", return_eigenvectors=False, tol=1e-2)[0])
            lmax = 1.1 * self.rho[l]
            lmin = lmax / 4.0
            d, c = (lmax + lmin) / 2, (lmax - lmin) / 2
            r = (b - A @ x) / D
            p = r / d
            x = x + p
            alpha = 1.0 / d
            for k in range(1, sweeps):
                r = (b - A @ x) / D
                beta = (c * alpha / 2) ** 2 if k > 1 else 0.5 * (c * alpha) ** 2
                alpha = 1.0 / (d - beta / alpha)
                p = alpha * r + beta * p
                x = x + p
            return x
        raise ValueError(kind)

    def vcycle(self, l, b, kind="jacobi", omega=2 / 3, pre=1, post=1, gamma=1):
        if l == len(self.A) - 1:
            return self.Cinv @ b
        x = self.smooth(l, np.zeros_like(b), b, kind, omega, pre)
        r = b - self.A[l] @ x
        rc = self.P[l].T @ r
        xc = self.vcycle(l + 1, rc, kind, omega, pre, post, gamma)
        for _ in range(gamma - 1):
            xc = xc + self.vcycle(l + 1, rc - self.A[l + 1] @ xc, kind, omega, pre, post, gamma)
        x = x + self.P[l] @ xc
        return self.smooth(l, x, b, kind, omega, post)


def pcg(A, b, M, rtol=1e-10, maxit=400):
    """numpy PCG (prototype / cross-check of the C port); returns the iteration count"""
    b = b - b.mean()
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if r @ r <= rtol ** 2 * bb:
            return it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return maxit


# ---- the hierarchy handed to cg_port.c -----------------------------------------------------------
class _Level(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in
                                     ("a_rp", "a_ci", "a_v", "p_rp", "p_ci", "p_v", "r_rp", "r_ci", "r_v",
                                      "dinv", "x", "b", "t")]


class AmgPcg:
    """CG preconditioned by one V(1,1) damped-Jacobi cycle of `Hier`, all loops in oracle/_ref/libcgport.so
    (OpenMP).  solve(b, x0) -> (x, iterations, relres)."""

    def __init__(self, A, omega=2 / 3, **hier_kw):
        self.H = Hier(A, **hier_kw)
        self.omega = omega
        self.lib = cgport.load()
        self.lib.cgport_pcg_amg.restype = C.c_int
        self.lib.cgport_pcg_amg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                            C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
        self.lib.cgport_vcycle.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        self._keep = []
        nlev = len(self.H.A)
        self.levels = (_Level * nlev)()

        def csr(M):
            M = M.tocsr()
            M.sort_indices()
            arrs = (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32),
                    np.ascontiguousarray(M.data, dtype=np.float64))
            self._keep.append(arrs)
            return [a.ctypes.data for a in arrs]

        for l in range(nlev):
            lv = self.levels[l]
            n = self.H.A[l].shape[0]
            lv.n = n
            lv.a_rp, lv.a_ci, lv.a_v = csr(self.H.A[l])
            if l + 1 < nlev:
                lv.p_rp, lv.p_ci, lv.p_v = csr(self.H.P[l])
                lv.r_rp, lv.r_ci, lv.r_v = csr(self.H.P[l].T)
            work = [np.ascontiguousarray(1.0 / self.H.D[l]), np.zeros(n), np.zeros(n), np.zeros(n)]
            self._keep.append(work)
            lv.dinv, lv.x, lv.b, lv.t = (w.ctypes.data for w in work)
        self.cinv = np.ascontiguousarray(self.H.Cinv, dtype=np.float64)
        self.n = self.H.A[0].shape[0]

    def precond(self, r):
        z = np.empty(self.n)
        r = np.ascontiguousarray(r, dtype=np.float64)
        self.lib.cgport_vcycle(C.addressof(self.levels), len(self.levels), self.cinv.ctypes.data, self.omega,
                               r.ctypes.data, z.ctypes.data)
        return z

    def solve(self, b, x0=None, rtol=1e-10, maxit=1000, project_mean=True):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        rr = C.c_double(0)
        it = self.lib.cgport_pcg_amg(C.addressof(self.levels), len(self.levels), self.cinv.ctypes.data, self.omega,
                                     b.ctypes.data, x.ctypes.data, rtol, maxit, 1 if project_mean else 0, C.byref(rr))
        return x, it, rr.value
