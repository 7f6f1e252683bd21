"""CPU prototype of the AMG-PCG (scipy) to compare cycle / coarsening variants by iteration count.
Not the product: a research tool for the next round.  python scripts/amg_prototype.py [n_theta n_r]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import fluidsim_b200 as fb
from oracle import restated as R

def hash_edges(i, j):
    a = np.minimum(i, j).astype(np.uint64); b = np.maximum(i, j).astype(np.uint64)
    z = (a << np.uint64(32)) | b
    z = z + np.uint64(0x9e3779b97f4a7c15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))

def pairwise(A, rounds=8):
    """one pairwise pass: handshake matching over strong couplings (>= 0.5 max), hashed priorities"""
    A = A.tocsr(); n = A.shape[0]
    C = A.tocoo()
    off = C.row != C.col
    i, j, w = C.row[off], C.col[off], -C.data[off]
    wmax = np.zeros(n); np.maximum.at(wmax, i, w)
    strong = (w > 0) & (w >= 0.5 * wmax[i])
    i, j, w = i[strong], j[strong], w[strong]
    h = hash_edges(i, j)
    state = -np.ones(n, dtype=np.int64); partner = -np.ones(n, dtype=np.int64)
    for _ in range(rounds):
        ok = (state[i] < 0) & (state[j] < 0)
        ii, jj, hh = i[ok], j[ok], h[ok]
        if len(ii) == 0: break
        order = np.lexsort((hh, ii))           # per row: last entry = highest hash
        ii, jj = ii[order], jj[order]
        last = np.r_[ii[1:] != ii[:-1], True]
        best = -np.ones(n, dtype=np.int64); best[ii[last]] = jj[last]
        cand = np.nonzero(best >= 0)[0]
        mutual = cand[best[best[cand]] == cand]
        partner[mutual] = best[mutual]; state[mutual] = 1
    leader = np.arange(n)
    m = state >= 0
    leader[m] = np.minimum(np.arange(n)[m], partner[m])
    # leftovers: join strongest matched neighbour
    C = A.tocoo(); off = C.row != C.col
    i2, j2, w2 = C.row[off], C.col[off], -C.data[off]
    ok = (state[i2] < 0) & (state[j2] >= 0) & (w2 > 0)
    i2, j2, w2 = i2[ok], j2[ok], w2[ok]
    order = np.lexsort((w2, i2)); i2, j2 = i2[order], j2[order]
    last = np.r_[i2[1:] != i2[:-1], True]
    leader[i2[last]] = leader[j2[last]]
    ids = np.unique(leader, return_inverse=True)[1]
    return ids, ids.max() + 1

def aggregates(A, passes):
    agg, nc = pairwise(A)
    Ak = None
    for _ in range(1, passes):
        Pt = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)), shape=(len(agg), nc))
        Ak = (Pt.T @ A @ Pt).tocsr()
        a2, n2 = pairwise(Ak)
        agg, nc = a2[agg], n2
    return agg, nc

def prolongator(A, agg, nc, omega_p=2/3, theta=0.25, steps=1):
    n = A.shape[0]
    Pt = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
    C = A.tocoo(); off = C.row != C.col
    w = -C.data; wmax = np.zeros(n); np.maximum.at(wmax, C.row[off], w[off])
    weak = off & (w < theta * wmax[C.row])
    data = C.data.copy(); lump = np.zeros(n); np.add.at(lump, C.row[weak], data[weak]); data[weak] = 0
    AF = sp.csr_matrix((data, (C.row, C.col)), shape=A.shape) + sp.diags(lump)
    AF.eliminate_zeros()
    dF = AF.diagonal()
    P = Pt
    for _ in range(steps):
        P = P - sp.diags(omega_p / dF) @ (AF @ P)
    return P.tocsr()

class Hier:
    def __init__(self, A, passes0=2, passes=2, omega_p=2/3, theta=0.25, psteps=1, min_rows=2048):
        self.A = [A.tocsr()]; self.P = []
        while self.A[-1].shape[0] > min_rows:
            Ak = self.A[-1]
            agg, nc = aggregates(Ak, passes0 if len(self.A) == 1 else passes)
            if nc >= 0.8 * Ak.shape[0]: break
            P = prolongator(Ak, agg, nc, omega_p, theta, psteps)
            self.P.append(P); self.A.append((P.T @ Ak @ P).tocsr())
        Ac = self.A[-1].toarray(); n = Ac.shape[0]
        self.Cinv = np.linalg.pinv(Ac + np.ones((n, n)) * (np.abs(Ac.diagonal()).max() / n)) if abs(Ac.sum(1)).max() < 1e-9 * abs(Ac.diagonal()).max() else np.linalg.inv(Ac)
        self.D = [a.diagonal() for a in self.A]
        self.rho = [None] * len(self.A)
    def sizes(self): return [a.shape[0] for a in self.A], [a.nnz for a in self.A]
    def smooth(self, l, x, b, kind, omega, sweeps):
        A, D = self.A[l], self.D[l]
        if kind == "jacobi":
            for _ in range(sweeps): x = x + omega * (b - A @ x) / D
            return x
        if kind == "cheby":      # Chebyshev on D^-1 A over [rho/30*?...]: standard [0.3 rho? ] use [rho/4, 1.1 rho] hmm: smoothing interval
            if self.rho[l] is None:
                Dinv = sp.diags(1.0 / D)
                self.rho[l] = abs(spla.eigs(Dinv @ A, k=1, which="LM", return_eigenvectors=False, tol=1e-2)[0])
            lmax = 1.1 * self.rho[l]; lmin = lmax / 4.0
            d, c = (lmax + lmin) / 2, (lmax - lmin) / 2
            r = (b - A @ x) / D
            p = r / d; x = x + p
            alpha = 1.0 / d
            for k in range(1, sweeps):
                r = (b - A @ x) / D
                beta = (c * alpha / 2) ** 2 if k > 1 else 0.5 * (c * alpha) ** 2
                alpha = 1.0 / (d - beta / alpha)
                p = alpha * r + beta * p   # simplified three-term
                x = x + p
            return x
        raise ValueError(kind)
    def vcycle(self, l, b, kind="jacobi", omega=2/3, pre=1, post=1, gamma=1):
        if l == len(self.A) - 1: return self.Cinv @ b
        x = self.smooth(l, np.zeros_like(b), b, kind, omega, pre)
        r = b - self.A[l] @ x
        rc = self.P[l].T @ r
        xc = self.vcycle(l + 1, rc, kind, omega, pre, post, gamma)
        for _ in range(gamma - 1):
            xc = xc + self.vcycle(l + 1, rc - self.A[l + 1] @ xc, kind, omega, pre, post, gamma)
        x = x + self.P[l] @ xc
        return self.smooth(l, x, b, kind, omega, post)

def pcg(A, b, M, rtol=1e-10, maxit=400):
    b = b - b.mean(); x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, maxit + 1):
        Ap = A @ p; alpha = rz / (p @ Ap); x += alpha * p; r -= alpha * Ap
        if r @ r <= rtol ** 2 * bb: return it
        z = M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit

if __name__ == "__main__":
    nt, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 128)
    nodes, markers, tris = fb.square_with_hole(nt, nr)
    pairs = fb.filter_wall_pairs(nodes, fb.find_boundary_pairs(nodes))
    ps = R.PressureSystem(nodes, tris, pairs)
    A = sp.csr_matrix((ps.vals, ps.colidx, ps.rowptr))
    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    print("n", A.shape[0], "nnz", A.nnz)
    def run(tag, hk, ck):
        t0 = time.time(); H = Hier(A, **hk); ts = time.time() - t0
        n, z = H.sizes()
        it = pcg(A, b, lambda r: H.vcycle(0, r, **ck))
        cx = sum(z) / z[0]
        print(f"{tag:46s} levels {n}  op-complexity {cx:.2f}  iters {it}  (setup {ts:.1f}s)", flush=True)
    run("baseline 2+2 passes, V(1,1) jacobi", {}, {})
    run("V(2,2) jacobi", {}, dict(pre=2, post=2))
    run("V(1,1) jacobi omega 0.8", {}, dict(omega=0.8))
    run("W-cycle (gamma=2) V(1,1)", {}, dict(gamma=2))
    run("theta 0 (unfiltered)", dict(theta=0.0), {})
    run("omega_p 0.5", dict(omega_p=0.5), {})
    run("prolongator smoothed twice", dict(psteps=2), {})
    run("1 pass at finest level (pairs)", dict(passes0=1), {})
    run("3 passes everywhere", dict(passes0=3, passes=3), {})
    run("3 passes + prolongator smoothed twice", dict(passes0=3, passes=3, psteps=2), {})
    run("cheby(2) smoother", {}, dict(kind="cheby", pre=2, post=2))
    run("cheby(3) smoother", {}, dict(kind="cheby", pre=3, post=3))
