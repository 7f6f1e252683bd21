"""Iteration counts of the pressure solves over the bench trajectory: time-extrapolated warm start (round 1) against
Fischer's A-orthonormal projection onto previous solutions with a restart, a sliding-window Gram formulation, and the
compressed basis that csrc/recycle.cu implements.  CPU prototype on oracle/cpu_step.py.
    python scripts/proto_projected_guess.py 512 256 25"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import cpu_step as CS
import importlib.util
spec = importlib.util.spec_from_file_location("hostmesh", "puc-fluidsimulation-project_b200/hostmesh.py")
hm = importlib.util.module_from_spec(spec); spec.loader.exec_module(hm)

class Fischer:
    def __init__(self, K, kmax, drop="restart"):
        self.K, self.kmax, self.X, self.drop = K, kmax, [], drop
    def guess(self, r):
        if not self.X: return None
        x0 = np.zeros_like(r)
        self.alpha = [xt @ r for xt in self.X]
        for a, xt in zip(self.alpha, self.X): x0 += a * xt
        return x0
    def update(self, q, x0):
        d = q - (x0 if x0 is not None else 0.0)
        d = d - d.mean()
        if len(self.X) >= self.kmax:
            if self.drop == "restart":
                self.X = []
                d = q - q.mean()
            elif self.drop == "oldest":
                self.X.pop(0)
            elif self.drop == "minalpha":
                self.X.pop(int(np.argmin(np.abs(self.alpha))))
        Ad = self.K.dot(d)
        for xt in self.X:
            d = d - (xt @ Ad) * xt
        Ad = self.K.dot(d)
        nrm = np.sqrt(d @ Ad)
        if nrm > 0: self.X.append(d / nrm)

class FischerC:
    """Fischer's A-orthonormal basis; when it is full it is compressed to an orthonormal basis of the span of the last m
    solutions (known through their coordinates in the old basis) instead of being thrown away."""
    def __init__(self, K, kmax, drop="compress3"):
        self.K, self.kmax, self.m, self.X, self.C = K, kmax, int(drop.replace("compress", "")), [], []
    def guess(self, r):
        if not self.X: return None
        self.alpha = np.array([xt @ r for xt in self.X])
        x0 = np.zeros_like(r)
        for a, xt in zip(self.alpha, self.X): x0 += a * xt
        return x0
    def update(self, q, x0):
        k = len(self.X)
        d = q - (x0 if x0 is not None else 0.0)
        d = d - d.mean()
        Ad = self.K.dot(d)
        c = np.array([xt @ Ad for xt in self.X]) if k else np.zeros(0)
        for ci, xt in zip(c, self.X): d = d - ci * xt
        Ad = self.K.dot(d)
        nrm = np.sqrt(d @ Ad)
        self.X.append(d / nrm)
        coords = np.r_[(self.alpha + c) if k else np.zeros(0), nrm]
        self.C = [np.r_[cc, 0.0] for cc in self.C] + [coords]
        self.C = self.C[-self.m:]
        if len(self.X) >= self.kmax:
            M = np.array(self.C[::-1]).T            # newest first: it becomes the first basis vector
            Q, R = np.linalg.qr(M)
            keep = np.abs(np.diag(R)) > 1e-10 * np.abs(R[0, 0])
            Q, R = Q[:, keep], R[np.ix_(keep, keep)]
            newX = [sum(Q[i, j] * self.X[i] for i in range(len(self.X))) for j in range(Q.shape[1])]
            self.C = [Q.T @ cc for cc in self.C]
            self.X = newX

class Gram:
    """sliding window of the last k solutions; G_ij = x_i . A x_j ~ x_i . b_j is filled from the projection's own dots"""
    def __init__(self, K, kmax, drop="gram", eps=1e-12):
        self.k, self.X, self.G, self.eps = kmax, [], np.zeros((0, 0)), eps
    def guess(self, r):
        if not self.X: return None
        self.g = np.array([x @ r for x in self.X])
        w, V = np.linalg.eigh((self.G + self.G.T) / 2)
        keep = w > self.eps * w.max()
        c = V[:, keep] @ ((V[:, keep].T @ self.g) / w[keep])
        x0 = np.zeros_like(r)
        for ci, x in zip(c, self.X): x0 += ci * x
        self.r = r
        return x0
    def update(self, q, x0):
        q = q - q.mean()
        k = len(self.X)
        g = self.g if k else np.zeros(0)
        G = np.zeros((k + 1, k + 1))
        G[:k, :k] = self.G
        G[:k, k] = g; G[k, :k] = g
        G[k, k] = q @ self.r if k else None
        if not k: G[0, 0] = q @ self._b0
        self.X.append(q); self.G = G
        if len(self.X) > self.k:
            self.X.pop(0); self.G = self.G[1:, 1:]

def run(nt, nr, steps, mode, kmax=8, drop="restart"):
    nodes, markers, tris = hm.square_with_hole(nt, nr)
    s = CS.CpuStokes(nodes, markers, tris, B1=-2.0, B2=-5.0, precond="amg", recycle=False)    # the variants below replace the guess
    if mode == "fischer":
        F = [(Gram if drop.startswith('gram') else FischerC if drop.startswith('compress') else Fischer)(s.K, kmax, drop) for _ in range(2)]
        def pressure(b_nodes, h):
            ps = s.psys
            r = np.bincount(ps.dof, weights=ps.M * b_nodes, minlength=ps.nd)
            rm = r - r.mean()
            f = F[0] if h is s.hist[0] else F[1]
            f._b0 = rm; f.r = rm
            x0 = f.guess(rm)
            q, it, _ = s.amg.solve(r, x0=x0, rtol=s.rtol_p, project_mean=True)
            f.update(q, x0)
            return q[ps.dof], it
        s._pressure = pressure
    its = []
    for k in range(steps):
        its.append(s.step())
    return its

nt, nr, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
a = run(nt, nr, steps, "extrap")
print("time-extrapolated guess   :", [(i[1], i[2]) for i in a], "sum over steps 6..", sum(i[1] + i[2] for i in a[5:]), flush=True)
for k, drop in ((12, "restart"), (12, "gram"), (12, "compress6"), (16, "compress8")):
    b = run(nt, nr, steps, "fischer", k, drop)
    print(f"k={k:2d} {drop:10s}           :", [(i[1], i[2]) for i in b], "sum over steps 6..", sum(i[1] + i[2] for i in b[5:]), flush=True)
