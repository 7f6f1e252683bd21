"""Does an independent dense LU (unblocked, partial pivoting, no FMA) on the reference's literal A_pressure reproduce the
literal trajectory (np.linalg.solve = LAPACK getrf/getrs)?  And how sensitive is it to 1e-12 noise in the viscous solve?"""
import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import literal as L

def lu_factor(A):
    A = A.copy(); n = A.shape[0]; piv = np.arange(n)
    for k in range(n - 1):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]] = A[[p, k]]; piv[[k, p]] = piv[[p, k]]
        A[k + 1:, k] /= A[k, k]
        A[k + 1:, k + 1:] -= np.outer(A[k + 1:, k], A[k, k + 1:])
    return A, piv

def lu_solve(LU, piv, b):
    n = LU.shape[0]; y = b[piv].copy()
    for i in range(1, n):
        y[i] -= LU[i, :i] @ y[:i]
    for i in range(n - 1, -1, -1):
        y[i] = (y[i] - LU[i, i + 1:] @ y[i + 1:]) / LU[i, i]
    return y

mesh = sys.argv[1] if len(sys.argv) > 1 else "mesh5.1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ref = L.LiteralStokes(mesh, B1=-2.0, B2=-5.0)
cache = {}
rng = np.random.default_rng(0)
def make_solver(noise):
    def solve(A, b):
        if A is test.A_pressure:
            if "lu" not in cache: cache["lu"] = lu_factor(A)
            return lu_solve(*cache["lu"], b)
        x = np.linalg.solve(A, b)
        return x * (1.0 + noise * rng.standard_normal(x.shape)) if noise else x
    return solve
for noise in (0.0, 1e-12):
    test = L.LiteralStokes(mesh, B1=-2.0, B2=-5.0)
    test.solve = make_solver(noise)
    cache.clear()
    ref = L.LiteralStokes(mesh, B1=-2.0, B2=-5.0)
    for k in range(steps):
        ref.flow_step(); test.flow_step()
        if k in (0, 1, 9, steps - 1):
            pr, pt = ref.p - ref.p.mean(), test.p - test.p.mean()
            print(f"noise {noise:g} step {k}: N={ref.N} rel u {np.linalg.norm(test.u - ref.u) / np.linalg.norm(ref.u):.2e}  rel p(mean-free) {np.linalg.norm(pt - pr) / np.linalg.norm(pr):.2e}", flush=True)
