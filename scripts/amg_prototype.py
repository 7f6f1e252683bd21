"""CPU prototype of the AMG-PCG (scipy) to compare cycle / coarsening variants by iteration count.
Not the product: a research tool for the next round.  python scripts/amg_prototype.py [n_theta n_r]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import fluidsim_b200 as fb
from oracle import restated as R

from oracle.amg_cpu import Hier, pcg      # the hierarchy moved to the oracle (CPU baseline of bench.py)

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
