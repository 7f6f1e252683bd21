"""Configs 1-3 of BASELINE.json on the shipped meshes (stored in tests/golden/*.npz): GPU time through
the public API next to the CPU oracle (restated tier; the literal reference cannot run on the GPU box).
    python scripts/bench_small_configs.py      -> JSON lines
These meshes are a few hundred nodes: everything is launch/latency bound, the numbers document the
per-step cost of the drop-in path, not a roofline.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidsim_b200 as fb
from oracle import restated as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))


def timeit(fn, reps=1):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


# config 1: Poisson on mesh2.1 (code/poisson.py)
g = G("mesh2_1_poisson")
t_setup = timeit(lambda: fb.PoissonProblem(g["nodes32"], g["markers"], g["tris"]))
pb = fb.PoissonProblem(g["nodes32"], g["markers"], g["tris"])
t_solve = timeit(pb.solve, 5)
A = np.zeros((pb.N, pb.N)); rp, ci, v = pb.A.arrays()
for i in range(pb.N): A[i, ci[rp[i]:rp[i + 1]]] = v[rp[i]:rp[i + 1]]
t_cpu = timeit(lambda: np.linalg.solve(A, pb.b), 5)
print(json.dumps({"config": "1: poisson mesh2.1 (N=277)", "gpu_setup_ms": 1e3 * t_setup, "gpu_solve_ms": 1e3 * t_solve,
                  "bicgstab_iters": pb.iters, "cpu_dense_solve_ms (np.linalg.solve, as the reference)": 1e3 * t_cpu,
                  "rel_err_vs_golden": float(np.linalg.norm(pb.solve() - g["f"]) / np.linalg.norm(g["f"]))}))

# config 2: heat on mesh_fine.1, 1000 steps (code/heatEq.py)
g = G("mesh_fine_1_poisson")
hp = fb.HeatProblem(g["nodes32"], g["markers"], g["tris"], DT=0.02)
hp.step()
t0 = time.perf_counter()
for _ in range(1000):
    hp.step()
t_heat = time.perf_counter() - t0
rp, ci, v = hp.A.arrays()
Ah = np.zeros((hp.N, hp.N))
for i in range(hp.N): Ah[i, ci[rp[i]:rp[i + 1]]] = v[rp[i]:rp[i + 1]]
u = np.zeros(hp.N)
t_cpu = timeit(lambda: np.linalg.solve(Ah, u + 1.0), 3)
print(json.dumps({"config": "2: heat mesh_fine.1 (N=1067), 1000 steps", "gpu_ms_per_step": t_heat, "gpu_total_s": t_heat,
                  "bicgstab_iters_last": hp.iters, "cpu_dense_solve_ms_per_step (reference: one np.linalg.solve per step)": 1e3 * t_cpu}))

# config 3: StokesColor on mesh5.1, pusher, 100 steps (+ dye), and 1M tracer queries per step
g = G("mesh5_1_ops")
sim = fb.StokesColor(g["nodes"], g["markers"], g["tris"], B1=-2.0, B2=-5.0, DT=0.05, v=0.1)
sim.step_all()
t0 = time.perf_counter()
for _ in range(100):
    sim.step_all()
t_gpu = (time.perf_counter() - t0) / 100
o = R.RestatedStokes(g["nodes"], g["markers"], g["tris"], B1=-2.0, B2=-5.0, DT=0.05, v=0.1)
loc = R.Locator(g["nodes"], g["tris"]); c = (g["nodes"][:, 0] < 0.5).astype(float)
def cpu_step():
    o.flow_step(); R.advect_semilagrange(c, o.u, 0.05, g["nodes"], g["tris"], loc)
t_cpu = timeit(cpu_step, 20)
print(json.dumps({"config": "3: StokesColor mesh5.1 (N=331), flow step + dye + mixing index", "gpu_ms_per_step": 1e3 * t_gpu,
                  "cpu_oracle_ms_per_step (restated, sparse direct solves)": 1e3 * t_cpu,
                  "reference_literal_ms_per_step (SURVEY 3.1, measured in the authoring container)": 77.0,
                  "pressure_cg_iters": [sim.stats.iters_p1, sim.stats.iters_p2]}))
