"""Tracer / dye throughput (configs 3 and 4 of BASELINE.json), single GPU, device-resident arrays.
    python scripts/bench_tracers.py            -> JSON lines
Config 3: mesh5.1 (N=331, T=522), pusher; 1M query points (g=1000 grid minus the hole) through the
          reference-semantics locator (k=10 nearest centroids) and through the food tracer step.
Config 4: one of the 64 sweep configurations with 4M tracers (g=2257 grid minus the hole).
The mesh is a few KB (L2/L1 resident): these kernels are latency/compute-bound, so points/s is the
figure, not an HBM fraction.
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "mesh5_1_ops.npz"))


def timed(fn, reps):
    fn(); _lib.call("fs_sync")
    ms = C.c_float()
    _lib.call("fs_timer_start")
    for _ in range(reps):
        fn()
    _lib.call("fs_timer_stop", C.byref(ms))
    return ms.value / reps * 1e-3


for name, gdens in (("config3_1M", 1000), ("config4_4M", 2257)):
    sim = fb.StokesFood(g["nodes"], g["markers"], g["tris"], B1=-2.0, B2=-5.0, DT=0.01, v=1.0,
                        tracer_points=fb.food_tracer_grid(gdens))
    for _ in range(20):
        sim.step()
    P = sim.num_tracers
    pts = torch.from_numpy(sim.tracer_points).cuda()
    status = torch.zeros(P, dtype=torch.int32, device="cuda")
    hint = torch.full((P,), -1, dtype=torch.int32, device="cuda")
    ids = torch.empty(P, dtype=torch.int32, device="cuda")
    u = torch.from_numpy(sim.u).cuda()
    t_loc = timed(lambda: _lib.call("fs_locate", sim.mesh._h, _lib.ptr(pts), P, _lib.ptr(ids)), 5)
    eaten = C.c_int64()
    step = lambda: _lib.call("fs_tracer_step", sim.mesh._h, _lib.ptr(pts), _lib.ptr(status), _lib.ptr(hint), P,
                             _lib.ptr(u), 0.01, 1.0, 0.5, 0.5, 0.28, C.byref(eaten))
    t_step = timed(step, 20)
    print(json.dumps({"workload": name, "mesh": "mesh5.1", "tracers": P,
                      "locate_knn10_Mpts_per_s": P / t_loc / 1e6, "tracer_step_Mpts_per_s": P / t_step / 1e6,
                      "ms_locate": 1e3 * t_loc, "ms_tracer_step": 1e3 * t_step, "eaten_after_21_steps": eaten.value,
                      "bytes_per_tracer_step": 16 * 2 + 4 * 2 + 4 * 2, "tracer_step_GBs": P * 48 / t_step / 1e9}))

# dye advection on the 4M-triangle mesh (the N mesh nodes are the "tracers" there)
nodes, markers, tris = fb.square_with_hole(2048, 1024)
m = fb.Mesh(nodes, tris, markers)
c = torch.from_numpy((nodes[:, 0] < 0.5).astype(np.float64)).cuda()
u = torch.from_numpy(0.1 * np.random.default_rng(0).standard_normal((m.N, 2))).cuda()
m.advect_dye(c, u, 0.05)
t = timed(lambda: m.advect_dye(c, u, 0.05), 5)
print(json.dumps({"workload": "advect_dye_4M_tri_mesh", "nodes": m.N, "ms": 1e3 * t, "Mnodes_per_s": m.N / t / 1e6}))
