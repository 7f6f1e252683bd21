"""Config 4 of BASELINE.json: StokesFood capture sweep, 64 squirmer (B1,B2) configurations x 4M tracers,
sharded over the GPUs with no data-path communication (one rank per GPU, round-robin over the configs).

    python scripts/bench_food_sweep.py [--steps 200] [--configs 64] [--grid 2257]
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_food_sweep.py

Every rank advances its configurations on mesh5.1 (nu=1, DT=0.01, code/StokesFood.py:38-42): flow step +
tracer step (interpolate, Euler, wrap, sticky capture) with the tracers resident on the device; the eaten
fractions are summed over the ranks at the end (the only collective).  Prints one JSON line (rank 0).
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib, parallel as par

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--configs", type=int, default=64)
ap.add_argument("--grid", type=int, default=2257)      # g x g tracer grid minus the body: 3.86M tracers
ap.add_argument("--batched", action="store_true",
                help="advance this rank's configurations TOGETHER (fs_stokes_step_batch / fs_tracer_step_batch): one launch "
                     "sequence per step for all of them instead of one per configuration")
ap.add_argument("--group", type=int, default=16, help="--batched: configurations per batch (memory: 24 B x tracers each)")
args = ap.parse_args()
rank, world, local, dist = par.init_distributed()
torch.cuda.set_device(local)
_lib.call("fs_set_device", local)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "mesh5_1_ops.npz"))
mine = par.shard_list(par.sweep_configs()[:args.configs], rank, world)
pts0 = fb.food_tracer_grid(args.grid)
P = len(pts0)
results = []
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
t0 = time.perf_counter()
if args.batched:
    for k0 in range(0, len(mine), args.group):
        grp = mine[k0:k0 + args.group]
        sw = fb.StokesSweep(g["nodes"], g["markers"], g["tris"], grp, DT=0.01, v=1.0, tracer_points=pts0, device_arrays=True)
        for _ in range(args.steps):
            eaten = sw.step_all()
        results += [(b1, b2, int(e) / P) for (b1, b2), e in zip(grp, eaten)]
        del sw
    mine = []
for (B1, B2) in mine:
    sim = fb.StokesSolver(g["nodes"], g["markers"], g["tris"], B1=B1, B2=B2, DT=0.01, v=1.0)
    u = torch.from_numpy(sim.u.copy()).cuda()
    pts = torch.from_numpy(pts0).cuda()
    status = torch.zeros(P, dtype=torch.int32, device="cuda")
    hint = torch.full((P,), -1, dtype=torch.int32, device="cuda")
    eaten = C.c_int64(0)
    for _ in range(args.steps):
        sim.step(u)
        _lib.call("fs_tracer_step", sim.mesh._h, _lib.ptr(pts), _lib.ptr(status), _lib.ptr(hint), P, _lib.ptr(u),
                  0.01, 1.0, 0.5, 0.5, 0.28, C.byref(eaten))
    results.append((B1, B2, eaten.value / P))
torch.cuda.synchronize()
dt = par.allreduce(time.perf_counter() - t0, "max", dist)
tot_eaten = par.allreduce(sum(r[2] for r in results), "sum", dist)
if rank == 0:
    n_cfg = min(args.configs, 64)
    print(json.dumps({"workload": f"StokesFood sweep: {n_cfg} (B1,B2) configs x {P} tracers, mesh5.1, {args.steps} steps each",
                      "mode": f"batched, {args.group} configurations per launch sequence" if args.batched else "one configuration at a time",
                      "n_gpus": world, "seconds": dt, "config_steps_per_s": n_cfg * args.steps / dt,
                      "tracer_updates_per_s": n_cfg * args.steps * P / dt,
                      "ms_per_config_step": 1e3 * dt * world / (n_cfg * args.steps),
                      "mean_eaten_fraction": tot_eaten / n_cfg, "rank0_first": results[:3]}))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
