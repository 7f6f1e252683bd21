"""Partitioned Stokes step across GPUs: correctness check and timing (torchrun, one rank per GPU; also runs
with a single process = one block).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dist_step.py --n-theta 512 --n-r 256 --steps 5 [--check] [--gather-rows 100000]

--check: rank 0 also advances the same mesh with the single-GPU StokesSolver and compares the gathered
velocity and pressure after the last step (relative L2).  Prints one JSON line (rank 0).
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib, parallel as par

ap = argparse.ArgumentParser()
ap.add_argument("--n-theta", type=int, default=512)
ap.add_argument("--n-r", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=0)
ap.add_argument("--gather-rows", type=int, default=100000)
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--check", action="store_true")
ap.add_argument("--profile-pcg", type=int, default=0, help="also time this many PCG iterations at a fixed count")
args = ap.parse_args()

rank, world, local, dist = par.init_distributed()
torch.cuda.set_device(local)
_lib.call("fs_set_device", local)
nodes, markers, tris = fb.square_with_hole(args.n_theta, args.n_r)
kw = dict(B1=-2.0, B2=-5.0, DT=0.05, v=0.1, rtol_pressure=args.rtol, rtol_visc=1e-12)
t0 = time.perf_counter()
ps = fb.PartitionedStokes(nodes, markers, tris, rank=rank, world=world, dist=dist, align=args.n_theta,
                          gather_rows=args.gather_rows, **kw)
t_setup = time.perf_counter() - t0
u = torch.from_numpy(ps.u.copy()).cuda()
iters = []
for _ in range(args.warmup):
    ps.step(u)
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
l0 = fb.launch_count()
t0 = time.perf_counter()
for _ in range(args.steps):
    st = ps.step(u)
    iters.append((st.iters_visc, st.iters_p1, st.iters_p2))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tmax = par.allreduce(dt, "max", dist)
out = {"n_gpus": world, "n_theta": args.n_theta, "n_r": args.n_r, "triangles": int(len(tris)), "steps": args.steps,
       "ms_per_step": 1e3 * tmax / max(args.steps, 1), "cg_iters_per_step": iters, "setup_s": t_setup,
       "levels_partitioned": ps.levels_partitioned, "n_own": ps.n_own, "n_halo_nodes": ps.n_halo_nodes,
       "n_own_dofs": ps.n_own_dofs, "n_halo_dofs": ps.n_halo_dofs, "launches_per_step": (fb.launch_count() - l0) / max(args.steps, 1)}
if args.profile_pcg:
    ps.trace()                                  # reset the event log
    out["us_per_pcg_iteration_fixed"] = ps.profile_pcg(args.profile_pcg)
    tr = ps.trace()
    if len(tr):
        np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"trace_rank{rank}.npy"), tr)
    out["debug_skip"] = os.environ.get("FS_DIST_DEBUG_SKIP", "0")
if args.check:
    ug = ps.gather(u)
    pg = ps.gather(ps.pressure()[0])
    if rank == 0:
        ref = fb.StokesSolver(nodes, markers, tris, precond=fb.PRECOND_AMG, **kw)
        ur = torch.from_numpy(ref.u.copy()).cuda()
        it1 = []
        for _ in range(args.warmup + args.steps):
            s1 = ref.step(ur)
            it1.append((s1.iters_visc, s1.iters_p1, s1.iters_p2))
        ur = ur.cpu().numpy()
        pr = ref.pressure()[0]
        out["check_rel_err_u_vs_1gpu"] = float(np.linalg.norm(ug - ur) / np.linalg.norm(ur))
        out["check_rel_err_p_vs_1gpu"] = float(np.linalg.norm(pg - pr) / np.linalg.norm(pr))
        out["single_gpu_iters"] = it1[args.warmup:]
if dist is not None:
    dist.barrier()
if rank == 0:
    print(json.dumps(out), flush=True)
if dist is not None:
    dist.destroy_process_group()
