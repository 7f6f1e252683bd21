"""Partitioned pressure CG across GPUs: correctness check and benchmark (torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dist_cg.py --n-theta 4096 --n-r 4096 [--check] [--maxit K]

Builds the periodic-merged SPD pressure operator of the square-with-hole mesh (every rank
assembles it on its own GPU, then keeps its row block), solves K q = rhs with the
partitioned persistent CG and prints one JSON line (rank 0).  --check also solves on one
GPU with the single-GPU kernel and compares (small meshes).
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib, parallel as par
from fluidsim_b200.distributed import PartitionedCG

ap = argparse.ArgumentParser()
ap.add_argument("--n-theta", type=int, default=2048)
ap.add_argument("--n-r", type=int, default=1024)
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--maxit", type=int, default=200000)
ap.add_argument("--fixed-iters", type=int, default=0, help="run exactly this many iterations (timing)")
ap.add_argument("--check", action="store_true")
args = ap.parse_args()

rank, world, local, dist = par.init_distributed()
torch.cuda.set_device(local)
_lib.call("fs_set_device", local)
nodes, markers, tris = fb.square_with_hole(args.n_theta, args.n_r)
sim = fb.StokesSolver(nodes, markers, tris, B1=-2.0, B2=-5.0)
_, kp, dof = sim.matrices()
rowptr, colidx, vals = kp.arrays()
nd = kp.n
# right-hand side of the first pressure solve: -(1/DT) M div(u*) with u* = u0 (BC only), merged
div = sim.mesh.divergence(sim.u)
rhs = np.bincount(dof, weights=sim.M_lumped_diag * (-(1.0 / sim.DT) * div), minlength=nd)
pc = PartitionedCG(rowptr, colidx, vals, rank, world, dist)
b_own = torch.from_numpy(rhs[pc.lo:pc.hi].copy()).cuda()
maxit = args.fixed_iters or args.maxit
rtol = 1e-300 if args.fixed_iters else args.rtol
# warm-up solve (few iterations), then the timed one
try:
    pc.solve(b_own, rtol=1e-300, maxit=20, project_mean=True)
except fb.FluidsimError as e:
    if e.code != -4:          # -4 = hit maxit, expected for the warm-up
        raise
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
try:
    x, it, rr = pc.solve(b_own, rtol=rtol, maxit=maxit, project_mean=True)
    conv = True
except fb.FluidsimError as e:
    if e.code != -4:
        raise
    x, it, rr, conv = None, maxit, float("nan"), False
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tmax = par.allreduce(dt, "max", dist)
ns = pc.last_ns.copy()
nnz = len(colidx)
out = {"workload": f"partitioned pressure CG, square-with-hole n_theta={args.n_theta} n_r={args.n_r} "
                   f"(T={2 * args.n_theta * args.n_r}, merged dofs {nd}, nnz {nnz})",
       "n_gpus": world, "iterations": it, "converged": conv, "relres": rr, "seconds": tmax,
       "us_per_iteration": 1e6 * tmax / max(it, 1), "iterations_per_s": it / tmax,
       "algorithmic_bytes_per_iteration": 12.0 * nnz + 108.0 * nd,
       "aggregate_GBs": (12.0 * nnz + 108.0 * nd) * it / tmax / 1e9,
       "rank0_us_pass_A_B_C": (ns / max(it, 1) / 1e3).tolist(), "n_own_rank0": pc.n_own, "n_halo_rank0": pc.n_halo,
       "n_send_rank0": pc.n_send}
if args.check and conv:
    xs, its, _ = kp.cg(rhs, rtol=args.rtol, project_mean=True)
    err = float(np.linalg.norm(x.cpu().numpy() - xs[pc.lo:pc.hi]) / np.linalg.norm(xs))
    out["check_rel_err_vs_single_gpu"] = par.allreduce(err, "max", dist)
    out["single_gpu_iterations"] = its
if rank == 0:
    print(json.dumps(out))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
