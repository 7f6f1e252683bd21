#!/usr/bin/env python
"""bench.py -- Stokes steps/sec on the 4M-triangle synthetic mesh (BASELINE.json metric)
and the pressure-CG SpMV's achieved HBM GB/s against the measured roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the operator-split Stokes step (code/StokesColor.py:537-575):
2-RHS viscous CG, BCs, divergence, pressure CG, gradient update, BCs, second pressure
CG, interior update -- on the square-with-hole mesh n_theta=2048 x n_r=1024
(T=4 194 304, N=2 099 200), pusher squirmer B1=-2 B2=-5, nu=0.1, DT=0.05.

N>1 (torchrun, one rank per GPU): weak scaling of the PARTITIONED step (BASELINE config 5): the mesh grows
with N (N x 4M triangles: 2048x2048, 4096x2048, 4096x4096 = the 32M-triangle config at N=8), is cut into
N contiguous blocks of rings, and every rank advances its block -- halo values and dot products travel as
peer stores over NVLink from inside the kernels (csrc/dist.cuh).  value = N x (steps/s of the big mesh),
i.e. 4M-triangle-equivalent steps per second, so the metric keeps its meaning across N; the line also
carries check_rel_err_vs_1gpu (the partitioned solver against the single-GPU one on a 262k-triangle mesh).
--sweep keeps the old N>1 behaviour (independent (B1,B2) configurations, no collective).

--impl reference times the reference's CPU path.  The literal reference (dense numpy LU) cannot hold a
4M-triangle mesh (N x N doubles = 32 TB), so the arm runs the oracle's multi-threaded CPU statement of the
same step (oracle/cpu_step.py + oracle/cg_port.c, OpenMP, all host threads) with the SAME algorithm as the
GPU arm (AMG-preconditioned pressure CG): W + K complete, measured steps, nothing extrapolated.  A single
measured step with the Jacobi-CG (the closest sparse analogue of the reference's own solve) is reported next
to it as cpu_baseline_jacobi.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_THETA, N_R = 2048, 1024
PARAMS = dict(B1=-2.0, B2=-5.0, DT=0.05, v=0.1)
RTOL_P, RTOL_V = 1e-10, 1e-12
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "spmv_traffic.json")


def sweep_params(rank, world):
    """Config 4: the 64 (B1,B2) pairs of fluidsim_b200.parallel.sweep_configs(), sharded round-robin;
    rank 0 keeps the N=1 pusher so that its work is the N=1 workload."""
    if rank == 0:
        return PARAMS["B1"], PARAMS["B2"]
    from fluidsim_b200.parallel import shard_list, sweep_configs
    return shard_list(sweep_configs(), rank, world)[0]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken in [t_begin, t_end] (wall clock; all samples if the window holds none:
        the sampler is started before the warm-up so that it is already streaming when the timed region begins)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.1)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ CPU baseline
def load_hostmesh():
    """The numpy-only mesh helpers of the package, loaded by path: the CPU arm must not load libfluidsim.so."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fs_hostmesh", os.path.join(ROOT, "puc-fluidsimulation-project_b200", "hostmesh.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_steps(nodes, markers, tris, precond, warmup, steps, threads, budget_s=None):
    """W + K complete Stokes steps of the oracle's CPU statement (oracle/cpu_step.py), measured.  Stops early
    (after at least one timed step) once budget_s seconds of timed work are spent."""
    from oracle import cpu_step
    t0 = time.perf_counter()
    sim = cpu_step.CpuStokes(nodes, markers, tris, B1=PARAMS["B1"], B2=PARAMS["B2"], DT=PARAMS["DT"], v=PARAMS["v"],
                             precond=precond, rtol_pressure=RTOL_P, rtol_visc=RTOL_V, threads=threads)
    t_setup = time.perf_counter() - t0
    for _ in range(warmup):
        sim.step()
    iters, times = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        iters.append(tuple(int(v) for v in sim.step()))
        times.append(time.perf_counter() - t0)
        if budget_s is not None and sum(times) >= budget_s:
            break
    return {"s_per_step": float(np.mean(times)), "steps_timed": len(times), "warmup": warmup, "cg_iters_per_step": iters,
            "setup_s": t_setup}


def cpu_baseline_line(res, precond, threads, nt, nr):
    algo = ("AMG-preconditioned pressure CG started from the projection onto the previous solutions (same algorithm as "
            "the GPU arm: oracle/amg_cpu.py hierarchy, oracle/cpu_step.py Recycler)" if precond == "amg" else
            "Jacobi-preconditioned pressure CG (closest sparse analogue of the reference's dense solve), same projected guess")
    return {"value": 1.0 / res["s_per_step"], "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": f"{res['steps_timed']} complete measured Stokes steps after {res['warmup']} warm-up steps on the "
                      f"{2 * nt * nr}-triangle mesh, oracle/cpu_step.py + oracle/cg_port.c (OpenMP): {algo}; "
                      f"2-RHS-equivalent viscous Jacobi-CG, sparse div/grad; nothing extrapolated",
            "ms_per_step": 1e3 * res["s_per_step"], "cg_iters_per_step": res["cg_iters_per_step"], "setup_s": res["setup_s"]}


def run_reference(args):
    """--impl reference: CPU only (rank 0).  Does not import the product."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hm = load_hostmesh()
    threads = host_threads()                      # torchrun exports OMP_NUM_THREADS=1: set the count explicitly
    # weak scaling: the CPU arm advances ONE 4M-triangle share of the N x 4M-triangle mesh (the metric is in
    # 4M-triangle-equivalent steps/s, and the host does not grow with N)
    nodes, markers, tris = hm.square_with_hole(args.n_theta, args.n_r)
    res = cpu_steps(nodes, markers, tris, "amg", args.warmup, args.steps, threads)
    cpu = cpu_baseline_line(res, "amg", threads, args.n_theta, args.n_r)
    v = cpu["value"]
    out = {"impl": "reference", "metric": "stokes_steps_per_sec_4M_tri", "value": v, "unit": "steps/s",
           "n_gpus": args.gpus, "steps": res["steps_timed"], "warmup": args.warmup, "ms_per_step": 1e3 / v,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world=1, parallelism=f"CPU, {threads} OpenMP threads; one 4M-triangle share of the "
                                     f"N x 4M-triangle mesh of the GPU arm"),
           "cpu_baseline": cpu,
           "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if not args.no_extra:
        rj = cpu_steps(nodes, markers, tris, "jacobi", 0, 1, threads)
        out["cpu_baseline_jacobi"] = cpu_baseline_line(rj, "jacobi", threads, args.n_theta, args.n_r)
    out["product_loaded"] = any(m.startswith("fluidsim_b200") for m in sys.modules)      # must stay False: CPU arm only
    emit(out)


# N x 4M triangles at the N=1 mesh's element aspect ratio (n_theta = 2 n_r; n_theta a multiple of 8): the AMG-PCG
# iteration count depends on the aspect ratio (43-48 iterations at 2:1, 54-55 for n_theta = n_r on the CPU prototype),
# so the weak-scaling series keeps it fixed.  N = 2 and 8 miss N x 4 194 304 triangles by 0.02 %; `value` is scaled by
# the exact triangle count.
WEAK_SHAPES = {1: (2048, 1024), 2: (2896, 1448), 4: (4096, 2048), 8: (5792, 2896)}


def workload_config(args, world=1, **extra):
    nt, nr = (args.n_theta, args.n_r) if world == 1 else weak_shape(args, world)
    cfg = {"workload": f"stokes_step square-with-hole n_theta={nt} n_r={nr} "
                       f"(T={2 * nt * nr}, N={nt * (nr + 1)}), pusher B1=-2 B2=-5, nu=0.1, DT=0.05",
           "solver": f"pressure: fp64 CG (A, vectors, dot products, convergence test fp64) preconditioned by a smoothed-aggregation "
                     f"AMG V(1,1) cycle folded to two SELL-32 SpMVs per level; the cycle's operators are stored as packed 32-bit "
                     f"entries (fp16 value | 16-bit column offset), its two finest levels gather from fp32 mirrors on one GPU "
                     f"[--precond amg], or Jacobi persistent CG [--precond jacobi]; rtol_pressure={RTOL_P:g}; "
                     f"viscous: 2-RHS Jacobi CG rtol={RTOL_V:g}; every pressure solve starts from the projection of its solution "
                     f"onto the span of the previous ones (A-orthonormal basis of <= 12 vectors, csrc/recycle.cu) and runs to rtol",
           "l2": "inputs larger than L2 (per PCG iteration: A 190 MB fp64 SELL + V-cycle operators ~200 MB packed SELL "
                 "+ 5 vectors 84 MB > 126 MB per GPU), no flush needed",
           "steps_from": "t=0 (u=0 + squirmer BC); warm-up steps advance the same trajectory"}
    cfg.update(extra)
    return cfg


def weak_shape(args, world):
    if args.n_theta != N_THETA or args.n_r != N_R:          # explicit per-GPU share: grow radially
        return args.n_theta, args.n_r * world
    return WEAK_SHAPES[world]


# ------------------------------------------------------------------------------ GPU arm
def timed_steps(sim, u, steps, _lib, barrier):
    """K steps on the library stream between two CUDA events; returns (seconds, [(iters...)])."""
    iters = []
    barrier()
    _lib.call("fs_timer_start")
    for _ in range(steps):
        st = sim.step(u)
        iters.append((st.iters_visc, st.iters_p1, st.iters_p2))
    ms = C.c_float(0)
    _lib.call("fs_timer_stop", C.byref(ms))
    barrier()
    return ms.value / 1e3, iters


def read_profile(_lib):
    pms = np.zeros(3)
    ns, ni = C.c_int64(0), C.c_int64(0)
    _lib.call("fs_profile_read", _lib.ptr(pms), C.byref(ns), C.byref(ni))
    return pms, max(int(ns.value), 1)


def run_ours(args):
    import torch
    import fluidsim_b200 as fb
    from fluidsim_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    _lib.call("fs_set_device", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.call("fs_sync")

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    partitioned = world > 1 and not args.sweep
    precond = {"amg": fb.PRECOND_AMG, "jacobi": fb.PRECOND_JACOBI}[args.precond]
    kw = dict(DT=PARAMS["DT"], v=PARAMS["v"], rtol_pressure=RTOL_P, rtol_visc=RTOL_V)
    check = None
    if partitioned:
        # cross-check first (small mesh): the partitioned solver against the single-GPU one, same steps
        cn, cm, ct = fb.square_with_hole(512, 256)
        pc = fb.PartitionedStokes(cn, cm, ct, rank=rank, world=world, dist=dist, align=512, gather_rows=5000,
                                  B1=PARAMS["B1"], B2=PARAMS["B2"], **kw)
        uc = torch.from_numpy(pc.u.copy()).cuda()
        itc = []
        for _ in range(5):
            st = pc.step(uc)
            itc.append((st.iters_visc, st.iters_p1, st.iters_p2))
        ug = pc.gather(uc)
        if rank == 0:
            ref = fb.StokesSolver(cn, cm, ct, B1=PARAMS["B1"], B2=PARAMS["B2"], precond=fb.PRECOND_AMG, **kw)
            ur = torch.from_numpy(ref.u.copy()).cuda()
            it1 = []
            for _ in range(5):
                st = ref.step(ur)
                it1.append((st.iters_visc, st.iters_p1, st.iters_p2))
            ur = ur.cpu().numpy()
            check = {"value": float(np.linalg.norm(ug - ur) / np.linalg.norm(ur)), "mesh": "square-with-hole 512x256 (262144 triangles)",
                     "steps": 5, "cg_iters_partitioned": itc, "cg_iters_1gpu": it1, "levels_partitioned": pc.levels_partitioned}
            del ref, ur
        del pc, uc
        barrier()
        nt, nr = weak_shape(args, world)
        nodes, markers, tris = fb.square_with_hole(nt, nr)
        t0 = time.perf_counter()
        sim = fb.PartitionedStokes(nodes, markers, tris, rank=rank, world=world, dist=dist, align=nt,
                                   B1=PARAMS["B1"], B2=PARAMS["B2"], **kw)
        t_setup = time.perf_counter() - t0
        N = sim.n_own
        nd, nnz = sim.n_own_dofs, None
        u_dev = torch.from_numpy(sim.u.copy()).cuda()
        get_state, set_state = sim.get_state, sim.set_state
    else:
        nodes, markers, tris = fb.square_with_hole(args.n_theta, args.n_r)
        B1, B2 = sweep_params(rank, world) if world > 1 else (PARAMS["B1"], PARAMS["B2"])
        t0 = time.perf_counter()
        sim = fb.StokesSolver(nodes, markers, tris, B1=B1, B2=B2, precond=precond, **kw)
        t_setup = time.perf_counter() - t0
        N = sim.N
        _, kp, _ = sim.matrices()
        nd, nnz = kp.n, kp.nnz
        u_dev = torch.from_numpy(sim.u.copy()).cuda()
        get_state, set_state = sim.get_warm_state, sim.set_warm_state

    # ---- device-resident arm: W warm-up steps, then exactly K timed steps (no profiling events inside)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                      # streaming 100 ms samples from here on; summarised over the two timed regions
    for _ in range(args.warmup):
        sim.step(u_dev)
    barrier()
    u_snap = u_dev.clone()                  # loop state, so that the other arms repeat the same K steps
    warm_snap = get_state()
    t_clk0 = time.time()
    launches0 = fb.launch_count()
    _lib.call("fs_profile", 0)
    t_dev, iters = timed_steps(sim, u_dev, args.steps, _lib, barrier)
    launches = fb.launch_count() - launches0
    t_dev = max_over_ranks(t_dev)
    # steps/s in units of the N=1 mesh: N independent configurations count N; the partitioned mesh counts its
    # triangles / 4 194 304
    work = (2.0 * nt * nr) / (2.0 * N_THETA * N_R) if partitioned else float(world)
    value = work * args.steps / t_dev

    # ---- end-to-end arm: the same K steps through the host-buffer C-ABI call (pinned numpy view):
    # every step copies this rank's u host->device and device->host inside the timed region
    u_pin = torch.empty((N, 2), dtype=torch.float64).pin_memory()
    u_pin.copy_(u_snap.cpu())
    u_host = u_pin.numpy()
    set_state(warm_snap)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.step(u_host)
    barrier()
    e2e = work * args.steps / max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop(t_clk0, time.time()) if rank == 0 else None      # samples of the device arm and the end-to-end arm

    # ---- sampled pass (NOT part of value): per-kernel event timing for the roofline
    pms = samples = None
    top_ms, top_n, top_bytes = C.c_double(0), C.c_int64(0), C.c_double(0)
    us_iter_part = None
    if partitioned:
        us_iter_part = max_over_ranks(sim.profile_pcg(40))
    else:
        u_prof = u_snap.clone()
        set_state(warm_snap)
        _lib.call("fs_profile", 1)
        timed_steps(sim, u_prof, args.steps, _lib, barrier)
        pms, samples = read_profile(_lib)
        _lib.call("fs_profile_read_top", C.byref(top_ms), C.byref(top_n), C.byref(top_bytes))
        _lib.call("fs_profile", 0)
        del u_prof

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()       # everything below is rank-0 only and must not touch the group
        dist = None
    if rank != 0:
        return

    def barrier():                          # local from here on
        torch.cuda.synchronize()
        _lib.call("fs_sync")

    peak, peak_src = measured_peak()
    it_arr = np.array(iters, dtype=np.float64)
    out_extra = {}
    if partitioned:
        # per-rank algorithmic bytes of one AMG-PCG iteration at 4M triangles per GPU: A*p (SELL fp64)
        # 210 MB + folded V-cycle with packed operators and fp64 gathers ~300 MB + the two PCG vector kernels 151 MB
        it_bytes = 661e6
        roof = {"bound": "hbm", "kernel": "whole AMG-PCG iteration per rank (k_spmv_sell A*p + folded V-cycle SpMVs + k_ppcg_xr / k_ppcg_p, "
                                          "halo pushes and all-reduces inside the kernels), fixed-count run timed with CUDA events",
                "achieved": it_bytes / (us_iter_part * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": it_bytes / (us_iter_part * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": it_bytes, "us_per_launch": us_iter_part,
                "note": "per-rank figure; one 'launch' = one PCG iteration (11 kernels + 1 graph)"}
        out_extra["check_rel_err_vs_1gpu"] = check
        out_extra["setup_s"] = t_setup
    elif args.precond == "amg":
        # dominant kernel of this configuration: k_spmv_sell, the SELL-32 SpMV that runs the CG's A*p (fp64
        # values) and every large operator of the folded V-cycle (fp32 values).  Its longest launch is the
        # finest level's up-sweep z = [G | SP] [r; x_c] (+ fused r.z): timed with its own CUDA event pair on
        # every 8th PCG iteration of the sampled pass (that cycle runs outside its CUDA graph); the A*p
        # launch is bracketed by events in every iteration of the sampled pass.
        spmv_bytes = 12.0 * nnz + 20.0 * nd
        t_spmv = pms[0] / samples / 1e3
        ap_bytes = 12.0 * nnz + 16.25 * nd      # values + columns, slice pointers (8 B / 32 rows), p gathered, A*p written
        traffic = json.load(open(TRAFFIC_FILE)) if os.path.exists(TRAFFIC_FILE) else {}
        ap_line = {"kernel": "k_spmv_sell<plain,DOT,f64> (A*p of the pressure CG, fused p.Ap)",
                   "achieved": ap_bytes / t_spmv / 1e9, "frac": ap_bytes / t_spmv / 1e9 / peak,
                   "algorithmic_bytes_per_launch": ap_bytes, "us_per_launch": 1e6 * t_spmv, "sampled_launches": samples,
                   "traffic": traffic.get("sell_ap_dram_bytes_per_launch")}
        if top_n.value > 0:
            t_top = top_ms.value / top_n.value / 1e3
            roof = {"bound": "hbm", "kernel": "k_spmv_sell<split,DOT,packed> (finest up-sweep of the folded AMG V-cycle: "
                                              "z = [G | SP] [r; x_c], fused r.z; packed fp16 | 16-bit-offset entries, fp32 gathers)",
                    "achieved": top_bytes.value / t_top / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": top_bytes.value / t_top / 1e9 / peak, "traffic": traffic.get("sell_up0_dram_bytes_per_launch"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": top_bytes.value, "us_per_launch": 1e6 * t_top,
                    "sampled_launches": int(top_n.value), "frac_of_8TBps_spec": top_bytes.value / t_top / 8e12,
                    "same_kernel_A*p": ap_line}
        else:   # unfolded cycle (FS_AMG_FOLD=0): the CG's A*p is the largest launch
            roof = dict(ap_line, bound="hbm", peak=peak, unit="GB/s", peak_source=peak_src,
                        frac_of_8TBps_spec=ap_bytes / t_spmv / 8e12)
        us_it = {"A*p": 1e6 * t_spmv, "V-cycle": 1e3 * pms[1] / samples, "vector ops + dots": 1e3 * pms[2] / samples}
        us_it["total"] = sum(us_it.values())
        roof["us_per_pcg_iteration"] = us_it
        # the whole iteration against the same peak: A*p + the cycle's algorithmic bytes (library figure) + the two vector
        # kernels (k_pcg_xr: p, Ap, x, r read, x, r, r32 written; k_pcg_p: z, p read, p written)
        cyc = C.c_double(0)
        _lib.call("fs_precond_bytes", kp._h, C.byref(cyc))
        it_bytes = ap_bytes + cyc.value + 9.5 * 8.0 * nd
        roof["whole_pcg_iteration"] = {"algorithmic_bytes": it_bytes, "us": us_it["total"],
                                       "achieved": it_bytes / (us_it["total"] * 1e-6) / 1e9,
                                       "frac": it_bytes / (us_it["total"] * 1e-6) / 1e9 / peak,
                                       "bytes": {"A*p": ap_bytes, "V-cycle": cyc.value, "vector kernels": 9.5 * 8.0 * nd}}
        roof["note"] = "timed in a separate sampled pass over the same K steps (events inside the solver); value is timed without them"
    if args.precond == "amg" and not args.no_extra and world == 1:
        # the Jacobi-preconditioned persistent CG kernel on the same operator: the path's HBM-roofline kernel
        # (2000 fixed iterations, not part of `value`)
        cg_bytes = 12.0 * nnz + 92.0 * nd + 16.0 * nd      # Jacobi PCG iteration: + dinv read in passes B and C
        b = torch.randn(nd, dtype=torch.float64, device="cuda")
        x = torch.zeros_like(b)
        it_c, rr_c = C.c_int(0), C.c_double(0)
        _lib.call("fs_profile", 1)
        _lib.lib.fs_cg(kp._h, _lib.ptr(b), _lib.ptr(x), 1, 1e-300, 2000, fb.PRECOND_JACOBI, 1, C.byref(it_c), C.byref(rr_c))
        pj, sj = read_profile(_lib)
        _lib.call("fs_profile", 0)
        t_it = pj.sum() / sj / 1e3
        traffic_it = json.load(open(TRAFFIC_FILE)).get("dram_bytes_per_iteration") if os.path.exists(TRAFFIC_FILE) else None
        out_extra["roofline_persistent_cg"] = {
            "kernel": "k_cg_persistent (Jacobi PCG, one launch per solve: SpMV+dot | x,r update+dots | p update)",
            "bound": "hbm", "achieved": cg_bytes / t_it / 1e9, "peak": peak, "unit": "GB/s",
            "frac": cg_bytes / t_it / 1e9 / peak, "traffic": None if traffic_it is None else traffic_it * sj,
            "iterations_per_launch": sj, "algorithmic_bytes_per_iteration": cg_bytes, "us_per_iteration": 1e6 * t_it,
            "dram_bytes_per_iteration_ncu": traffic_it, "us_pass_A_B_C": [1e3 * v / sj for v in pj],
            "spmv_pass_GBs": spmv_bytes / (pj[0] / sj / 1e3) / 1e9,
            "spmv_pass_frac": spmv_bytes / (pj[0] / sj / 1e3) / 1e9 / peak,
            "note": "frac can exceed 1: the five CG vectors are kept L2-resident (evict_last), DRAM traffic is below the algorithmic bytes"}
        # same step with the Jacobi persistent CG (the closest analogue of the reference's own solve)
        simj = fb.StokesSolver(nodes, markers, tris, B1=B1, B2=B2, precond=fb.PRECOND_JACOBI, **kw)
        uj = torch.from_numpy(simj.u.copy()).cuda()
        for _ in range(args.warmup):          # same point of the trajectory as the timed AMG steps
            simj.step(uj)
        tj, itj = timed_steps(simj, uj, 2, _lib, barrier)
        out_extra["value_jacobi_pcg"] = {"value": 2 / tj, "unit": "steps/s", "cg_iters_per_step": itj,
                                         "note": "same algorithm as cpu_baseline_jacobi"}
        del simj, uj
    elif args.precond == "jacobi" and not partitioned:
        spmv_bytes = 12.0 * nnz + 20.0 * nd
        cg_bytes = 12.0 * nnz + 92.0 * nd + 16.0 * nd
        t_spmv = pms[0] / samples / 1e3
        t_it = pms.sum() / samples / 1e3
        n_launch = 2 * args.steps
        traffic_it = json.load(open(TRAFFIC_FILE)).get("dram_bytes_per_iteration") if os.path.exists(TRAFFIC_FILE) else None
        roof = {"bound": "hbm", "kernel": "k_cg_persistent (pressure PCG: SpMV+dot | x,r update+dots | p update, 3 grid barriers)",
                "achieved": cg_bytes / t_it / 1e9, "peak": peak, "unit": "GB/s", "frac": cg_bytes / t_it / 1e9 / peak,
                "traffic": None if traffic_it is None else traffic_it * samples / n_launch, "peak_source": peak_src,
                "launches_timed": n_launch, "iterations_per_launch": samples / n_launch,
                "algorithmic_bytes_per_launch": cg_bytes * samples / n_launch, "us_per_launch": 1e6 * t_it * samples / n_launch,
                "algorithmic_bytes_per_iteration": cg_bytes, "us_per_iteration": 1e6 * t_it,
                "dram_bytes_per_iteration_ncu": traffic_it, "us_pass_A_B_C": [1e3 * v / samples for v in pms],
                "spmv_pass_GBs": spmv_bytes / t_spmv / 1e9, "spmv_pass_frac": spmv_bytes / t_spmv / 1e9 / peak,
                "frac_of_8TBps_spec": cg_bytes / t_it / 8e12,
                "note": "frac can exceed 1: the five CG vectors are kept L2-resident (evict_last), DRAM traffic is below the algorithmic bytes"}
    cpu = None
    if world == 1 and not args.no_cpu:
        del sim, u_dev, u_snap
        torch.cuda.empty_cache()
        threads = host_threads()
        hm_nodes, hm_markers, hm_tris = nodes, markers, tris
        # bounded sample: complete CPU steps at the same point of the trajectory is not needed for a rate -- 2 warm-up
        # steps (the transient of the warm start), then measured steps until ~20 s of CPU work are spent
        res = cpu_steps(hm_nodes, hm_markers, hm_tris, "amg", 2, 12, threads, budget_s=20.0)
        cpu = cpu_baseline_line(res, "amg", threads, args.n_theta, args.n_r)
    if world == 1:
        par = "single GPU"
    elif partitioned:
        par = (f"mesh cut into {world} contiguous blocks of rings, one per GPU: row-block partitioned operators (A_visc, Z^T K Z, AMG "
               f"levels above 100k rows; smaller levels replicated), halo values pushed over NVLink peer memory from inside the "
               f"producing kernels, dot products all-reduced inside the PCG vector kernels (deterministic, rank order); one reduction "
               f"wave + 7 halo exchanges per PCG iteration; setup replicated")
    else:
        par = f"{world} independent squirmer (B1,B2) configs, one per GPU, no collective"
    out = {"metric": "stokes_steps_per_sec_4M_tri", "value": value, "unit": "steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world=world if partitioned else 1, precond=args.precond, cg_iters_per_step=iters,
                                     parallelism=par,
                                     value_note=(f"value = steps/s of the partitioned mesh x its triangle count / 4 194 304 "
                                                             f"(= {work:.4f}; 4M-triangle-equivalent steps/s)"
                                                 if partitioned else "value = steps/s")),
           "roofline": roof, "cpu_baseline": cpu,
           "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": 16 * N * (world if partitioned else 1),
                   "d2h_bytes_per_step": 16 * N * (world if partitioned else 1)},
           "gpu_launches": int(launches), "clocks": clk}
    out.update(out_extra)
    emit(out)


_REAL_STDOUT = None


def quiet_stdout():
    """Send fd 1 to stderr while the benchmark runs: NCCL / the driver write banners such as
    'NCCL version ...' straight to fd 1, and stdout must carry exactly one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-theta", dest="n_theta", type=int, default=N_THETA)
    ap.add_argument("--n-r", dest="n_r", type=int, default=N_R)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the Jacobi persistent-CG legs (roofline_persistent_cg, value_jacobi_pcg); profiling runs")
    ap.add_argument("--sweep", action="store_true",
                    help="N>1: independent (B1,B2) configurations, one per GPU (config 4), instead of the partitioned step")
    ap.add_argument("--precond", default="amg", choices=["amg", "jacobi"],
                    help="pressure-CG preconditioner of the timed steps (default: amg, the fastest)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
