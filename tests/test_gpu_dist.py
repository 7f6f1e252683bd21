"""GPU: the partitioned persistent CG (fs_dist_*).  world=1 exercises the whole peer-memory
protocol against the rank's own mailbox on a single GPU; the 2-rank test needs two GPUs and
is skipped otherwise (run it with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import fluidsim_b200 as fb
from fluidsim_b200.distributed import PartitionedCG
from oracle import restated as R
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_partitioned_cg_world1_matches_oracle():
    nodes, markers, tris = fb.square_with_hole(256, 96)
    pairs = fb.filter_wall_pairs(nodes, fb.find_boundary_pairs(nodes))
    ps = R.PressureSystem(nodes, tris, pairs)
    b = np.random.default_rng(2).standard_normal(len(nodes))
    want = ps.solve(b)                                   # direct solve, mean-free
    rhs = np.bincount(ps.dof, weights=ps.M * b, minlength=ps.nd)
    pc = PartitionedCG(ps.rowptr, ps.colidx, ps.vals)
    x, it, rr = pc.solve(rhs, rtol=1e-12, project_mean=True)
    assert it > 10 and rr <= 1e-12
    assert np.linalg.norm(x[ps.dof] - want) <= 1e-9 * np.linalg.norm(want)
    # a second solve on the same handle (epochs keep counting) gives the same answer
    x2, it2, _ = pc.solve(rhs, rtol=1e-12, project_mean=True)
    assert it2 == it and np.array_equal(x, x2)
    # same iterates as the single-GPU persistent kernel
    A = fb.CsrMatrix.from_arrays(ps.rowptr, ps.colidx, ps.vals)
    xs, its, _ = A.cg(rhs, rtol=1e-12, project_mean=True)
    assert abs(its - it) <= 2 and np.abs(xs - x).max() <= 1e-9 * np.abs(xs).max()


def test_partitioned_cg_two_ranks():
    if fb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_cg.py"), "--n-theta", "512",
           "--n-r", "256", "--rtol", "1e-11", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["converged"] and out["n_gpus"] == 2
    assert out["check_rel_err_vs_single_gpu"] <= 1e-9
    assert abs(out["iterations"] - out["single_gpu_iterations"]) <= 3


# ---- the partitioned Stokes step (fs_pstokes_*) -------------------------------------------------------
def _run_dist_step(nproc, extra, port):
    script = os.path.join(ROOT, "scripts", "dist_step.py")
    if nproc == 1:
        cmd = [sys.executable, script] + extra
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
               "127.0.0.1", "--master-port", str(port), script] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_partitioned_step_one_block_matches_single_gpu():
    """world = 1: the partitioned code path (row-block extraction, arena vectors, in-kernel reductions, partitioned
    AMG levels + replicated tail) against fs_stokes_step on the same mesh."""
    out = _run_dist_step(1, ["--n-theta", "256", "--n-r", "128", "--steps", "6", "--check", "--gather-rows", "3000"], 0)
    assert out["levels_partitioned"] >= 2
    assert out["check_rel_err_u_vs_1gpu"] <= 1e-9 and out["check_rel_err_p_vs_1gpu"] <= 1e-7
    for a, b in zip(out["cg_iters_per_step"], out["single_gpu_iters"]):
        assert a[0] == b[0] and abs(a[1] - b[1]) <= 1 and abs(a[2] - b[2]) <= 1


def test_partitioned_step_two_ranks():
    if fb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run_dist_step(2, ["--n-theta", "512", "--n-r", "256", "--steps", "6", "--check", "--gather-rows", "5000"], 29541)
    assert out["n_gpus"] == 2 and out["levels_partitioned"] >= 2 and out["n_halo_nodes"] > 0
    assert out["check_rel_err_u_vs_1gpu"] <= 1e-9 and out["check_rel_err_p_vs_1gpu"] <= 1e-7
    for a, b in zip(out["cg_iters_per_step"], out["single_gpu_iters"]):
        assert a[0] == b[0] and abs(a[1] - b[1]) <= 1 and abs(a[2] - b[2]) <= 1
