"""CPU, world_size 2, gloo: the N>1 host logic -- sweep / tracer sharding with the final count
reduction, and the row-block partition + halo exchange of the partitioned CG (the SpMV itself is a
numpy stand-in here; the CUDA kernels are covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from fluidsim_b200 import parallel as par
    from fluidsim_b200.mesh import square_with_hole, find_boundary_pairs, filter_wall_pairs
    from oracle import restated as R
    r, w, _, d = par.init_distributed("gloo")
    assert (r, w) == (rank, world) and d is not None
    # --- sweeps: 64 configs, no overlap, union complete
    mine = par.shard_list(par.sweep_configs(), rank, world)
    got = [None] * world
    dist.all_gather_object(got, mine)
    flat = sorted(c for g in got for c in g)
    assert flat == sorted(par.sweep_configs()) and len(set(flat)) == 64
    # --- tracer ensemble: shard, count locally, allreduce
    pts = R.food_tracer_init(25)
    lo, hi = par.shard_range(len(pts), rank, world)
    local_eaten = int((np.hypot(pts[lo:hi, 0] - 0.5, pts[lo:hi, 1] - 0.5) <= 0.28).sum())
    total = par.allreduce(local_eaten, "sum", d)
    assert total == int((np.hypot(pts[:, 0] - 0.5, pts[:, 1] - 0.5) <= 0.28).sum())
    assert par.allreduce(float(rank), "max", d) == world - 1
    # --- partitioned SpMV with halo exchange == global SpMV
    nodes, markers, tris = square_with_hole(64, 48)
    pairs = filter_wall_pairs(nodes, find_boundary_pairs(nodes))
    ps = R.PressureSystem(nodes, tris, pairs)
    part = par.row_block_partition(ps.rowptr, ps.colidx, rank, world, align=32)
    lo, hi = part["lo"], part["hi"]
    a, b = part["nnz_range"]
    x = np.random.default_rng(0).standard_normal(ps.nd)        # same on both ranks
    xl = np.concatenate([x[lo:hi], np.zeros(len(part["halo_global"]))])
    # halo exchange: every rank publishes which global ids it needs; owners answer
    need = [None] * world
    dist.all_gather_object(need, part["halo_global"])
    for peer in range(world):
        if peer == rank:
            continue
        want = need[peer]
        sel = want[(want >= lo) & (want < hi)]
        dist.send(torch.from_numpy(x[sel].copy()), dst=peer) if rank < peer else None
        if rank > peer:
            buf = torch.empty(len(part["recv_slots"].get(peer, [])), dtype=torch.float64)
            dist.recv(buf, src=peer)
            xl[(hi - lo) + part["recv_slots"][peer]] = buf.numpy()
    for peer in range(world):                                   # second half of the pairwise exchange
        if peer == rank:
            continue
        want = need[peer]
        sel = want[(want >= lo) & (want < hi)]
        if rank > peer:
            dist.send(torch.from_numpy(x[sel].copy()), dst=peer)
        else:
            buf = torch.empty(len(part["recv_slots"].get(peer, [])), dtype=torch.float64)
            dist.recv(buf, src=peer)
            xl[(hi - lo) + part["recv_slots"][peer]] = buf.numpy()
    import scipy.sparse as sp
    Al = sp.csr_matrix((ps.vals[a:b], part["colidx"], part["rowptr"]), shape=(hi - lo, len(xl)))
    yl = Al @ xl
    yg = ps.K @ x
    assert np.array_equal(yl, yg[lo:hi]) or np.abs(yl - yg[lo:hi]).max() < 1e-12 * np.abs(yg).max()
    # dot product by allreduce
    assert abs(par.allreduce(float(yl @ xl[:hi - lo]), "sum", d) - float(yg @ x)) < 1e-9 * abs(float(yg @ x))
    # --- partitioned Stokes step, host logic: node blocks, sub-meshes, local index sets; the nodal operator
    # evaluated block by block (oracle arithmetic on the sub-mesh, halo velocities taken from the owner)
    # reproduces the whole-mesh result bit for bit
    from fluidsim_b200 import node_block_split, sub_mesh, local_index_sets, index_sets
    N = len(nodes)
    split = node_block_split(N, world, align=64)
    lo, hi = split[rank], split[rank + 1]
    ln, lm, lt, l2g, eids = sub_mesh(nodes, markers, tris, lo, hi)
    wall, inner, _, interior = index_sets(nodes, markers)
    lw, li, lint, lpairs = local_index_sets(lo, hi, wall, inner, interior, pairs)
    u_full = np.random.default_rng(7).standard_normal((N, 2))               # same on both ranks
    u_loc = np.zeros((len(l2g), 2))
    u_loc[:hi - lo] = u_full[lo:hi]
    need = [None] * world
    dist.all_gather_object(need, l2g[hi - lo:])                             # halo node ids of every rank
    answers = [None] * world
    mine = {peer: u_full[ids[(ids >= lo) & (ids < hi)]] for peer, ids in enumerate(need) if peer != rank}
    dist.all_gather_object(answers, mine)
    halo = l2g[hi - lo:]
    for peer in range(world):
        if peer == rank:
            continue
        sel = (halo >= split[peer]) & (halo < split[peer + 1])
        u_loc[(hi - lo) + np.nonzero(sel)[0]] = answers[peer][rank]
    div_blocks = [None] * world
    dist.all_gather_object(div_blocks, R.divergence(ln, lt, u_loc)[:hi - lo])
    assert np.array_equal(np.concatenate(div_blocks), R.divergence(nodes, tris, u_full))
    counts = [None] * world
    dist.all_gather_object(counts, (len(lw), len(li), len(lint), len(lpairs)))
    assert tuple(np.sum(counts, axis=0)) == (len(wall), len(inner), len(interior), len(pairs))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in procs)
    assert res == [(0, "ok"), (1, "ok")]


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from fluidsim_b200 import parallel as par
    for n, w in ((10, 3), (64, 8), (7, 8), (0, 2)):
        rs = [par.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    assert len(par.sweep_configs()) == 64 and par.sweep_configs()[0] == (-4.0, -5.0)
    rp = np.arange(0, 4001, 4, dtype=np.int32)
    ci = np.tile(np.arange(4, dtype=np.int32), 1000)
    p0 = par.row_block_partition(rp, ci, 0, 2, align=100)
    p1 = par.row_block_partition(rp, ci, 1, 2, align=100)
    assert p0["hi"] == p1["lo"] and p1["hi"] == 1000 and p0["lo"] == 0
    assert len(p0["halo_global"]) == 0 and list(p1["halo_global"]) == [0, 1, 2, 3]


# ---- the projected pressure guesses on a row-partitioned system (csrc/pstokes.cu keeps one basis block per rank) ----------
def _recycler_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from fluidsim_b200.mesh import square_with_hole, find_boundary_pairs, filter_wall_pairs
    from oracle import restated as R
    from oracle.cpu_step import Recycler
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes, markers, tris = square_with_hole(64, 48)
    ps = R.PressureSystem(nodes, tris, filter_wall_pairs(nodes, find_boundary_pairs(nodes)))
    K = ps.K.tocsr()
    n = K.shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world
    lu = spla.splu((K + sp.identity(n) * 1e-3).tocsc())          # SPD stand-in with a unique solution, same on both ranks
    A = (K + sp.identity(n) * 1e-3).tocsr()

    def allsum(v):
        t = torch.from_numpy(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
        dist.all_reduce(t)
        return t.numpy()

    class RowBlockRecycler(Recycler):
        """vectors hold the rows [lo, hi) only; dots are all-reduced, A d gathers the blocks of d first"""
        def _dots(self, v):
            return allsum(super()._dots(v))

        def _dot(self, a, b):
            return float(allsum(a @ b)[0])

        def _Kdot(self, d):
            parts = [None] * world
            dist.all_gather_object(parts, d)
            return (A @ np.concatenate(parts))[lo:hi]

    class _Op:                                                   # the whole-system twin's operator
        def dot(self, x):
            return A @ x

    part, whole = RowBlockRecycler(None, kmax=6, keep=3), Recycler(_Op(), kmax=6, keep=3)
    rng = np.random.default_rng(3)
    modes = rng.standard_normal((7, n))
    sizes, x_prev = [], None
    anorm = lambda e: float(np.sqrt(e @ (A @ e)))
    for step in range(14):
        t = 0.1 * step
        b = sum(np.cos(0.7 * j * t + j) * m for j, m in enumerate(modes))      # slowly varying right-hand side, 7 modes
        x = lu.solve(b)
        g_part, g_whole = part.guess(b[lo:hi]), whole.guess(b)
        if step:
            assert np.abs(g_part - g_whole[lo:hi]).max() <= 1e-10 * np.abs(g_whole).max()
            # the projection is the A-norm best approximation in a span that contains the previous solution
            assert anorm(g_whole - x) <= anorm(x_prev - x) * (1 + 1e-9)
        x_prev = x
        part.update(x[lo:hi])
        whole.update(x)
        sizes.append(part.k)
        assert part.k == whole.k and np.allclose(part.C, whole.C, rtol=0, atol=1e-9 * np.abs(whole.C).max())
    # the same coordinates on every rank (they decide the compression): exchanged and compared
    got = [None] * world
    dist.all_gather_object(got, (part.k, [list(c) for c in part.C]))
    assert all(g == got[0] for g in got)
    assert max(sizes) <= 5 and any(b < a for a, b in zip(sizes, sizes[1:]))      # at least one compression happened
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_world2_recycler_row_blocks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_recycler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5)[0] for _ in range(2)) == [0, 1]
