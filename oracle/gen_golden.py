"""Fixture generator: runs the LITERAL reference functions (oracle/literal.py, needs
/root/reference, authoring container only) on the three shipped meshes, checks
the restated oracle (oracle/restated.py) against them, and writes the golden
vectors to tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY (see oracle/restated.py header).

    python -m oracle.gen_golden            # regenerates every fixture (~3 min)

The asserts in here ARE the pinning of the restated oracle: K, M, div, pairs,
index sets, A_visc, locator ids, dye interpolation must equal the literal
reference bit for bit; grad / BC values / mixing index to rounding (1e-15).
For the two unpinned boundaries (pressure solve, food interpolation) the
measured gap to the literal code is stored in the fixture ("gap_*" keys).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy.sparse as sp

from . import literal as L
from . import restated as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
MESHES = ["mesh2.1", "mesh5.1", "mesh_fine.1"]
SNAP = [0, 1, 9, 49, 99]


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def gen_operators(mesh):
    ns = L.load_functions("StokesColor.py")
    npth, epth = L.mesh_paths(mesh)
    with L.quiet():
        nodes, markers = ns["readNode"](npth)
        tris = ns["readEle"](epth)
    n2, m2 = R.read_node(npth)
    assert np.array_equal(nodes, n2) and np.array_equal(markers, m2)
    assert np.array_equal(tris, R.read_ele(epth))
    N = nodes.shape[0]
    with L.quiet():
        K, _ = ns["buildStiffnessMatrix"](nodes, tris, g_source=0.0)
        M = ns["buildLumpedMassMatrix"](nodes, tris)
        pairs_all = [(int(a), int(b)) for a, b in ns["find_boundary_pairs"](nodes, L=1.0)]
    rowptr, colidx, scatter = R.csr_pattern(N, tris)
    kv = R.assemble_stiffness(nodes, tris, rowptr, colidx, scatter)
    Ks = sp.csr_matrix((kv, colidx, rowptr), shape=(N, N))
    assert np.array_equal(Ks.toarray(), K), "K not bit-exact"
    # structural pattern covers every literal nonzero
    nzr, nzc = np.nonzero(K)
    assert set(zip(nzr.tolist(), nzc.tolist())) <= set(
        zip(np.repeat(np.arange(N), np.diff(rowptr)).tolist(), colidx.tolist()))
    assert np.array_equal(R.lumped_mass(nodes, tris), M), "M not bit-exact"
    assert R.find_boundary_pairs(nodes) == pairs_all
    pairs = R.filter_wall_pairs(nodes, pairs_all)
    wall, inner_b, dirichlet, interior = R.index_sets(nodes, markers)

    out = dict(nodes=nodes, markers=markers, tris=tris, rowptr=rowptr, colidx=colidx,
               scatter=scatter, K=kv, M=M, pairs_all=np.array(pairs_all, dtype=np.int32),
               pairs=np.array(pairs, dtype=np.int32).reshape(-1, 2), wall=wall.astype(np.int32),
               inner_b=inner_b.astype(np.int32), interior=interior.astype(np.int32))

    # KAT fields (scripts/stokes_report.py:388-431 Tests A, B; scripts/final_test.py:124-156)
    rng = np.random.default_rng(0)
    fields_u = {
        "rand": rng.standard_normal((N, 2)),
        "katB": np.stack([2 * nodes[:, 0], 3 * nodes[:, 1]], axis=1),     # div = 5
        "final_test": nodes.copy(),                                        # div = 2
    }
    fields_p = {"rand": rng.standard_normal(N), "katA": 2 * nodes[:, 0] + 3 * nodes[:, 1]}
    for k, u in fields_u.items():
        d = ns["calculate_divergence"](nodes, tris, u)
        assert np.array_equal(d, R.divergence(nodes, tris, u)), "div not bit-exact"
        out["u_" + k] = u
        out["div_" + k] = d
    for k, p in fields_p.items():
        gx, gy = ns["calculate_gradiant"](nodes, tris, p)
        rx, ry = R.gradient(nodes, tris, p)
        assert np.allclose(gx, rx, rtol=0, atol=1e-13 * np.abs(gx).max())
        assert np.allclose(gy, ry, rtol=0, atol=1e-13 * np.abs(gy).max())
        out["p_" + k] = p
        out["gx_" + k] = gx
        out["gy_" + k] = gy

    # squirmer BC values (code/StokesColor.py:405-427)
    ns.update(dict(wall_node_indices=wall, inner_boundary_indices=inner_b, nodes_coords=nodes,
                   pairs=pairs, OUTER_BOUNDARY_VALUE=[0.0, 0.0]))
    bcs = []
    for B1, B2 in [(-2.0, 0.0), (-2.0, -5.0), (-2.0, 5.0)]:
        ns["B1"], ns["B2"] = B1, B2
        u = rng.standard_normal((N, 2))
        u0 = u.copy()
        ns["makePerBCU"](u)
        ns["makeDirBCU"](u)
        ur = u0.copy()
        R.make_per_bcu(ur, pairs)
        R.make_dir_bcu(ur, nodes, wall, inner_b, B1, B2)
        assert np.allclose(u, ur, rtol=0, atol=1e-15)
        bcs.append((B1, B2, u0, u))
    out["bc_B"] = np.array([(b[0], b[1]) for b in bcs])
    out["bc_in"] = np.stack([b[2] for b in bcs])
    out["bc_out"] = np.stack([b[3] for b in bcs])

    # A_visc on the pattern for both scripts' constants
    for tag, DT, v in [("color", 0.05, 0.1), ("food", 0.01, 1.0)]:
        A = np.eye(N) + DT * v * K
        A[dirichlet, :] = 0.0
        A[:, dirichlet] = 0.0
        A[dirichlet, dirichlet] = 1.0
        av = R.viscous_matrix(N, rowptr, colidx, kv, dirichlet, DT, v)
        assert np.array_equal(sp.csr_matrix((av, colidx, rowptr), shape=(N, N)).toarray(), A)
        out["avisc_" + tag] = av

    # locator ids (code/StokesColor.py:314-345): seeded cloud + the mesh nodes
    loc = ns["PointLocator"](nodes, tris)
    pts = np.concatenate([rng.random((20000, 2)), nodes])
    ids = np.array([-1 if (r := loc.find(x, y)) is None else int(r) for x, y in pts], dtype=np.int32)
    rl = R.Locator(nodes, tris)
    assert np.array_equal(ids, rl.find(pts)), "locator ids differ"
    out["loc_pts"] = pts
    out["loc_ids"] = ids

    # one dye advection + mixing index with a seeded velocity (code/StokesColor.py:347-403)
    ns.update(dict(N=N, triangles=tris, point_locator=loc))
    c = (nodes[:, 0] < 0.5).astype(np.float64)
    u = 0.5 * rng.standard_normal((N, 2))
    c_l = c.copy()
    ns["advect_semilagrange"](c_l, u, 0.05)
    c_r = c.copy()
    R.advect_semilagrange(c_r, u, 0.05, nodes, tris, rl)
    assert np.array_equal(c_l, c_r), "dye advection not bit-exact"
    mask = np.where(markers == 0)[0]
    mi_l = ns["mixing_index"](c_l, M, mask=mask)
    mi_r = R.mixing_index(c_r, M, mask=mask)
    assert np.allclose(mi_l, mi_r, rtol=1e-14)
    out.update(dye_u=u, dye_c0=c, dye_c1=c_l, dye_mix=np.array(mi_l))
    np.savez_compressed(os.path.join(OUT, mesh.replace(".", "_") + "_ops.npz"), **out)
    print(f"[{mesh}] operators: N={N} T={len(tris)} nnz={len(colidx)} pairs={len(pairs)} OK")


def gen_trajectory(mesh, tag, B1, B2, DT, v, steps=100, dye=True, food=False):
    t0 = time.time()
    lit = L.LiteralStokes(mesh, B1=B1, B2=B2, DT=DT, v=v)
    res = R.RestatedStokes(lit.nodes, lit.markers, lit.tris, B1=B1, B2=B2, DT=DT, v=v)
    assert lit.pairs == res.pairs
    nodes, tris = lit.nodes, lit.tris
    loc = R.Locator(nodes, tris)
    c_res = lit.c.copy()
    mask = lit.inner_mask
    _, _, var0 = R.mixing_index(c_res, res.M, mask)
    out = dict(B1=B1, B2=B2, DT=DT, v=v, steps=steps, snap=np.array([s for s in SNAP if s < steps]))
    if food:
        pts = R.food_tracer_init()
        status = np.zeros(len(pts), dtype=np.int64)
        out["food_pts0"] = pts.copy()
        eaten = []
    prog_l, prog_r, gap_u, gap_p = [], [], [], []
    for s in range(steps):
        lit.flow_step()
        res.flow_step()
        if dye:
            prog_l.append(lit.dye_step())
            R.advect_semilagrange(c_res, res.u, DT, nodes, tris, loc)
            _, _, var = R.mixing_index(c_res, res.M, mask)
            prog_r.append(1.0 - var / (var0 + 1e-16))
        if food:
            eaten.append(R.food_tracer_step(nodes, tris, res.u, pts, status, DT))
        pl = lit.p - lit.p.mean()
        gap_u.append(_rel(res.u, lit.u))
        gap_p.append(_rel(res.p - res.p.mean(), pl))
        if s in SNAP:
            out[f"lit_u_{s}"] = lit.u.copy()
            out[f"lit_p_{s}"] = pl
            out[f"res_u_{s}"] = res.u.copy()
            out[f"res_p_{s}"] = res.p.copy()
            out[f"res_p2_{s}"] = res.p2.copy()
            if dye:
                out[f"lit_c_{s}"] = lit.c.copy()
                out[f"res_c_{s}"] = c_res.copy()
            if food:
                out[f"food_pts_{s}"] = pts.copy()
                out[f"food_status_{s}"] = status.copy()
    out["gap_u"] = np.array(gap_u)
    out["gap_p"] = np.array(gap_p)
    if dye:
        out["lit_progress"] = np.array(prog_l)
        out["res_progress"] = np.array(prog_r)
    if food:
        out["food_eaten"] = np.array(eaten)
    np.savez_compressed(os.path.join(OUT, f"{mesh.replace('.', '_')}_traj_{tag}.npz"), **out)
    msg = f"[{mesh}] trajectory {tag}: {steps} steps, gap(u) {gap_u[-1]:.2e} gap(p) {gap_p[-1]:.2e}"
    if dye:
        msg += f", |progress lit-res| max {np.abs(np.array(prog_l) - np.array(prog_r)).max():.2e}"
    if food:
        msg += f", eaten {eaten[-1]}/{len(pts)}"
    print(msg, f"({time.time() - t0:.0f}s)")


def gen_poisson_heat(mesh, heat_steps=1000):      # BASELINE config 2: 1000 time steps
    """Literal poisson.py / heatEq.py module bodies re-enacted with the reference
    functions (code/poisson.py:216-285, code/heatEq.py:219-333), fp32 coordinates."""
    ns = L.load_functions("poisson.py")
    nh = L.load_functions("heatEq.py")
    npth, epth = L.mesh_paths(mesh)
    H, tol = 1.0, 1e-6
    with L.quiet():
        nodes, markers = ns["readNode"](npth)          # float32
        tris = ns["readEle"](epth)
        pairs = ns["find_boundary_pairs"](nodes, L=1.0)
    assert nodes.dtype == np.float32
    N = nodes.shape[0]

    def g_source_fun(x, y):                             # code/poisson.py:234-235
        return 50 * np.sin(3 * y)

    A, b = ns["buildFemSystem"](nodes, tris, g_source=g_source_fun)
    A0, b0 = A.copy(), b.copy()
    filt = []
    for m, s in pairs:
        my = nodes[m, 1]
        if not (np.abs(my - 0.0) < tol or np.abs(my - H) < tol):
            filt.append((m, s))
    ns["apply_periodic_bc"](A, b, filt)
    for i in range(N):                                   # code/poisson.py:258-278
        y = nodes[i, 1]
        is_wall = np.abs(y - 0.0) < tol or np.abs(y - H) < tol
        is_inner = markers[i] == 2
        if is_wall or is_inner:
            A[i, :] = 0.0
            A[i, i] = 1.0
            b[i] = 0.0 if is_inner else 1.0
    f = np.linalg.solve(A, b)

    # restated check
    cx, cy = R.centroids(nodes, tris)
    rp, ci, vals, rb = R.fem_system(nodes, tris, g_centroid=g_source_fun(cx, cy))
    assert np.array_equal(sp.csr_matrix((vals, ci, rp), shape=(N, N)).toarray(), A0), "fem A not bit-exact"
    assert np.allclose(rb, b0, rtol=1e-6, atol=1e-9)
    gap_b = _rel(rb, b0)
    Ar, br, pa, pf = R.poisson_system(nodes, markers, tris, g_source_fun(cx, cy))
    assert [(int(a), int(c)) for a, c in pairs] == pa
    assert np.array_equal(Ar.toarray(), A), "poisson matrix not bit-exact"
    As = sp.csr_matrix(A)

    # heat (code/heatEq.py:304-333)
    nh.update(dict(N=N, nodes_boundary_markers=markers, nodes_coords=nodes, tol=tol, H=H,
                   INNER_BOUNDARY_MARKER=2, INNER_BOUNDARY_VALUE=0.0, OUTER_BOUNDARY_VALUE=1.0,
                   pairs=pairs))
    DT = 0.02
    Ah = np.eye(N) + DT * A
    u = np.zeros(N)
    u = nh["reapply_periodic_u"](u)
    u = nh["reapply_dirchlect_u"](u)
    ur = R.heat_reapply(np.zeros(N), nodes, markers, pa)
    assert np.array_equal(u, ur)
    heat = {"heat_u_init": u.copy()}
    snaps = sorted({0, 1, 9, min(99, heat_steps - 1), heat_steps - 1})
    for n in range(heat_steps):
        rhs = u + DT * b * 0
        u = np.linalg.solve(Ah, rhs)
        u = nh["reapply_periodic_u"](u)
        u = nh["reapply_dirchlect_u"](u)
        if n in snaps:
            heat[f"heat_u_{n}"] = u.copy()
    np.savez_compressed(
        os.path.join(OUT, mesh.replace(".", "_") + "_poisson.npz"),
        nodes32=nodes, markers=markers, tris=tris, pairs_all=np.array(pa, dtype=np.int32),
        pairs=np.array(pf, dtype=np.int32).reshape(-1, 2),
        fem_rowptr=rp, fem_colidx=ci, fem_vals=vals, fem_b=b0, gap_b=gap_b,
        A_rowptr=As.indptr, A_colidx=As.indices, A_vals=As.data, b=b, f=f,
        heat_steps=heat_steps, heat_snap=np.array(snaps), heat_DT=DT, **heat)
    print(f"[{mesh}] poisson/heat: N={N} pairs={len(pf)} fem b gap {gap_b:.1e} OK")


def gen_variants(mesh="mesh5.1"):
    """Fixtures for the draft-script physics variants.  build_mass_and_convection is a function of
    code/StokesColor.py and is executed literally; the rotating-cylinder data, the dye diffusion and the pressure
    smoothing are inline script statements (scripts/stokes_report.py, scripts/good_visualization2.py), restated."""
    ns = L.load_functions("StokesColor.py")
    npth, epth = L.mesh_paths(mesh)
    with L.quiet():
        nodes, markers = ns["readNode"](npth)
        tris = ns["readEle"](epth)
    N = nodes.shape[0]
    rng = np.random.default_rng(11)
    u = rng.standard_normal((N, 2))
    M, Cm = ns["build_mass_and_convection"](nodes, tris, u)
    rp, ci, scatter = R.csr_pattern(N, tris)
    mv, cv = R.mass_convection(nodes, tris, u, rp, ci, scatter)
    Md = sp.csr_matrix((mv, ci, rp), shape=(N, N)).toarray()
    Cd = sp.csr_matrix((cv, ci, rp), shape=(N, N)).toarray()
    assert np.array_equal(Md, M), "consistent mass not bit-exact"
    gap_c = np.abs(Cd - Cm).max() / np.abs(Cm).max()
    assert gap_c <= 1e-15, gap_c                      # np.dot of 2-vectors: BLAS summation / FMA
    wall, inner, _, _ = R.index_sets(nodes, markers)
    urot = rng.standard_normal((N, 2))
    omega = R.ramp_omega(37)
    R.rotating_cylinder_bcu(urot, nodes, wall, inner, omega)
    kvals = R.assemble_stiffness(nodes, tris, rp, ci, scatter)
    K = sp.csr_matrix((kvals, ci, rp), shape=(N, N))
    c_adv = rng.uniform(0.0, 1.0, N)
    c_dif = R.dye_diffuse(c_adv, K, 0.05, 1e-3)
    p_raw = rng.standard_normal(N)
    p_s = R.helmholtz_smooth(K.toarray(), p_raw, ref=3, alpha=0.01)
    np.savez_compressed(os.path.join(OUT, mesh.replace(".", "_") + "_variants.npz"),
                        u=u, M_vals=mv, C_vals_literal=sp.csr_matrix(Cm)[np.repeat(np.arange(N), np.diff(rp)), ci].A1,
                        C_vals=cv, gap_c=gap_c, urot_in=rng.standard_normal((N, 2)) * 0 + 7.0, urot=urot, omega=omega,
                        c_adv=c_adv, c_dif=c_dif, DT=0.05, D=1e-3, p_raw=p_raw, p_smooth=p_s, ref=3, alpha=0.01)
    print(f"[{mesh}] variants: mass bit-exact, convection gap {gap_c:.1e} OK")


def main():
    if not L.available():
        sys.exit("reference checkout not found; fixtures can only be generated in the authoring container")
    os.makedirs(OUT, exist_ok=True)
    for m in MESHES:
        gen_operators(m)
        gen_poisson_heat(m)
    gen_trajectory("mesh5.1", "color_pusher", -2.0, -5.0, 0.05, 0.1, steps=100)
    gen_trajectory("mesh5.1", "food_pusher", -2.0, -5.0, 0.01, 1.0, steps=100, dye=False, food=True)
    gen_trajectory("mesh5.1", "food_neutral", -2.0, 0.0, 0.01, 1.0, steps=100, dye=False, food=True)
    gen_trajectory("mesh_fine.1", "color_puller", -2.0, 5.0, 0.05, 0.1, steps=20)
    gen_variants()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--heat-only":      # regenerate the Poisson / heat fixtures alone
        for m in MESHES:
            gen_poisson_heat(m)
    elif len(sys.argv) > 1 and sys.argv[1] == "--variants-only":
        gen_variants()
    else:
        main()
