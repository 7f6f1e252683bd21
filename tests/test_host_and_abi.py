"""CPU: host-side logic of the product (ingest parser in libfluidsim, pair finder, index
sets, synthetic mesh, CSR row surgery) and the C-ABI surface.  No compute call is made:
without a GPU every compute entry point must fail loudly (there is no CPU fallback)."""
import ctypes
import ctypes as C
import os
import re

import numpy as np
import pytest

import fluidsim_b200 as fb
from fluidsim_b200 import _lib
from conftest import ROOT, load_golden, MESHES
from oracle import restated as R


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "fluidsim.h")).read()
    declared = set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fs_last_error"} - set(_lib.SIGNATURES)
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib._build.LIB)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fs_version() >= 100


def test_no_cpu_fallback():
    if fb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fb.FluidsimError):
        fb.Mesh(np.array([[0., 0.], [1., 0.], [0., 1.]]), np.array([[0, 1, 2]], dtype=np.int32))


@pytest.mark.parametrize("mesh", MESHES)
def test_triangle_reader_roundtrip(mesh, tmp_path):
    g = load_golden(mesh + "_ops")
    npth, epth = str(tmp_path / "m.1.node"), str(tmp_path / "m.1.ele")
    fb.write_node(npth, g["nodes"], g["markers"])
    fb.write_ele(epth, g["tris"])
    nodes, markers = fb.readNode(npth)
    tris = fb.readEle(epth)
    assert nodes.dtype == np.float64 and markers.dtype == np.int32 and tris.dtype == np.int32
    assert np.array_equal(nodes, g["nodes"]) and np.array_equal(markers, g["markers"])
    assert np.array_equal(tris, g["tris"])
    n32, _ = fb.readNode(npth, dtype=np.float32)          # code/poisson.py:40
    assert np.array_equal(n32, g["nodes"].astype(np.float32))


@pytest.mark.parametrize("name,gold", [("mesh2.1", "mesh2_1"), ("mesh5.1", "mesh5_1"), ("mesh_fine.1", "mesh_fine_1")])
def test_reader_on_reference_files(name, gold):
    base = os.path.join("/root/reference/resources", name)
    if not os.path.exists(base + ".node"):
        pytest.skip("reference checkout not present (GPU box)")
    g = load_golden(gold + "_ops")
    nodes, markers = fb.readNode(base + ".node")
    assert np.array_equal(nodes, g["nodes"]) and np.array_equal(markers, g["markers"])
    assert np.array_equal(fb.readEle(base + ".ele"), g["tris"])
    seg, mk = fb.readPoly(base + ".poly")
    assert seg.shape[1] == 2 and len(mk) == len(seg)


def test_reader_errors(tmp_path):
    with pytest.raises(fb.FluidsimError):
        fb.readNode(str(tmp_path / "missing.node"))
    p = tmp_path / "short.node"
    p.write_text("3 2 0 1\n1 0 0 1\n")
    with pytest.raises(fb.FluidsimError):
        fb.readNode(str(p))
    p = tmp_path / "bad.ele"
    p.write_text("1 3 0\n1 1 x 3\n")
    with pytest.raises(fb.FluidsimError):
        fb.readEle(str(p))


def test_six_node_ele_keeps_corners(tmp_path):
    p = tmp_path / "p2.ele"
    p.write_text("2 6 0\n1 1 2 3 4 5 6\n2 3 2 7 8 9 10\n")
    assert np.array_equal(fb.readEle(str(p)), np.array([[0, 1, 2], [2, 1, 6]], dtype=np.int32))


def test_pairs_and_sets_match_reference(ops):
    nodes, markers = ops["nodes"], ops["markers"]
    pa = fb.find_boundary_pairs(nodes)
    assert np.array_equal(np.array(pa, dtype=np.int32), ops["pairs_all"])
    pf = fb.filter_wall_pairs(nodes, pa)
    assert np.array_equal(np.array(pf, dtype=np.int32).reshape(-1, 2), ops["pairs"])
    wall, inner, _, interior = fb.index_sets(nodes, markers)
    assert np.array_equal(wall, ops["wall"]) and np.array_equal(inner, ops["inner_b"])
    assert np.array_equal(interior, ops["interior"])


def test_pairs_empty_side():
    nodes = np.array([[0.0, 0.0], [0.5, 0.5], [0.0, 1.0]])
    assert fb.find_boundary_pairs(nodes) == []


def test_synthetic_mesh():
    c, m, t = fb.square_with_hole(64, 24)
    assert c.shape == (64 * 25, 2) and t.shape == (2 * 64 * 24, 3)
    x = c[t]
    det = (x[:, 1, 0] - x[:, 0, 0]) * (x[:, 2, 1] - x[:, 0, 1]) - (x[:, 2, 0] - x[:, 0, 0]) * (x[:, 1, 1] - x[:, 0, 1])
    assert det.min() > 0                                          # all CCW like the shipped meshes
    assert np.isclose(0.5 * det.sum(), 1.0 - np.pi * 0.25 ** 2, rtol=2e-3)
    r = np.hypot(c[m == 2, 0] - 0.5, c[m == 2, 1] - 0.5)
    assert np.allclose(r, 0.25, atol=1e-14)
    pairs = fb.filter_wall_pairs(c, fb.find_boundary_pairs(c))
    assert len(pairs) == 64 // 4 - 1
    assert max(abs(c[a, 1] - c[b, 1]) for a, b in pairs) == 0.0   # sides match bit for bit
    wall, inner, _, _ = fb.index_sets(c, m)
    assert len(inner) == 64 and len(wall) == 2 * (64 // 4 + 1)
    with pytest.raises(ValueError):
        fb.square_with_hole(30, 4)


def test_row_editor_matches_dense_surgery():
    """apply_periodic_bc / Dirichlet rows on CSR == the reference's dense in-place edits
    (code/poisson.py:187-213,258-278), including a slave claimed by two masters (mesh2.1)."""
    from fluidsim_b200.poisson import _RowEditor
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 12
    A = sp.random(n, n, density=0.3, random_state=4, format="csr") + sp.eye(n, format="csr")
    A.sort_indices()
    D = A.toarray()
    b = rng.standard_normal(n)
    bd = b.copy()
    pairs = [(0, 5), (2, 7), (3, 7)]
    ed = _RowEditor(A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy())
    for m, s in pairs:
        rm, rs = ed.row(m), ed.row(s)
        for c, v in rs.items():
            rm[c] = rm.get(c, 0.0) + v
        rs.clear(); rs[s] = 1.0; rs[m] = -1.0
        D[m, :] += D[s, :]; D[s, :] = 0.0; D[s, s] = 1.0; D[s, m] = -1.0
    rp, ci, v = ed.finish()
    assert np.array_equal(sp.csr_matrix((v, ci, rp), shape=(n, n)).toarray(), D)
    assert np.all(np.diff(rp) >= 0) and all(np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0) for i in range(n))


def test_refine_mesh():
    g = load_golden("mesh5_1_ops")
    n, m, t = fb.refine_mesh(g["nodes"], g["markers"], g["tris"], 2)
    assert t.shape == (16 * len(g["tris"]), 3) and len(n) == len(m)
    x = n[t]
    det = (x[:, 1, 0] - x[:, 0, 0]) * (x[:, 2, 1] - x[:, 0, 1]) - (x[:, 2, 0] - x[:, 0, 0]) * (x[:, 1, 1] - x[:, 0, 1])
    assert det.min() > 0                                                   # orientation preserved
    assert abs(0.5 * det.sum() - (1 - np.pi * 0.0625)) < 2e-4              # circle nodes projected: area converges
    assert np.allclose(np.hypot(n[m == 2, 0] - 0.5, n[m == 2, 1] - 0.5), 0.25, atol=1e-9)
    assert (m == 2).sum() == 4 * (g["markers"] == 2).sum() and (m == 1).sum() == 4 * (g["markers"] == 1).sum()
    pairs = fb.filter_wall_pairs(n, fb.find_boundary_pairs(n))
    assert len(pairs) == 4 * (len(g["pairs"]) + 1) - 1 and max(abs(n[a, 1] - n[b, 1]) for a, b in pairs) == 0.0
    # Euler characteristic of the annulus: V - E + T = 0
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    E = len(np.unique(np.sort(e, axis=1), axis=0))
    assert len(n) - E + len(t) == 0


def test_raster_argument_validation():
    """Host-side checks of the output sink run before any device call."""
    import fluidsim_b200 as fb
    with pytest.raises(ValueError):
        fb.colorize(np.zeros((4, 4), dtype=np.float32), 0.0, 1.0, cmap=np.zeros((10, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        fb.colorize(np.zeros((4, 4), dtype=np.float32), 0.0, 1.0, background=(0, 0, 0))
    with pytest.raises(ValueError):
        fb.splat_points(np.zeros((4, 4, 3), dtype=np.uint8), np.zeros((1, 2)))
    with pytest.raises(ValueError):
        fb.write_png("/dev/null", np.zeros((4, 4), dtype=np.uint8))
    lut = fb.colormap_lut("gray")
    assert np.array_equal(lut[:, 0], lut[:, 1]) and lut[0, 0] == 0 and lut[255, 0] == 255 and (np.diff(lut[:, 0].astype(int)) >= 0).all()


# ---- mesh generation from a PSLG (`triangle -p -q30 -a...` is not available here) ----
def _quality(P, T):
    a, b, c = P[T[:, 0]], P[T[:, 1]], P[T[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    l = np.sort(np.stack([np.linalg.norm(b - c, axis=1), np.linalg.norm(c - a, axis=1), np.linalg.norm(a - b, axis=1)], 1), axis=1)
    return area2, np.degrees(np.arcsin(np.clip(np.abs(area2) / (l[:, 1] * l[:, 2]), 0, 1)))


def _edges(T):
    e = np.sort(np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]), axis=1)
    return np.unique(e, axis=0, return_counts=True)


@pytest.mark.parametrize("name", ["mesh2_1", "mesh5_1"])
def test_meshgen_reproduces_the_reference_triangulation(name):
    """Triangle's output is a conforming Delaunay triangulation: from the shipped nodes and boundary segments
    the mesher must return exactly the reference's triangles (and add no point: the meshes already satisfy -q30)."""
    import fluidsim_b200 as fb
    g = load_golden(name + "_ops")
    nodes, markers, tris = g["nodes"], g["markers"], g["tris"]
    e, cnt = _edges(tris)
    segs = e[cnt == 1]                                           # boundary edges = the .poly segments
    smark = markers[segs[:, 0]]
    assert np.array_equal(smark, markers[segs[:, 1]]) and set(smark) == {1, 2}
    P, M, T, S, SM = fb.triangulate(nodes, markers, segs, smark, holes=[(0.5, 0.5)], min_angle=30.0)
    assert np.array_equal(P, nodes) and np.array_equal(M, markers)
    assert set(map(tuple, np.sort(T, axis=1).tolist())) == set(map(tuple, np.sort(tris, axis=1).tolist()))
    assert (_quality(P, T)[0] > 0).all()                         # CCW


def test_meshgen_quality_bounds_and_boundaries(tmp_path):
    import fluidsim_b200 as fb
    v, vm, s, sm, h = fb.box_with_hole_pslg(60)
    amax = 5e-4
    P, M, T, S, SM = fb.triangulate(v, vm, s, sm, h, min_angle=30.0, max_area=amax, curves={2: (0.5, 0.5, 0.25)})
    area2, ang = _quality(P, T)
    assert (area2 > 0).all() and ang.min() >= 30.0 - 1e-9 and 0.5 * area2.max() <= amax * (1 + 1e-12)
    assert np.array_equal(P[:len(v)], v) and np.array_equal(M[:len(v)], vm)       # input vertices first, unchanged
    # boundary: every output segment is an edge of exactly one triangle, and nothing else is a boundary edge
    e, cnt = _edges(T)
    bnd = set(map(tuple, e[cnt == 1].tolist()))
    assert bnd == set(map(tuple, np.sort(S, axis=1).tolist())) and (cnt <= 2).all()
    assert len(P) - len(e) + len(T) == 0                                          # Euler characteristic of an annulus
    # markers: 1 on the box, 2 on the circle (new circle points were projected), 0 strictly inside
    on_box = (np.abs(P[:, 0]) < 1e-14) | (np.abs(P[:, 0] - 1) < 1e-14) | (np.abs(P[:, 1]) < 1e-14) | (np.abs(P[:, 1] - 1) < 1e-14)
    rad = np.hypot(P[:, 0] - 0.5, P[:, 1] - 0.5)
    assert np.array_equal(M == 1, on_box) and np.abs(rad[M == 2] - 0.25).max() < 1e-14 and (rad[M == 0] > 0.25).all()
    assert (M == 2).sum() >= 60 and np.array_equal(SM == 2, M[S[:, 0]] == 2)
    # area = box minus the polygon actually meshed (between the 60-gon and the circle)
    assert 1 - np.pi * 0.25 ** 2 < 0.5 * area2.sum() <= 1 - 0.5 * 60 * 0.25 ** 2 * np.sin(2 * np.pi / 60) + 1e-12
    # the host-side set-up of the Stokes solver accepts the mesh; .node/.ele/.poly round trip
    pairs = fb.filter_wall_pairs(P, fb.find_boundary_pairs(P))
    sets = fb.index_sets(P, M)
    assert len(pairs) > 5 and len(sets[0]) > 0 and len(sets[1]) == (M == 2).sum()
    fb.write_node(str(tmp_path / "m.node"), P, M)
    fb.write_ele(str(tmp_path / "m.ele"), T)
    P2, M2 = fb.readNode(str(tmp_path / "m.node"))
    assert np.array_equal(P2, P) and np.array_equal(M2, M) and np.array_equal(fb.readEle(str(tmp_path / "m.ele")), T)


def test_meshgen_nonconvex_outline_and_poly_reader(tmp_path):
    import fluidsim_b200 as fb
    poly = tmp_path / "L.poly"
    poly.write_text("6 2 0 1\n1 0 0 1\n2 1 0 1\n3 1 0.5 1\n4 0.5 0.5 1\n5 0.5 1 1\n6 0 1 1\n"
                    "6 1\n1 1 2 1\n2 2 3 1\n3 3 4 1\n4 4 5 1\n5 5 6 1\n6 6 1 1\n0\n# an L-shaped domain, no holes\n")
    v, vm, s, sm, h = fb.read_poly_full(str(poly))
    assert v.shape == (6, 2) and len(h) == 0 and s.tolist()[0] == [0, 1] and (sm == 1).all() and (vm == 1).all()
    P, M, T, S, SM = fb.triangulate_poly(str(poly), min_angle=28.0, max_area=2e-3)
    area2, ang = _quality(P, T)
    assert abs(0.5 * area2.sum() - 0.75) < 1e-12 and ang.min() >= 28.0 - 1e-9 and 0.5 * area2.max() <= 2e-3 * (1 + 1e-12)
    cen = P[T].mean(1)
    assert not ((cen[:, 0] > 0.5) & (cen[:, 1] > 0.5)).any()      # the notch is outside the domain
    assert (M[(P[:, 0] > 0.51) & (P[:, 0] < 0.99) & (P[:, 1] > 0.01) & (P[:, 1] < 0.49)] == 0).all()
    with pytest.raises(ValueError):
        fb.triangulate(v, vm, s, sm, min_angle=40.0)


def test_meshgen_two_holes_internal_segment_and_sharp_corner():
    import fluidsim_b200 as fb
    v, vm, s, sm, _ = fb.box_with_hole_pslg(40, centre=(0.3, 0.5), radius=0.15)
    v2, _, s2, _, _ = fb.box_with_hole_pslg(30, centre=(0.72, 0.5), radius=0.1)
    V = np.concatenate([v, v2[4:], [[0.1, 0.9], [0.9, 0.9]]])
    VM = np.concatenate([vm, 3 * np.ones(30, dtype=np.int32), [0, 0]])
    S = np.concatenate([s, s2[4:] - 4 + len(v), [[len(v) + 30, len(v) + 31]]])          # last: a constraint inside the domain
    SM = np.concatenate([sm, 3 * np.ones(30, dtype=np.int32), [7]])
    P, M, T, S2, SM2 = fb.triangulate(V, VM, S, SM, [(0.3, 0.5), (0.72, 0.5)], min_angle=30.0, max_area=1e-3)
    area2, ang = _quality(P, T)
    assert ang.min() >= 30.0 - 1e-9 and 0.5 * area2.max() <= 1e-3 * (1 + 1e-12) and (area2 > 0).all()
    want = 1 - 0.5 * 40 * 0.15 ** 2 * np.sin(2 * np.pi / 40) - 0.5 * 30 * 0.1 ** 2 * np.sin(2 * np.pi / 30)
    assert abs(0.5 * area2.sum() - want) < 1e-12
    e, cnt = _edges(T)
    es = set(map(tuple, e.tolist()))
    assert all(tuple(sorted(x)) in es for x in S2.tolist())                              # every (sub)segment is an edge
    assert np.allclose(P[M == 7, 1], 0.9) and (M == 7).sum() >= 1                        # points created on the constraint
    assert len(P) - len(e) + len(T) == -1                                                # two holes
    # a 20-degree input corner: the refinement terminates and leaves only the corner triangle below the bound
    a = np.radians(20.0)
    W = np.array([[0.0, 0.0], [1.0, 0.0], [np.cos(a), np.sin(a)]])
    P, M, T, _, _ = fb.triangulate(W, np.ones(3, dtype=np.int32), np.array([[0, 1], [1, 2], [2, 0]]), np.ones(3, dtype=np.int32),
                                   min_angle=30.0, max_area=2e-3)
    area2, ang = _quality(P, T)
    assert (ang < 30.0 - 1e-9).sum() == 1 and abs(ang.min() - 20.0) < 1e-9 and abs(0.5 * area2.sum() - 0.5 * np.sin(a)) < 1e-12


# ---- round 2: host-side partition helpers, keep-alive pointers, CPU reference arm ---------------------
def test_ptr_keeps_converted_temporaries_alive():
    """ptr(as_f64(x)) must hold the converted copy until the call has returned."""
    import gc
    from fluidsim_b200 import _lib
    x = np.arange(12, dtype=np.float32).reshape(6, 2)[::2]          # needs a copy: wrong dtype, strided
    p = _lib.ptr(_lib.as_f64(x))
    gc.collect()
    assert p._keep is not None and p._keep.dtype == np.float64 and p._keep.flags["C_CONTIGUOUS"]
    assert p.value == p._keep.ctypes.data
    got = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(6,))
    assert np.array_equal(got, np.asarray(x, dtype=np.float64).ravel())


def test_sub_mesh_local_numbering_and_sums():
    """The sub-mesh of a node block reproduces the whole-mesh nodal sums on its owned nodes bit for bit
    (same elements, same order) and numbers own nodes first, halo nodes in ascending global id."""
    nodes, markers, tris = fb.square_with_hole(64, 24)
    N = len(nodes)
    split = fb.node_block_split(N, 3, align=64)
    assert split[0] == 0 and split[-1] == N and all(b % 64 == 0 for b in split[:-1])
    u = np.random.default_rng(5).standard_normal((N, 2))
    div_g = R.divergence(nodes, tris, u)
    mass_g = R.lumped_mass(nodes, tris)
    seen = np.zeros(len(tris), dtype=int)
    for r in range(3):
        lo, hi = split[r], split[r + 1]
        ln, lm, lt, l2g, eids = fb.sub_mesh(nodes, markers, tris, lo, hi)
        n_own = hi - lo
        assert np.array_equal(l2g[:n_own], np.arange(lo, hi)) and np.all(np.diff(l2g[n_own:]) > 0)
        assert np.array_equal(l2g[lt], tris[eids]) and np.array_equal(ln, nodes[l2g]) and np.array_equal(lm, markers[l2g])
        assert np.all(np.diff(eids) > 0)
        seen[eids] += 1
        assert np.array_equal(R.divergence(ln, lt, u[l2g])[:n_own], div_g[lo:hi])
        assert np.array_equal(R.lumped_mass(ln, lt)[:n_own], mass_g[lo:hi])
    assert seen.min() >= 1                                          # every element belongs to at least one block


def test_local_index_sets_and_cut_pairs():
    nodes, markers, tris = fb.square_with_hole(64, 24)
    pairs = fb.filter_wall_pairs(nodes, fb.find_boundary_pairs(nodes))
    wall, inner, _, interior = fb.index_sets(nodes, markers)
    N = len(nodes)
    split = fb.node_block_split(N, 2, align=64)
    tot = 0
    for r in range(2):
        lo, hi = split[r], split[r + 1]
        w, i, it, pr = fb.local_index_sets(lo, hi, wall, inner, interior, pairs)
        assert np.array_equal(w + lo, wall[(wall >= lo) & (wall < hi)])
        assert np.array_equal(i + lo, inner[(inner >= lo) & (inner < hi)])
        tot += len(pr)
        for m, s in pr:
            assert (m + lo, s + lo) in pairs
    assert tot == len(pairs)                                        # all pairs sit on the outer ring: one block has them all
    a, b = pairs[0]
    cut = (min(a, b) + max(a, b)) // 2 + 1                          # a block boundary between the two nodes of a pair
    with pytest.raises(ValueError):
        fb.local_index_sets(0, cut, wall, inner, interior, pairs)


def test_cpu_step_matches_restated_oracle():
    """oracle/cpu_step.py (the CPU arm of bench.py): sparse div/grad operators and both pressure solvers against
    the restated oracle's direct solves."""
    from oracle import cpu_step
    nodes, markers, tris = fb.square_with_hole(128, 32)
    Dx, Dy = cpu_step.grad_operators(nodes, tris)
    u = np.random.default_rng(0).standard_normal((len(nodes), 2))
    d0 = R.divergence(nodes, tris, u)
    assert np.abs(d0 - (Dx @ u[:, 0] + Dy @ u[:, 1])).max() <= 1e-13 * np.abs(d0).max()
    gx, gy = R.gradient(nodes, tris, u[:, 0])
    assert np.abs(gx - Dx @ u[:, 0]).max() <= 1e-13 * np.abs(gx).max() and np.abs(gy - Dy @ u[:, 0]).max() <= 1e-13 * np.abs(gy).max()
    ref = R.RestatedStokes(nodes, markers, tris, B1=-2.0, B2=-5.0)
    sims = {p: cpu_step.CpuStokes(nodes, markers, tris, B1=-2.0, B2=-5.0, precond=p) for p in ("amg", "jacobi")}
    for _ in range(4):
        ref.flow_step()
        for p, s in sims.items():
            it = s.step()
            assert np.linalg.norm(s.u - ref.u) <= 1e-9 * np.linalg.norm(ref.u), p
            assert np.linalg.norm(s.p - ref.p) <= 1e-7 * np.linalg.norm(ref.p), p
    assert sims["amg"].iters[1] < 60 < sims["jacobi"].iters[1]


def test_cpu_recycler_projection_matches_direct_solves_with_fewer_iterations():
    """oracle/cpu_step.py: Recycler (the CPU restatement of csrc/recycle.cu): starting every pressure solve from the
    projection onto the span of the previous solutions leaves the fields where the restated oracle's direct solves put
    them and needs fewer PCG iterations than the time-extrapolated guess once the basis exists; the run passes through
    a compression of the basis (12 -> <= 6 vectors at the 12th solve)."""
    from oracle import cpu_step
    nodes, markers, tris = fb.square_with_hole(128, 32)
    ref = R.RestatedStokes(nodes, markers, tris, B1=-2.0, B2=-5.0)
    a = cpu_step.CpuStokes(nodes, markers, tris, B1=-2.0, B2=-5.0, precond="amg", recycle=True)
    b = cpu_step.CpuStokes(nodes, markers, tris, B1=-2.0, B2=-5.0, precond="amg", recycle=False)
    ia, ib, sizes = [], [], []
    for _ in range(16):
        ref.flow_step()
        ia.append(a.step()[1:])
        ib.append(b.step()[1:])
        sizes.append(a.rec[0].k)
    assert np.linalg.norm(a.u - ref.u) <= 1e-9 * np.linalg.norm(ref.u)
    assert np.linalg.norm(a.u - b.u) <= 1e-9 * np.linalg.norm(b.u)
    assert max(sizes) == 11 and sizes[11] <= 6 and sizes[-1] > sizes[11]        # grew, was compressed, grows again
    assert sum(map(sum, ia[8:])) <= 0.8 * sum(map(sum, ib[8:])), (ia, ib)


def test_bench_reference_arm_runs_without_the_product():
    import json, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-theta", "128", "--n-r", "32",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["product_loaded"] is False and out["gpu_launches"] == 0
    assert out["steps"] == 2 and out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1
    assert out["value"] > 0 and out["e2e"]["h2d_bytes_per_step"] == 0 and "cpu_baseline_jacobi" in out
