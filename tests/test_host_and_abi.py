"""CPU: host-side logic of the product (ingest parser in libfluidsim, pair finder, index
sets, synthetic mesh, CSR row surgery) and the C-ABI surface.  No compute call is made:
without a GPU every compute entry point must fail loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fluidsim_b200 as fb
from fluidsim_b200 import _lib
from conftest import ROOT, load_golden, MESHES


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
