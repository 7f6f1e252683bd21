"""Host-side mesh helpers in plain numpy: periodic pairs and index sets (the reference's setup
rules), the synthetic square-with-hole generator of the benchmarks, uniform refinement and the
Triangle writers.  No import of the shared library: `mesh.py` re-exports everything here, and
CPU-only tools (bench.py --impl reference) load this file by path.
"""
from __future__ import annotations

import numpy as np


def write_node(path, coords, markers):
    """Triangle .node writer (1-based ids, 17 significant digits)."""
    coords = np.asarray(coords, dtype=np.float64)
    with open(path, "w") as f:
        f.write(f"{coords.shape[0]}  2  0  1\n")
        for i, ((x, y), m) in enumerate(zip(coords, markers)):
            f.write(f"{i + 1:4d}    {float(x)!r}  {float(y)!r}    {int(m)}\n")
        f.write("# written by fluidsim_b200\n")


def write_ele(path, tris):
    tris = np.asarray(tris)
    with open(path, "w") as f:
        f.write(f"{tris.shape[0]}  3  0\n")
        for e, t in enumerate(tris):
            f.write(f"{e + 1:4d}    {t[0] + 1:4d}  {t[1] + 1:4d}  {t[2] + 1:4d}\n")
        f.write("# written by fluidsim_b200\n")


# --------------------------------------------------------------------------- boundary sets
def find_boundary_pairs(nodes_coords, L=1.0, tol=1e-6):
    """[(left_id, right_id)] in ascending left id: for each node on x=0 the node on
    x=L nearest in y.  code/StokesColor.py:169-203.  At an exact distance tie (the
    reference's KDTree result is implementation-defined there) the smaller y wins,
    which is what the reference returns on resources/mesh2.1."""
    nodes_coords = np.asarray(nodes_coords)
    left = np.where(np.abs(nodes_coords[:, 0]) < tol)[0]
    right = np.where(np.abs(nodes_coords[:, 0] - L) < tol)[0]
    if len(left) == 0 or len(right) == 0:
        print("Warning: One or both boundaries have no nodes.")
        return []
    ry = nodes_coords[right, 1].astype(np.float64)
    order = np.argsort(ry, kind="stable")
    rys = ry[order]
    ly = nodes_coords[left, 1].astype(np.float64)
    pos = np.searchsorted(rys, ly)
    lo = np.clip(pos - 1, 0, len(rys) - 1)
    hi = np.clip(pos, 0, len(rys) - 1)
    pick = np.where(np.abs(ly - rys[lo]) <= np.abs(rys[hi] - ly), lo, hi)
    # among equal y's keep the first in sorted order
    first = np.searchsorted(rys, rys[pick], side="left")
    return [(int(a), int(right[order[b]])) for a, b in zip(left, first)]


def filter_wall_pairs(nodes_coords, pairs, H=1.0, tol=1e-6):
    """code/StokesColor.py:449-457: drop pairs whose master sits on y=0 or y=H."""
    out = []
    for m, s in pairs:
        my = nodes_coords[m, 1]
        if not (abs(my - 0.0) < tol or abs(my - H) < tol):
            out.append((m, s))
    return out


def index_sets(nodes_coords, markers, H=1.0, tol=1e-6, inner_marker=2):
    """(wall, inner_boundary, dirichlet, interior) of code/StokesColor.py:461-464:
    walls by y-coordinate, inner boundary by marker."""
    y = np.asarray(nodes_coords)[:, 1]
    wall = np.where(np.isclose(y, 0.0, atol=tol) | np.isclose(y, H, atol=tol))[0]
    inner = np.where(np.asarray(markers) == inner_marker)[0]
    dirichlet = np.union1d(wall, inner)
    interior = np.setdiff1d(np.arange(len(y)), dirichlet)
    return wall, inner, dirichlet, interior


# --------------------------------------------------------------------------- synthetic mesh
def square_with_hole(n_theta, n_r, radius=0.25, center=(0.5, 0.5), half=0.5):
    """Structured triangulation of the unit square with a circular hole, the
    topological annulus of SURVEY.md section 8(d): n_theta angular x n_r radial
    quads, two CCW triangles each (T = 2*n_theta*n_r, N = n_theta*(n_r+1)).
    Ring 0 lies on the circle (marker 2), the last ring is snapped onto the box
    (marker 1) so the x=0 / x=1 nodes match in y.  Node id = ring*n_theta + j,
    i.e. contiguous rings: a row-block partition cuts along rings."""
    if n_theta % 8:
        raise ValueError("n_theta must be a multiple of 8")
    j = np.arange(n_theta)
    th = 2.0 * np.pi * j / n_theta
    ct, st = np.cos(th), np.sin(th)
    # exact box direction: scale the ray so the larger of |cos|,|sin| becomes `half`
    rbox = half / np.maximum(np.abs(ct), np.abs(st))
    s = np.linspace(0.0, 1.0, n_r + 1)[:, None]
    r = (1.0 - s) * radius + s * rbox[None, :]
    x = center[0] + r * ct[None, :]
    y = center[1] + r * st[None, :]
    # snap the outer ring exactly onto the box so the periodic sides match bit for bit
    xo, yo = x[-1], y[-1]
    on_v = np.abs(ct) >= np.abs(st)
    xo[on_v] = center[0] + half * np.sign(ct[on_v])
    yo[~on_v] = center[1] + half * np.sign(st[~on_v])
    corner = np.abs(np.abs(ct) - np.abs(st)) < 1e-12
    yo[corner] = center[1] + half * np.sign(st[corner])
    # symmetrise y on the two vertical sides (theta and pi-theta give the same y)
    k = n_theta // 2
    yr = yo.copy()
    for jj in np.where(on_v)[0]:
        mirror = (k - jj) % n_theta
        yr[jj] = yr[mirror] = 0.5 * (yo[jj] + yo[mirror])
    y[-1] = yr
    coords = np.stack([x.ravel(), y.ravel()], axis=1)
    markers = np.zeros((n_r + 1, n_theta), dtype=np.int32)
    markers[0] = 2
    markers[-1] = 1
    i = np.arange(n_r)[:, None]
    jn = (j + 1) % n_theta
    a = i * n_theta + j[None, :]
    b = i * n_theta + jn[None, :]
    c = (i + 1) * n_theta + jn[None, :]
    d = (i + 1) * n_theta + j[None, :]
    t1 = np.stack([a, c, b], axis=-1)
    t2 = np.stack([a, d, c], axis=-1)
    tris = np.stack([t1, t2], axis=2).reshape(-1, 3).astype(np.int32)
    return np.ascontiguousarray(coords), markers.ravel(), tris


# --------------------------------------------------------------------------- refinement
def refine_mesh(nodes_coords, markers, triangles, levels=1, circle=((0.5, 0.5), 0.25), inner_marker=2):
    """Uniform red refinement (every triangle -> 4) of a Triangle mesh, `levels` times.

    New mid-edge nodes on a boundary edge (an edge owned by a single triangle) inherit the marker its
    two end nodes share; nodes created on the inner boundary (marker ``inner_marker``) are projected
    onto the circle so that the refined squirmer boundary stays round.  Orientation is preserved.
    Gives refined UNSTRUCTURED meshes from the shipped ones without the external `triangle` binary
    (SURVEY section 8 f3)."""
    nodes = np.asarray(nodes_coords, dtype=np.float64)
    mk = np.asarray(markers, dtype=np.int32)
    tris = np.asarray(triangles, dtype=np.int64)
    (cx, cy), rad = circle
    for _ in range(levels):
        n = nodes.shape[0]
        e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])      # (3T,2): edges 01, 12, 20
        lo, hi = e.min(axis=1), e.max(axis=1)
        key = lo * n + hi
        uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        ua, ub = uniq // n, uniq % n
        mid = 0.5 * (nodes[ua] + nodes[ub])
        mm = np.where((cnt == 1) & (mk[ua] == mk[ub]) & (mk[ua] != 0), mk[ua], 0).astype(np.int32)
        on_circle = mm == inner_marker
        if on_circle.any():
            d = mid[on_circle] - np.array([cx, cy])
            mid[on_circle] = np.array([cx, cy]) + rad * d / np.linalg.norm(d, axis=1, keepdims=True)
        t = len(tris)
        m01, m12, m20 = n + inv[:t], n + inv[t:2 * t], n + inv[2 * t:]
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        tris = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1),
                               np.stack([m20, m12, c], 1), np.stack([m01, m12, m20], 1)])
        nodes = np.concatenate([nodes, mid])
        mk = np.concatenate([mk, mm])
    return np.ascontiguousarray(nodes), mk, np.ascontiguousarray(tris.astype(np.int32))


# --------------------------------------------------------------------------- partition (one block of nodes per GPU)
def node_block_split(n_nodes, world, align=1):
    """world+1 boundaries of contiguous node blocks, aligned to ``align`` nodes (n_theta for the ring-numbered
    synthetic mesh: a block is then a set of whole rings and touches at most two neighbours)."""
    nblk = (n_nodes + align - 1) // align
    if nblk < world:
        raise ValueError(f"{n_nodes} nodes in units of {align} cannot be split over {world} ranks")
    return [min(n_nodes, ((nblk * r) // world) * align) for r in range(world)] + [n_nodes]


def sub_mesh(nodes_coords, markers, triangles, lo, hi):
    """The part of the mesh rank [lo, hi) works on: every element that touches one of its nodes.

    Local numbering: the owned nodes first (global order), then the halo nodes -- the other corners of
    those elements -- in ascending global id.  Elements keep their global order, so nodal sums over
    incident elements add in the same order as on the whole mesh.
    Returns (local_nodes, local_markers, local_triangles, l2g, element_ids)."""
    tris = np.asarray(triangles)
    own = (tris >= lo) & (tris < hi)
    eids = np.nonzero(own.any(axis=1))[0]
    lt = tris[eids]
    ext = lt[(lt < lo) | (lt >= hi)]
    halo = np.unique(ext)
    l2g = np.concatenate([np.arange(lo, hi, dtype=np.int64), halo.astype(np.int64)])
    n_own = hi - lo
    inside = (lt >= lo) & (lt < hi)
    loc = np.where(inside, lt - lo, n_own + np.searchsorted(halo, lt)).astype(np.int32)
    return (np.ascontiguousarray(np.asarray(nodes_coords)[l2g]), np.ascontiguousarray(np.asarray(markers)[l2g]),
            np.ascontiguousarray(loc), l2g.astype(np.int32), eids)


def local_index_sets(lo, hi, wall, inner, interior, pairs):
    """The owned part of the boundary sets in local numbering (own node g -> g - lo).  A periodic pair must
    lie inside one block (the pressure dof it merges into belongs to one rank)."""
    def own(a):
        a = np.asarray(a, dtype=np.int64)
        return (a[(a >= lo) & (a < hi)] - lo).astype(np.int32)
    lp = []
    for m, s in pairs:
        mi, si = lo <= m < hi, lo <= s < hi
        if mi != si:
            raise ValueError(f"periodic pair ({m}, {s}) is cut by the node block [{lo}, {hi})")
        if mi:
            lp.append((m - lo, s - lo))
    return own(wall), own(inner), own(interior), lp
