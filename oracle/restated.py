"""Restated tier of the oracle: CPU (numpy / scipy.sparse) restatement of the
reference's Stokes-step hot path in sparse, well-posed form.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py`` may import this module, and only as the
checker / the timed CPU baseline -- never as part of the product path.

Pinning: every function here is checked against the *literal* reference
functions (``oracle/literal.py``: the reference's own code exec'd in the
authoring container) by ``oracle/gen_golden.py`` when it writes the fixtures in
``tests/golden/``, and again by ``tests/test_oracle_golden.py`` against those
committed fixtures.  Two boundaries stay **parity unpinned** (SURVEY.md §8c):
  * the pressure solve: the reference's ``A_pressure`` is singular (cond 2.6e17)
    and its LU answer is rounding-noise-determined at the 1e-3 level; this
    module restates it as the SPD, periodic-merged, mean-free system below and
    the fixtures record the measured gap to the literal LU;
  * the StokesFood tracer interpolation, which the reference delegates to
    ``matplotlib.tri.LinearTriInterpolator`` (absent here, unpinned version).

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial import KDTree


# --------------------------------------------------------------------------- ingest
def read_node(path, dtype=np.float64):
    """code/StokesColor.py:54-78 (dtype float64) / code/poisson.py:27-56 (float32)."""
    with open(path) as fh:
        n = int(fh.readline().split()[0])
        coords = np.zeros((n, 2), dtype=dtype)
        markers = np.zeros(n, dtype=np.int32)
        for _ in range(n):
            t = fh.readline().split()
            i = int(t[0]) - 1
            coords[i, 0] = float(t[1])
            coords[i, 1] = float(t[2])
            if int(t[3]) != 0:
                markers[i] = int(t[3])
    return coords, markers


def read_ele(path):
    """code/StokesColor.py:82-95: line order is the element id, column 0 ignored."""
    with open(path) as fh:
        t = int(fh.readline().split()[0])
        tris = np.zeros((t, 3), dtype=np.int32)
        for e in range(t):
            tok = fh.readline().split()
            tris[e] = (int(tok[1]) - 1, int(tok[2]) - 1, int(tok[3]) - 1)
    return tris


# --------------------------------------------------------------------------- pattern
def dof_map_from_pairs(n, pairs):
    """Periodic merge: every (master, slave) pair shares one dof.  Classes are
    formed by union-find; the representative is the smallest node id of the
    class; dofs are numbered in ascending representative order.  Returns
    (dof (n,) int32, n_dof)."""
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for m, s in pairs:
        ra, rb = find(int(m)), find(int(s))
        if ra != rb:
            lo, hi = min(ra, rb), max(ra, rb)
            parent[hi] = lo
    rep = np.array([find(i) for i in range(n)])
    is_rep = rep == np.arange(n)
    new_id = np.cumsum(is_rep) - 1
    return new_id[rep].astype(np.int32), int(is_rep.sum())


def csr_pattern(n, tris, dof=None):
    """Structural CSR pattern of the P1 stiffness matrix: entry (a,b) exists iff
    some triangle holds both nodes (code/StokesColor.py:120-126 writes exactly
    those).  Columns sorted ascending.  With ``dof`` the node ids are first
    mapped (periodic merge).  Returns rowptr (n+1) int32, colidx int32 and
    scatter (T,9) int32 = position of element entry (i,j) in the value array."""
    t = np.asarray(tris, dtype=np.int64)
    if dof is not None:
        t = np.asarray(dof, dtype=np.int64)[t]
    rows = np.repeat(t, 3, axis=1).ravel()          # i-major: (i,j) -> i*3+j
    cols = np.tile(t, (1, 3)).ravel()
    key = rows * n + cols
    uniq, inv = np.unique(key, return_inverse=True)
    r = (uniq // n).astype(np.int64)
    c = (uniq % n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, c, inv.reshape(-1, 9).astype(np.int32)


# --------------------------------------------------------------------------- element maths
def _geom(nodes, tris):
    x = nodes[tris, 0]
    y = nodes[tris, 1]
    x1, x2, x3 = x[:, 0], x[:, 1], x[:, 2]
    y1, y2, y3 = y[:, 0], y[:, 1], y[:, 2]
    det = x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)
    yd = np.stack([y2 - y3, y3 - y1, y1 - y2], axis=1)
    xd = np.stack([x3 - x2, x1 - x3, x2 - x1], axis=1)
    return det, yd, xd


def element_stiffness(nodes, tris):
    """code/StokesColor.py:111-124.  Returns (T,9) values (i-major) and the skip mask."""
    det, yd, xd = _geom(nodes, tris)
    skip = np.abs(det) < 1e-14
    num = yd[:, :, None] * yd[:, None, :] + xd[:, :, None] * xd[:, None, :]
    den = 2 * np.abs(det)
    with np.errstate(divide="ignore", invalid="ignore"):
        ke = num / den[:, None, None]
    ke[skip] = 0.0
    return ke.reshape(-1, 9), skip


def assemble_stiffness(nodes, tris, rowptr, colidx, scatter):
    """Values of K on the structural pattern; contributions are added in
    ascending element order, (i,j) i-major inside an element, exactly like the
    dense ``AMatrix[tri[i], tri[j]] += ...`` loop (code/StokesColor.py:103-126),
    so the result is bit-identical to the literal dense matrix."""
    ke, skip = element_stiffness(nodes, tris)
    keep = ~skip
    return np.bincount(scatter[keep].ravel(), weights=ke[keep].ravel(),
                       minlength=len(colidx))


def lumped_mass(nodes, tris):
    """code/StokesColor.py:266-284 (no degenerate skip)."""
    det, _, _ = _geom(nodes, tris)
    third = (0.5 * np.abs(det)) / 3.0
    return np.bincount(np.asarray(tris).ravel(), weights=np.repeat(third, 3),
                       minlength=nodes.shape[0])


def _area_sum(nodes, tris, det):
    keep = np.abs(det) >= 1e-14
    third = (0.5 * np.abs(det)) / 3.0
    return keep, third, np.bincount(np.asarray(tris)[keep].ravel(),
                                    weights=np.repeat(third[keep], 3),
                                    minlength=nodes.shape[0])


def divergence_sum(nodes, tris, u):
    """Un-normalised lumped divergence and area_sum, code/StokesColor.py:139-163."""
    det, yd, xd = _geom(nodes, tris)
    keep, third, area_sum = _area_sum(nodes, tris, det)
    with np.errstate(divide="ignore"):
        inv2a = 1.0 / det
    ux = u[tris, 0]
    uy = u[tris, 1]
    dux = (ux[:, 0] * yd[:, 0] + ux[:, 1] * yd[:, 1] + ux[:, 2] * yd[:, 2]) * inv2a
    duy = (uy[:, 0] * xd[:, 0] + uy[:, 1] * xd[:, 1] + uy[:, 2] * xd[:, 2]) * inv2a
    lump = (dux + duy) * third
    div_sum = np.bincount(np.asarray(tris)[keep].ravel(), weights=np.repeat(lump[keep], 3),
                          minlength=nodes.shape[0])
    return div_sum, area_sum


def divergence(nodes, tris, u):
    """calculate_divergence, code/StokesColor.py:130-165."""
    div_sum, area_sum = divergence_sum(nodes, tris, u)
    return div_sum / (area_sum + 1e-12)


def gradient(nodes, tris, p):
    """calculate_gradiant, code/StokesColor.py:224-263.  The reference forms
    ``grads.T @ p_local`` with a BLAS matmul whose summation order is not
    defined; here it is ((g0*p0 + g1*p1) + g2*p2) with g_i = diff_i * inv2A."""
    det, yd, xd = _geom(nodes, tris)
    keep, third, area_sum = _area_sum(nodes, tris, det)
    with np.errstate(divide="ignore"):
        inv2a = 1.0 / det
    pl = p[tris]
    gx = (yd[:, 0] * inv2a) * pl[:, 0] + (yd[:, 1] * inv2a) * pl[:, 1] + (yd[:, 2] * inv2a) * pl[:, 2]
    gy = (xd[:, 0] * inv2a) * pl[:, 0] + (xd[:, 1] * inv2a) * pl[:, 1] + (xd[:, 2] * inv2a) * pl[:, 2]
    n = nodes.shape[0]
    t = np.asarray(tris)[keep].ravel()
    sx = np.bincount(t, weights=np.repeat((gx * third)[keep], 3), minlength=n)
    sy = np.bincount(t, weights=np.repeat((gy * third)[keep], 3), minlength=n)
    return sx / (area_sum + 1e-12), sy / (area_sum + 1e-12)


# --------------------------------------------------------------------------- boundary sets
def find_boundary_pairs(nodes, L=1.0, tol=1e-6):
    """code/StokesColor.py:169-203: for each left node (ascending id) the right
    node nearest in y (cKDTree over the right y's, like the reference)."""
    left = np.where(np.abs(nodes[:, 0]) < tol)[0]
    right = np.where(np.abs(nodes[:, 0] - L) < tol)[0]
    if len(left) == 0 or len(right) == 0:
        return []
    tree = KDTree(nodes[right, 1].reshape(-1, 1))
    # one query per left node, like the reference: at exact distance ties the
    # batched query of KDTree can pick the other candidate (mesh2.1, y=0.6875)
    return [(int(a), int(right[tree.query([nodes[a, 1]])[1]])) for a in left]


def filter_wall_pairs(nodes, pairs, H=1.0, tol=1e-6):
    """code/StokesColor.py:449-457."""
    out = []
    for m, s in pairs:
        my = nodes[m, 1]
        if not (abs(my - 0.0) < tol or abs(my - H) < tol):
            out.append((m, s))
    return out


def index_sets(nodes, markers, H=1.0, tol=1e-6):
    """code/StokesColor.py:461-464: walls by y-coordinate, inner by marker 2."""
    wall = np.where(np.isclose(nodes[:, 1], 0.0, atol=tol)
                    | np.isclose(nodes[:, 1], H, atol=tol))[0]
    inner = np.where(markers == 2)[0]
    dirichlet = np.union1d(wall, inner)
    interior = np.setdiff1d(np.arange(nodes.shape[0]), dirichlet)
    return wall, inner, dirichlet, interior


def make_dir_bcu(u, nodes, wall, inner, B1, B2):
    """makeDirBCU, code/StokesColor.py:405-427."""
    u[wall] = 0.0
    th = np.arctan2(nodes[inner, 1] - 0.5, nodes[inner, 0] - 0.5)
    vt = B1 * np.sin(th) + B2 * np.sin(2 * th)
    u[inner, 0] = vt * (-np.sin(th))
    u[inner, 1] = vt * np.cos(th)


def make_per_bcu(u, pairs):
    """makePerBCU, code/StokesColor.py:429-431 (sequential)."""
    for m, s in pairs:
        u[s] = u[m]


# --------------------------------------------------------------------------- linear systems
def viscous_matrix(n, rowptr, colidx, kvals, dirichlet, DT, v):
    """A_visc = I + DT*v*K with Dirichlet rows and columns zeroed, diagonal 1
    (code/StokesColor.py:471-475), on K's structural pattern."""
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    isd = np.zeros(n, dtype=bool)
    isd[dirichlet] = True
    vals = (rows == colidx).astype(np.float64) + (DT * v) * kvals
    kill = isd[rows] | isd[colidx]
    vals[kill] = 0.0
    vals[kill & (rows == colidx)] = 1.0
    return vals


class PressureSystem:
    """Well-posed restatement of ``solve(A_pressure, b_p)``
    (code/StokesColor.py:478-479,554-555).

    The reference's A_pressure = diag(1/(M+1e-12)) K + 1e10-penalty on the
    periodic pairs is singular and non-symmetric.  Restated: merge each periodic
    pair into one dof (Z), multiply the equations by the lumped mass, and solve
        (Z^T K Z) q = Z^T (M * b) - mean,      p = Z (q - mean(q))
    which is SPD on the mean-free subspace.  SURVEY.md §5.9 / §7.2.
    """

    def __init__(self, nodes, tris, pairs):
        n = nodes.shape[0]
        self.n = n
        self.dof, self.nd = dof_map_from_pairs(n, pairs)
        self.rowptr, self.colidx, self.scatter = csr_pattern(self.nd, tris, self.dof)
        self.vals = assemble_stiffness(nodes, tris, self.rowptr, self.colidx, self.scatter)
        self.K = sp.csr_matrix((self.vals, self.colidx, self.rowptr), shape=(self.nd, self.nd))
        self.M = lumped_mass(nodes, tris)
        self._lu = None

    def reduce_rhs(self, b):
        r = np.bincount(self.dof, weights=self.M * b, minlength=self.nd)
        return r - r.sum() / self.nd

    def solve(self, b):
        """Direct solve (dof 0 pinned, then mean removed)."""
        r = self.reduce_rhs(b)
        if self._lu is None:
            self._lu = spla.splu(self.K[1:, 1:].tocsc())
        q = np.zeros(self.nd)
        q[1:] = self._lu.solve(r[1:])
        # one step of iterative refinement against the full singular operator
        res = r - self.K @ q
        res -= res.mean()
        dq = np.zeros(self.nd)
        dq[1:] = self._lu.solve(res[1:])
        q += dq
        q -= q.mean()
        return q[self.dof]

    def solve_cg(self, b, x0=None, rtol=1e-10, maxiter=100000, jacobi=False):
        r = self.reduce_rhs(b)
        q0 = None
        if x0 is not None:
            q0 = np.zeros(self.nd)
            q0[self.dof] = x0
        it = [0]
        Mop = None
        if jacobi:
            d = 1.0 / self.K.diagonal()
            Mop = spla.LinearOperator((self.nd, self.nd), matvec=lambda z: d * z)
        q, info = spla.cg(self.K, r, x0=q0, rtol=rtol, atol=0.0, maxiter=maxiter, M=Mop,
                          callback=lambda _: it.__setitem__(0, it[0] + 1))
        q = q - q.mean()
        return q[self.dof], it[0]


class RestatedStokes:
    """The operator-split step of code/StokesColor.py:537-575 in sparse form."""

    def __init__(self, nodes, markers, tris, B1=-2.0, B2=0.0, DT=0.05, v=0.1, H=1.0, tol=1e-6):
        self.nodes, self.markers, self.tris = nodes, markers, np.asarray(tris)
        n = nodes.shape[0]
        self.N = n
        self.B1, self.B2, self.DT, self.v = B1, B2, DT, v
        self.all_pairs = find_boundary_pairs(nodes, 1.0, tol)
        self.pairs = filter_wall_pairs(nodes, self.all_pairs, H, tol)
        self.wall, self.inner_b, self.dirichlet, self.interior = index_sets(nodes, markers, H, tol)
        self.rowptr, self.colidx, self.scatter = csr_pattern(n, tris)
        self.kvals = assemble_stiffness(nodes, tris, self.rowptr, self.colidx, self.scatter)
        self.M = lumped_mass(nodes, tris)
        av = viscous_matrix(n, self.rowptr, self.colidx, self.kvals, self.dirichlet, DT, v)
        self.A_visc = sp.csr_matrix((av, self.colidx, self.rowptr), shape=(n, n))
        self._visc_lu = spla.splu(self.A_visc.tocsc())
        self.psys = PressureSystem(nodes, tris, self.pairs)
        self.u = np.zeros((n, 2))
        make_dir_bcu(self.u, nodes, self.wall, self.inner_b, B1, B2)
        self.p = np.zeros(n)
        self.p2 = np.zeros(n)

    def _visc_solve(self, rhs):
        x = self._visc_lu.solve(rhs)
        x += self._visc_lu.solve(rhs - self.A_visc @ x)
        return x

    omega = None      # set to an angular velocity: rotating-cylinder data instead of the squirmer's (stokes_report.py:1155-1171)

    def _dirichlet(self, v):
        if self.omega is None:
            make_dir_bcu(v, self.nodes, self.wall, self.inner_b, self.B1, self.B2)
        else:
            rotating_cylinder_bcu(v, self.nodes, self.wall, self.inner_b, self.omega)

    def flow_step(self):
        DT, nodes, tris = self.DT, self.nodes, self.tris
        u = self.u
        us = np.stack([self._visc_solve(u[:, 0].copy()), self._visc_solve(u[:, 1].copy())], axis=1)
        make_per_bcu(us, self.pairs)
        self._dirichlet(us)
        self.div_u_star = divergence(nodes, tris, us)
        p = self.psys.solve(-(1.0 / DT) * self.div_u_star)
        gx, gy = gradient(nodes, tris, p)
        u[:, 0] = us[:, 0] - DT * gx
        u[:, 1] = us[:, 1] - DT * gy
        make_per_bcu(u, self.pairs)
        self._dirichlet(u)
        p2 = self.psys.solve(-(1.0 / DT) * divergence(nodes, tris, u))
        g2x, g2y = gradient(nodes, tris, p2)
        u[self.interior, 0] -= DT * g2x[self.interior]
        u[self.interior, 1] -= DT * g2y[self.interior]
        self.p, self.p2 = p, p2
        self.final_div = divergence(nodes, tris, u)


# --------------------------------------------------------------------------- locator / dye
class Locator:
    """PointLocator, code/StokesColor.py:314-345: k=10 nearest centroids in
    ascending distance, first triangle whose three weights are >= 0, else -1."""

    def __init__(self, nodes, tris):
        self.nodes, self.tris = nodes, np.asarray(tris)
        self.centroids = np.mean(nodes[self.tris], axis=1)
        self.tree = KDTree(self.centroids)

    def find(self, pts, k=10):
        pts = np.atleast_2d(pts)
        kk = min(k, len(self.tris))
        _, idx = self.tree.query(pts, k=kk)
        idx = idx.reshape(len(pts), kk)
        out = np.full(len(pts), -1, dtype=np.int32)
        x, y = pts[:, 0], pts[:, 1]
        for c in range(kk):
            tid = idx[:, c]
            n = self.nodes[self.tris[tid]]
            x1, y1, x2, y2, x3, y3 = n[:, 0, 0], n[:, 0, 1], n[:, 1, 0], n[:, 1, 1], n[:, 2, 0], n[:, 2, 1]
            det = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
            ok = np.abs(det) >= 1e-14
            with np.errstate(divide="ignore", invalid="ignore"):
                w1 = ((x2 - x) * (y3 - y) - (x3 - x) * (y2 - y)) / det
                w2 = ((x3 - x) * (y1 - y) - (x1 - x) * (y3 - y)) / det
            w3 = 1.0 - w1 - w2
            hit = ok & (w1 >= 0.0) & (w2 >= 0.0) & (w3 >= 0.0) & (out < 0)
            out[hit] = tid[hit]
        return out


def _pdx(a, b):
    d = a - b
    d = np.where(d > 0.5, d - 1.0, d)
    d = np.where(d < -0.5, d + 1.0, d)
    return d


def backtrace(nodes, u, DT):
    """code/StokesColor.py:361-366."""
    xb = (nodes[:, 0] - DT * u[:, 0] * 1.0) % 1.0
    yb = nodes[:, 1] - DT * u[:, 1] * 1.0
    yb = np.where(yb < 0.0, 1e-12, yb)
    yb = np.where(yb > 1.0, 1.0 - 1e-12, yb)
    return xb, yb


def advect_semilagrange(c, u, DT, nodes, tris, locator):
    """code/StokesColor.py:347-389, vectorised; returns the host triangle ids."""
    xb, yb = backtrace(nodes, u, DT)
    tid = locator.find(np.stack([xb, yb], axis=1))
    hit = tid >= 0
    t = np.asarray(tris)[np.where(hit, tid, 0)]
    i, j, k = t[:, 0], t[:, 1], t[:, 2]
    x1, y1, x2, y2, x3, y3 = nodes[i, 0], nodes[i, 1], nodes[j, 0], nodes[j, 1], nodes[k, 0], nodes[k, 1]
    det = _pdx(x2, x1) * (y3 - y1) - _pdx(x3, x1) * (y2 - y1)
    with np.errstate(divide="ignore", invalid="ignore"):
        w1 = (_pdx(x2, xb) * (y3 - yb) - _pdx(x3, xb) * (y2 - yb)) / det
        w2 = (_pdx(x3, xb) * (y1 - yb) - _pdx(x1, xb) * (y3 - yb)) / det
    w3 = 1.0 - w1 - w2
    cn = w1 * c[i] + w2 * c[j] + w3 * c[k]
    c[:] = np.where(hit, cn, c)
    return tid


def mixing_index(c, mass, mask=None):
    """code/StokesColor.py:391-403."""
    if mask is not None:
        c = c[mask]
        mass = mass[mask]
    W = mass.sum()
    mu = (mass @ c) / W
    var = (mass @ (c - mu) ** 2) / W
    return var / (mu * (1 - mu) + 1e-16), mu, var


# --------------------------------------------------------------------------- food tracers
def food_tracer_init(g=25, L=1.0, H=1.0, radius=0.25):
    """code/StokesFood.py:421-430."""
    xx = np.linspace(0.05, L - 0.05, g)
    yy = np.linspace(0.05, H - 0.05, g)
    gx, gy = np.meshgrid(xx, yy)
    pts = np.vstack([gx.ravel(), gy.ravel()]).T
    d = np.linalg.norm(pts - np.array([0.5, 0.5]), axis=1)
    return pts[d > radius].copy()


def locate_exact(nodes, tris, pts, chunk=4096):
    """Lowest-id triangle containing each point (weights of code/StokesColor.py:334-342
    all >= 0), -1 outside the mesh.  Brute force; small meshes only."""
    tris = np.asarray(tris)
    n = nodes[tris]
    x1, y1, x2, y2, x3, y3 = n[:, 0, 0], n[:, 0, 1], n[:, 1, 0], n[:, 1, 1], n[:, 2, 0], n[:, 2, 1]
    det = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    ok = np.abs(det) >= 1e-14
    out = np.full(len(pts), -1, dtype=np.int32)
    for s in range(0, len(pts), chunk):
        x = pts[s:s + chunk, 0][:, None]
        y = pts[s:s + chunk, 1][:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            w1 = ((x2 - x) * (y3 - y) - (x3 - x) * (y2 - y)) / det
            w2 = ((x3 - x) * (y1 - y) - (x1 - x) * (y3 - y)) / det
        w3 = 1.0 - w1 - w2
        inside = ok & (w1 >= 0) & (w2 >= 0) & (w3 >= 0)
        any_ = inside.any(axis=1)
        out[s:s + chunk] = np.where(any_, inside.argmax(axis=1), -1)
    return out


def interp_p1(nodes, tris, f, pts, tid):
    """P1 interpolation of nodal field(s) f at pts inside triangles tid; NaN where tid<0.
    Restates what matplotlib's LinearTriInterpolator does at code/StokesFood.py:482-486
    (parity unpinned: matplotlib is absent)."""
    tris = np.asarray(tris)
    t = tris[np.where(tid >= 0, tid, 0)]
    i, j, k = t[:, 0], t[:, 1], t[:, 2]
    x, y = pts[:, 0], pts[:, 1]
    x1, y1, x2, y2, x3, y3 = nodes[i, 0], nodes[i, 1], nodes[j, 0], nodes[j, 1], nodes[k, 0], nodes[k, 1]
    det = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    with np.errstate(divide="ignore", invalid="ignore"):
        w1 = ((x2 - x) * (y3 - y) - (x3 - x) * (y2 - y)) / det
        w2 = ((x3 - x) * (y1 - y) - (x1 - x) * (y3 - y)) / det
    w3 = 1.0 - w1 - w2
    f = np.asarray(f)
    if f.ndim == 1:
        val = w1 * f[i] + w2 * f[j] + w3 * f[k]
        return np.where(tid >= 0, val, np.nan)
    val = w1[:, None] * f[i] + w2[:, None] * f[j] + w3[:, None] * f[k]
    return np.where((tid >= 0)[:, None], val, np.nan)


def food_tracer_step(nodes, tris, u, pts, status, DT, L=1.0, center=(0.5, 0.5), rcap=0.28,
                     tid=None):
    """code/StokesFood.py:482-503: interpolate u, Euler move, wrap x, sticky capture.
    Points outside the mesh get NaN velocity and stay NaN (defined choice, see header)."""
    if tid is None:
        tid = locate_exact(nodes, tris, pts)
    vel = interp_p1(nodes, tris, u, pts, tid)
    pts[:, 0] += vel[:, 0] * DT
    pts[:, 1] += vel[:, 1] * DT
    pts[:, 0] = np.mod(pts[:, 0], L)
    with np.errstate(invalid="ignore"):
        d = np.sqrt((pts[:, 0] - center[0]) ** 2 + (pts[:, 1] - center[1]) ** 2)
        status[d <= rcap] = 1
    return int(status.sum())


# --------------------------------------------------------------------------- poisson / heat
def fem_system(nodes32, tris, g_centroid=None, g_const=1.0, scalar=np.float32):
    """buildFemSystem, code/poisson.py:100-146: signed 2*ADet, skip only ADet==0,
    arithmetic in the coordinate dtype (float32 in the reference), accumulated
    into float64.  ``g_centroid`` = g evaluated at the centroids (array, scalar dtype).
    Returns (rowptr, colidx, vals f64, b f64) with b already negated like the reference."""
    nd = np.asarray(nodes32, dtype=scalar)
    tris = np.asarray(tris)
    n = nd.shape[0]
    x = nd[tris, 0]
    y = nd[tris, 1]
    x1, x2, x3 = x[:, 0], x[:, 1], x[:, 2]
    y1, y2, y3 = y[:, 0], y[:, 1], y[:, 2]
    adet = x1 * y2 - x1 * y3 - x2 * y1 + x2 * y3 + x3 * y1 - x3 * y2
    keep = adet != 0
    yd = np.stack([y2 - y3, y3 - y1, y1 - y2], axis=1)
    xd = np.stack([x3 - x2, x1 - x3, x2 - x1], axis=1)
    num = yd[:, :, None] * yd[:, None, :] + xd[:, :, None] * xd[:, None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        ke = (num / (scalar(2.0) * adet)[:, None, None]).reshape(-1, 9)
    rowptr, colidx, scatter = csr_pattern(n, tris)
    vals = np.bincount(scatter[keep].ravel(), weights=ke[keep].ravel().astype(np.float64),
                       minlength=len(colidx))
    area = scalar(0.5) * adet
    g = np.asarray(g_centroid, dtype=scalar) if g_centroid is not None else scalar(g_const)
    src = (g * (area / scalar(3))).astype(np.float64)
    src = np.broadcast_to(src, adet.shape)
    b = np.bincount(tris[keep].ravel(), weights=np.repeat(src[keep], 3), minlength=n)
    return rowptr, colidx, vals, -b


def centroids(nodes, tris, scalar=np.float32):
    nd = np.asarray(nodes, dtype=scalar)
    tris = np.asarray(tris)
    x = nd[tris, 0]
    y = nd[tris, 1]
    return (x[:, 0] + x[:, 1] + x[:, 2]) / scalar(3), (y[:, 0] + y[:, 1] + y[:, 2]) / scalar(3)


def periodic_row_merge(A, b, pairs):
    """apply_periodic_bc(A,b,pairs), code/poisson.py:187-213, on a scipy LIL matrix
    (sequential, in place: later pairs see earlier pairs' edits)."""
    A = A.tolil()
    for m, s in pairs:
        A[m, :] = A[m, :] + A[s, :]
        b[m] += b[s]
        A[s, :] = 0.0
        A[s, s] = 1.0
        A[s, m] = -1.0
        b[s] = 0.0
    return A


def poisson_system(nodes32, markers, tris, g_centroid, wall_value=1.0, inner_value=0.0,
                   H=1.0, tol=1e-6, scalar=np.float32):
    """code/poisson.py:221-278: FEM system, filtered periodic row merge, Dirichlet
    identity rows (columns kept).  Returns (A csr, b, pairs_all, pairs_filtered)."""
    nd = np.asarray(nodes32, dtype=scalar)
    n = nd.shape[0]
    rowptr, colidx, vals, b = fem_system(nd, tris, g_centroid, scalar=scalar)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(n, n))
    pairs = find_boundary_pairs(nd, 1.0, tol)
    filt = filter_wall_pairs(nd, pairs, H, tol)
    A = periodic_row_merge(A, b, filt)
    y = nd[:, 1]
    is_wall = (np.abs(y - 0.0) < tol) | (np.abs(y - H) < tol)
    is_inner = markers == 2
    for i in np.where(is_wall | is_inner)[0]:
        A[i, :] = 0.0
        A[i, i] = 1.0
        b[i] = inner_value if is_inner[i] else wall_value
    return A.tocsr(), b, pairs, filt


def heat_reapply(u, nodes32, markers, pairs_all, wall_value=1.0, inner_value=0.0, H=1.0, tol=1e-6):
    """reapply_periodic_u then reapply_dirchlect_u, code/heatEq.py:282-301 (the
    periodic copy uses the UNfiltered pair list, :226,298)."""
    for m, s in pairs_all:
        u[s] = u[m]
    y = np.asarray(nodes32)[:, 1]
    is_wall = (np.abs(y - 0.0) < tol) | (np.abs(y - H) < tol)
    is_inner = markers == 2
    u[is_wall & ~is_inner] = wall_value
    u[is_inner] = inner_value
    return u


# ---- output sink (test infrastructure for fs_raster_field / fs_raster_colormap / fs_raster_points) -------
def raster_field(nodes, tris, field, width, height, extent=(0.0, 1.0, 0.0, 1.0)):
    """Restatement of what ``ax.tripcolor(triang, c, shading="gouraud")`` shows
    (code/StokesColor.py:508-511): the P1 field sampled at the pixel centres (row 0 = top), NaN
    outside the mesh.  Barycentric weights exactly as PointLocator.find computes them
    (code/StokesColor.py:334-340); a pixel centre on a shared edge takes the lower triangle id (the
    value is the same on both sides).  Plain loops over triangles: small meshes only."""
    x0, x1, y0, y1 = extent
    dx, dy = (x1 - x0) / width, (y1 - y0) / height
    xs = x0 + (np.arange(width) + 0.5) * dx
    ys = y1 - (np.arange(height) + 0.5) * dy
    img = np.full((height, width), np.nan, dtype=np.float64)
    owner = np.full((height, width), -1, dtype=np.int64)
    for t in range(len(tris) - 1, -1, -1):          # descending, so that the lowest id is written last
        a, b, c = tris[t]
        p1, p2, p3 = nodes[a], nodes[b], nodes[c]
        det = (p2[0] - p1[0]) * (p3[1] - p1[1]) - (p3[0] - p1[0]) * (p2[1] - p1[1])
        if abs(det) < 1e-14:
            continue
        lo_x, hi_x = min(p1[0], p2[0], p3[0]), max(p1[0], p2[0], p3[0])
        lo_y, hi_y = min(p1[1], p2[1], p3[1]), max(p1[1], p2[1], p3[1])
        ix = np.nonzero((xs >= lo_x - dx) & (xs <= hi_x + dx))[0]
        iy = np.nonzero((ys >= lo_y - dy) & (ys <= hi_y + dy))[0]
        if len(ix) == 0 or len(iy) == 0:
            continue
        X, Y = np.meshgrid(xs[ix], ys[iy])
        w1 = ((p2[0] - X) * (p3[1] - Y) - (p3[0] - X) * (p2[1] - Y)) / det
        w2 = ((p3[0] - X) * (p1[1] - Y) - (p1[0] - X) * (p3[1] - Y)) / det
        w3 = 1.0 - w1 - w2
        inside = (w1 >= 0) & (w2 >= 0) & (w3 >= 0)
        val = w1 * field[a] + w2 * field[b] + w3 * field[c]
        sub = np.ix_(iy, ix)
        img_sub, own_sub = img[sub], owner[sub]
        img_sub[inside] = val[inside]
        own_sub[inside] = t
        img[sub], owner[sub] = img_sub, own_sub
    return img, owner


def colorize(img, vmin, vmax, lut, background=(0, 0, 0, 255)):
    """NaN -> background, else lut[round(255 * clamp((v - vmin) / (vmax - vmin)))] in float32 like the kernel."""
    v = img.astype(np.float32)
    s = (v - np.float32(vmin)) / np.float32(np.float32(vmax) - np.float32(vmin))
    s = np.clip(s, np.float32(0), np.float32(1))
    k = np.where(np.isnan(v), 0, (s * np.float32(255) + np.float32(0.5))).astype(np.int64)
    rgba = np.empty(img.shape + (4,), dtype=np.uint8)
    rgba[..., :3] = np.asarray(lut, dtype=np.uint8)[k]
    rgba[..., 3] = 255
    rgba[np.isnan(v)] = np.asarray(background, dtype=np.uint8)
    return rgba


def splat_points(rgba, points, status, colors, radius_px, extent=(0.0, 1.0, 0.0, 1.0)):
    """Discs of radius_px pixels; overlapping discs: the largest point index is on top."""
    h, w = rgba.shape[:2]
    x0, x1, y0, y1 = extent
    r = np.float32(radius_px)
    for i, (px, py) in enumerate(points):
        if np.isnan(px) or np.isnan(py):
            continue
        cx = np.float32((px - x0) * (w / (x1 - x0)))
        cy = np.float32((y1 - py) * (h / (y1 - y0)))
        for yy in range(max(int(np.floor(cy - r)), 0), min(int(np.ceil(cy + r)), h - 1) + 1):
            for xx in range(max(int(np.floor(cx - r)), 0), min(int(np.ceil(cx + r)), w - 1) + 1):
                ddx = np.float32(xx + 0.5) - cx
                ddy = np.float32(yy + 0.5) - cy
                if ddx * ddx + ddy * ddy <= r * r:
                    k = 0 if status is None else int(status[i])
                    k = min(max(k, 0), len(colors) - 1)
                    rgba[yy, xx, :3] = colors[k]
                    rgba[yy, xx, 3] = 255
    return rgba


def draw_quiver(rgba, points, vectors, scale, half_width_px=0.6, head_frac=0.3, color=(0, 0, 0), extent=(0.0, 1.0, 0.0, 1.0)):
    """ax.quiver(x, y, u, v, angles='xy', scale_units='xy', scale=scale) (code/StokesColor.py:514-527,
    code/StokesFood.py:517-519), restated (matplotlib is absent: pinned to this restatement only): shaft from (x, y) to
    (x + u/scale, y + v/scale), two head strokes of head_frac x the shaft length at +-25 degrees; a pixel is painted when
    its centre is within half_width_px of a stroke.  float32 pixel arithmetic in the order of the device kernel."""
    f = np.float32
    h, w = rgba.shape[:2]
    x0, x1, y0, y1 = extent
    inv_dx, inv_dy, inv_s = w / (x1 - x0), h / (y1 - y0), 1.0 / scale
    hw, hf = f(half_width_px), f(head_frac)
    c, s = f(0.90630779), f(0.42261826)
    r2 = hw * hw

    def d2(px, py, ax, ay, bx, by):
        dx, dy = bx - ax, by - ay
        l2 = dx * dx + dy * dy
        t = ((px - ax) * dx + (py - ay) * dy) / l2 if l2 > 0 else f(0)
        t = min(max(t, f(0)), f(1))
        qx, qy = ax + t * dx - px, ay + t * dy - py
        return qx * qx + qy * qy

    for (px_, py_), (vx, vy) in zip(points, vectors):
        if np.isnan(px_) or np.isnan(py_) or np.isnan(vx) or np.isnan(vy):
            continue
        ax, ay = f((px_ - x0) * inv_dx), f((y1 - py_) * inv_dy)
        bx, by = f((px_ + vx * inv_s - x0) * inv_dx), f((y1 - (py_ + vy * inv_s)) * inv_dy)
        sx, sy = ax - bx, ay - by
        h1x, h1y = bx + hf * (c * sx - s * sy), by + hf * (s * sx + c * sy)
        h2x, h2y = bx + hf * (c * sx + s * sy), by + hf * (-s * sx + c * sy)
        xl, xh = min(ax, bx, h1x, h2x) - hw, max(ax, bx, h1x, h2x) + hw
        yl, yh = min(ay, by, h1y, h2y) - hw, max(ay, by, h1y, h2y) + hw
        for yy in range(max(int(np.floor(yl)), 0), min(int(np.ceil(yh)), h - 1) + 1):
            for xx in range(max(int(np.floor(xl)), 0), min(int(np.ceil(xh)), w - 1) + 1):
                px, py = f(xx + 0.5), f(yy + 0.5)
                if d2(px, py, ax, ay, bx, by) <= r2 or d2(px, py, bx, by, h1x, h1y) <= r2 or d2(px, py, bx, by, h2x, h2y) <= r2:
                    rgba[yy, xx, :3] = color
                    rgba[yy, xx, 3] = 255
    return rgba


# --------------------------------------------------------------------------- physics variants of the draft scripts
def ramp_omega(step, target=5.0, ramp_up_steps=200):
    """scripts/stokes_report.py:1156-1162."""
    return target * (step + 1) / ramp_up_steps if step < ramp_up_steps else target


def rotating_cylinder_bcu(u, nodes, wall, inner, omega, center=(0.5, 0.5)):
    """scripts/stokes_report.py:1150-1171 (the Dirichlet data the script's loop writes: walls at rest, the inner
    boundary moving tangentially with angular velocity omega)."""
    u[wall] = 0.0
    rx = nodes[inner, 0] - center[0]
    ry = nodes[inner, 1] - center[1]
    u[inner, 0] = -ry * omega
    u[inner, 1] = rx * omega


def mass_convection(nodes, tris, u, rowptr, colidx, scatter):
    """build_mass_and_convection, code/StokesColor.py:286-312, on the structural pattern: consistent mass and
    convection values, contributions added in ascending element order, (i, j) i-major (like the dense loop)."""
    tris = np.asarray(tris)
    det, yd, xd = _geom(nodes, tris)
    keep = np.abs(det) >= 1e-14
    area = 0.5 * np.abs(det)
    uc = ((u[tris[:, 0]] + u[tris[:, 1]]) + u[tris[:, 2]]) / 3.0            # u[idx].mean(axis=0)
    den = 2 * np.abs(det)
    with np.errstate(divide="ignore", invalid="ignore"):
        gx, gy = yd / den[:, None], xd / den[:, None]
    conv_j = (area / 3)[:, None] * (uc[:, 0:1] * gx + uc[:, 1:2] * gy)       # (T,3): depends on j only
    keC = np.repeat(conv_j[:, None, :], 3, axis=1).reshape(-1, 9)
    w = np.where(np.eye(3, dtype=bool), 2.0, 1.0).reshape(1, 9)
    keM = (area / 12.0)[:, None] * w
    m_vals = np.bincount(scatter[keep].ravel(), weights=keM[keep].ravel(), minlength=len(colidx))
    c_vals = np.bincount(scatter[keep].ravel(), weights=keC[keep].ravel(), minlength=len(colidx))
    return m_vals, c_vals


def dye_diffuse(c_adv, K, DT, D):
    """scripts/good_visualization2.py:704-715: explicit update with the stiffness matrix, clipped to [0, 1]
    (sign as in the script)."""
    return np.clip(c_adv + DT * D * (K @ c_adv), 0.0, 1.0)


def helmholtz_smooth(K_dense, p_raw, ref, alpha=0.01):
    """scripts/stokes_report.py:1187-1196: p = solve(I + alpha K with row / column `ref` replaced by e_ref, p_raw with
    p_raw[ref] = 0), then the mean removed."""
    n = len(p_raw)
    S = np.eye(n) + alpha * K_dense
    S[ref, :] = 0.0
    S[:, ref] = 0.0
    S[ref, ref] = 1.0
    pr = p_raw.copy()
    pr[ref] = 0.0
    p = np.linalg.solve(S, pr)
    return p - p.mean()
