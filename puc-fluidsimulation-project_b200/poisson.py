"""Poisson and heat sub-solvers (code/poisson.py, code/heatEq.py) on the device
CSR path: assembly and Krylov solves run in libfluidsim; the reference's in-place
row surgery on the dense matrix (periodic row merge, Dirichlet identity rows) is
setup-time host work on the few touched CSR rows, after which the matrix is a
device handle again.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import mesh as _mesh
from ._lib import call
from .core import CsrMatrix, Mesh


class _RowEditor:
    """Edits whole rows of a CSR matrix on the host; untouched rows are copied."""

    def __init__(self, rowptr, colidx, vals):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.n = len(rowptr) - 1
        self.rows: dict[int, dict[int, float]] = {}

    def row(self, i):
        r = self.rows.get(i)
        if r is None:
            a, b = self.rowptr[i], self.rowptr[i + 1]
            r = dict(zip(self.colidx[a:b].tolist(), self.vals[a:b].tolist()))
            self.rows[i] = r
        return r

    def finish(self):
        counts = np.diff(self.rowptr).astype(np.int64)
        for i, r in self.rows.items():
            counts[i] = len(r)
        new_ptr = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(counts, out=new_ptr[1:])
        nnz = int(new_ptr[-1])
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.float64)
        # bulk copy of untouched rows
        touched = np.zeros(self.n, dtype=bool)
        if self.rows:
            touched[list(self.rows.keys())] = True
        old_rows = np.repeat(np.arange(self.n), np.diff(self.rowptr))
        keep = ~touched[old_rows]
        dst = (new_ptr[old_rows] + (np.arange(len(old_rows)) - self.rowptr[old_rows]))[keep]
        col[dst] = self.colidx[keep]
        val[dst] = self.vals[keep]
        for i, r in self.rows.items():
            cs = sorted(r)
            a = new_ptr[i]
            col[a:a + len(cs)] = cs
            val[a:a + len(cs)] = [r[c] for c in cs]
        return new_ptr.astype(np.int32), col, val


def _swap(A: CsrMatrix, rowptr, colidx, vals):
    new = CsrMatrix.from_arrays(rowptr, colidx, vals)
    A.__dict__, new.__dict__ = new.__dict__, A.__dict__      # in-place semantics of the reference


def apply_periodic_bc(A: CsrMatrix, *args):
    """Both reference variants, chosen by arity like the reference's two files:
    ``apply_periodic_bc(A, pairs)``     penalty method,  code/StokesColor.py:206-221
    ``apply_periodic_bc(A, b, pairs)``  row merge,       code/poisson.py:187-213
    In place on the CsrMatrix handle (and on b); sequential pair order."""
    ed = _RowEditor(*A.arrays())
    if len(args) == 1:
        (pairs,) = args
        penalty = 1.0e10
        for m, s in pairs:
            m, s = int(m), int(s)
            rm, rs = ed.row(m), ed.row(s)
            rm[m] = rm.get(m, 0.0) + penalty
            rs[s] = rs.get(s, 0.0) + penalty
            rm[s] = rm.get(s, 0.0) - penalty
            rs[m] = rs.get(m, 0.0) - penalty
    else:
        b, pairs = args
        for m, s in pairs:
            m, s = int(m), int(s)
            rm, rs = ed.row(m), ed.row(s)
            for c, v in rs.items():
                rm[c] = rm.get(c, 0.0) + v
            b[m] += b[s]
            rs.clear()
            rs[s] = 1.0
            rs[m] = -1.0
            b[s] = 0.0
    _swap(A, *ed.finish())


def apply_dirichlet_rows(A: CsrMatrix, b, idx, values):
    """Row := e_i, b_i := value, columns kept.  code/poisson.py:258-278."""
    ed = _RowEditor(*A.arrays())
    for i, v in zip(np.asarray(idx).tolist(), np.broadcast_to(values, (len(idx),)).tolist()):
        r = ed.row(i)
        r.clear()
        r[i] = 1.0
        b[i] = v
    _swap(A, *ed.finish())


def add_identity_scaled(A: CsrMatrix, DT):
    """A := I + DT*A (code/heatEq.py:304-305), in place."""
    rowptr, colidx, vals = A.arrays()
    rows = np.repeat(np.arange(A.n), np.diff(rowptr))
    vals = DT * vals
    diag = rows == colidx
    if diag.sum() != A.n:
        raise ValueError("matrix has a structurally missing diagonal entry")
    vals[diag] = 1.0 + vals[diag]
    _swap(A, rowptr, colidx, vals)


class PoissonProblem:
    """code/poisson.py:216-285 as an object.  ``nodes_coords`` float32 reproduces the
    reference's float32 assembly arithmetic; float64 is the clean mode."""

    INNER_BOUNDARY_MARKER = 2
    OUTER_BOUNDARY_VALUE = 1.0
    INNER_BOUNDARY_VALUE = 0.0

    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, g_source=None, H=1.0, tol=1e-6):
        self.nodes_coords = nodes_coords
        self.markers = np.asarray(nodes_boundary_markers)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.N = len(self.markers)
        f32 = np.asarray(nodes_coords).dtype == np.float32
        if g_source is None:
            g_source = lambda x, y: 50 * np.sin(3 * y)                      # code/poisson.py:234-235
        self.mesh = Mesh(np.asarray(nodes_coords, dtype=np.float64), self.triangles, self.markers)
        vals, self.b = self.mesh.fem_system(g_source, f32=f32)
        self.A = self.mesh.matrix(vals)
        self.pairs = _mesh.find_boundary_pairs(nodes_coords, L=1.0, tol=tol)     # unfiltered, :228
        self.filtered_pairs = _mesh.filter_wall_pairs(nodes_coords, self.pairs, H=H, tol=tol)
        apply_periodic_bc(self.A, self.b, self.filtered_pairs)
        y = np.asarray(nodes_coords)[:, 1]
        is_wall = (np.abs(y - 0.0) < tol) | (np.abs(y - H) < tol)
        is_inner = self.markers == self.INNER_BOUNDARY_MARKER
        self.wall = np.where(is_wall)[0].astype(np.int32)
        self.inner = np.where(is_inner)[0].astype(np.int32)
        idx = np.where(is_wall | is_inner)[0]
        vals_bc = np.where(is_inner[idx], self.INNER_BOUNDARY_VALUE, self.OUTER_BOUNDARY_VALUE)
        apply_dirichlet_rows(self.A, self.b, idx, vals_bc)
        interior = np.setdiff1d(np.arange(self.N), idx).astype(np.int32)
        self.mesh.set_bc(self.wall, self.inner, self.filtered_pairs, interior)

    def solve(self, rtol=1e-13):
        """f = solve(A, b), code/poisson.py:283-285 (BiCGStab: the matrix is not symmetric)."""
        f, it, rr = self.A.bicgstab(self.b, rtol=rtol)
        self.iters, self.relres = it, rr
        return f


class HeatProblem(PoissonProblem):
    """code/heatEq.py:219-333: implicit Euler (I + DT*A) u+ = u with the boundary
    values re-imposed after every solve."""

    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, DT=0.02, **kw):
        super().__init__(nodes_coords, nodes_boundary_markers, triangles, **kw)
        self.DT = DT
        add_identity_scaled(self.A, DT)
        self.u = np.zeros(self.N)
        self.reapply(self.u)

    def reapply(self, u):
        """reapply_periodic_u then reapply_dirchlect_u, code/heatEq.py:282-301."""
        self.mesh.reapply_scalar_bc(u, self.pairs, self.OUTER_BOUNDARY_VALUE, self.INNER_BOUNDARY_VALUE)
        return u

    def step(self, rtol=1e-13):
        rhs = self.u.copy()                    # u + DT*b*0, code/heatEq.py:322
        self.u, it, rr = self.A.bicgstab(rhs, x0=self.u, rtol=rtol)
        self.iters = it
        self.reapply(self.u)
        return self.u


def helmholtz_smooth(K: CsrMatrix, p_raw, ref, alpha=0.01, rtol=1e-13):
    """The stabilised-pressure variant of scripts/stokes_report.py:1187-1196: solve (I + alpha K) p = p_raw with row and
    column ``ref`` replaced by the unit vector (p[ref] = 0), then remove the mean.  The system is SPD: CG on the device."""
    n = K.n
    rowptr, colidx, vals = K.arrays()
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    v = alpha * vals + (rows == colidx)
    kill = (rows == ref) | (colidx == ref)
    v = np.where(kill, (rows == colidx).astype(np.float64), v)
    S = CsrMatrix.from_arrays(rowptr, colidx, v)
    b = np.array(p_raw, dtype=np.float64)
    b[ref] = 0.0
    p, _, _ = S.cg(b, rtol=rtol)
    return p - p.mean()
