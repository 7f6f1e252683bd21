"""Device handles (Mesh, CsrMatrix) and the reference's function-style API on top
of them.  Every function keeps the reference's name, argument order and return
shape; dense (N,N) returns become CsrMatrix handles with
.toarray() / .tocsr() for inspection.
"""
from __future__ import annotations

import ctypes as C
import zlib

import numpy as np

from . import _lib
from . import mesh as _mesh
from ._lib import call, ptr, as_f64, as_i32

PRECOND_NONE, PRECOND_JACOBI, PRECOND_AMG, PRECOND_AUTO = 0, 1, 2, 3


def _empty_like_buf(ref, n, dtype):
    """Output buffer of n elements next to `ref`: numpy for host inputs, torch cuda
    tensor for device inputs."""
    if _lib._is_torch(ref):
        import torch
        td = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
        return torch.empty(n, dtype=td, device=ref.device)
    return np.empty(n, dtype=dtype)


class CsrMatrix:
    """Opaque device CSR matrix (fs_csr).  Stands in for the reference's dense
    numpy (N,N) matrices."""

    def __init__(self, handle, owner=None, borrowed=False):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self._owner = owner          # keeps the mesh / stokes state alive
        self._borrowed = borrowed
        n, nnz = C.c_int64(0), C.c_int64(0)
        call("fs_csr_sizes", self._h, C.byref(n), C.byref(nnz))
        self.n, self.nnz = n.value, nnz.value
        self.shape = (self.n, self.n)

    @classmethod
    def from_arrays(cls, rowptr, colidx, vals):
        rowptr, colidx, vals = as_i32(rowptr), as_i32(colidx), as_f64(vals)
        n, nnz = len(rowptr) - 1, len(colidx)
        h = C.c_void_p()
        call("fs_csr_create", n, nnz, ptr(rowptr, np.int32), ptr(colidx, np.int32), ptr(vals, np.float64), C.byref(h))
        return cls(h)

    @classmethod
    def from_scipy(cls, A):
        A = A.tocsr()
        A.sort_indices()
        return cls.from_arrays(A.indptr, A.indices, A.data)

    def __del__(self):
        try:
            if not self._borrowed and self._h:
                _lib.lib.fs_csr_destroy(self._h)
        except Exception:
            pass

    def arrays(self):
        rowptr = np.empty(self.n + 1, dtype=np.int32)
        colidx = np.empty(self.nnz, dtype=np.int32)
        vals = np.empty(self.nnz, dtype=np.float64)
        call("fs_csr_get", self._h, ptr(rowptr), ptr(colidx), ptr(vals))
        return rowptr, colidx, vals

    def tocsr(self):
        import scipy.sparse as sp
        rowptr, colidx, vals = self.arrays()
        return sp.csr_matrix((vals, colidx, rowptr), shape=self.shape)

    def toarray(self):
        return self.tocsr().toarray()

    def matvec(self, x):
        x = as_f64(x)
        y = _empty_like_buf(x, self.n, np.float64)
        call("fs_spmv", self._h, ptr(x, np.float64, (self.n,)), ptr(y))
        return y

    __matmul__ = matvec

    def precond_apply(self, r):
        """z = M^-1 r: one V(1,1) cycle of the AMG hierarchy behind PRECOND_AMG (built on first use)."""
        r = as_f64(r)
        z = _empty_like_buf(r, self.n, np.float64)
        call("fs_precond_apply", self._h, ptr(r, np.float64, (self.n,)), ptr(z))
        return z

    def cg(self, b, x0=None, rtol=1e-10, maxit=100000, precond=PRECOND_JACOBI, project_mean=False):
        b = as_f64(b)
        nrhs = 1 if b.ndim == 1 else b.shape[1]
        if x0 is None:
            x = np.zeros_like(b) if not _lib._is_torch(b) else b.new_zeros(b.shape)
        else:
            x = as_f64(x0).copy() if not _lib._is_torch(x0) else x0.clone()
        it, rr = C.c_int(0), C.c_double(0)
        call("fs_cg", self._h, ptr(b, np.float64), ptr(x, np.float64), nrhs, rtol, maxit, precond,
             1 if project_mean else 0, C.byref(it), C.byref(rr))
        return x, it.value, rr.value

    def bicgstab(self, b, x0=None, rtol=1e-10, maxit=100000, precond=PRECOND_JACOBI):
        b = as_f64(b)
        if x0 is None:
            x = np.zeros_like(b) if not _lib._is_torch(b) else b.new_zeros(b.shape)
        else:
            x = as_f64(x0).copy() if not _lib._is_torch(x0) else x0.clone()
        it, rr = C.c_int(0), C.c_double(0)
        call("fs_bicgstab", self._h, ptr(b, np.float64, (self.n,)), ptr(x, np.float64), rtol, maxit, precond,
             C.byref(it), C.byref(rr))
        return x, it.value, rr.value


def solve(A, b, symmetric=None, rtol=1e-12, x0=None):
    """Stand-in for ``np.linalg.solve(A, b)`` (code/StokesColor.py:544, code/heatEq.py:323)
    on a CsrMatrix: CG when ``symmetric`` else BiCGStab, Jacobi preconditioned."""
    if not isinstance(A, CsrMatrix):
        raise TypeError("solve() needs a CsrMatrix handle (dense matrices are the reference's path)")
    if symmetric:
        return A.cg(b, x0=x0, rtol=rtol)[0]
    return A.bicgstab(b, x0=x0, rtol=rtol)[0]


class Mesh:
    """Device mesh handle (fs_mesh): coordinates, triangles, structural CSR pattern,
    scatter map and incidence lists live on the GPU."""

    def __init__(self, nodes, triangles, markers=None):
        nodes = as_f64(nodes)
        triangles = as_i32(triangles)
        if nodes.ndim != 2 or nodes.shape[1] != 2:
            raise ValueError("nodes must have shape (N,2)")
        if triangles.ndim != 2 or triangles.shape[1] != 3:
            raise ValueError("triangles must have shape (T,3)")
        mk = as_i32(markers) if markers is not None else None
        h = C.c_void_p()
        call("fs_mesh_create", ptr(nodes, np.float64), nodes.shape[0], ptr(triangles, np.int32), triangles.shape[0],
             ptr(mk, np.int32, (nodes.shape[0],)) if mk is not None else None, C.byref(h))
        self._h = h
        n, t, nnz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        call("fs_mesh_sizes", h, C.byref(n), C.byref(t), C.byref(nnz))
        self.N, self.T, self.nnz = n.value, t.value, nnz.value
        self._bc = None

    def __del__(self):
        try:
            if self._h:
                _lib.lib.fs_mesh_destroy(self._h)
        except Exception:
            pass

    # -- pattern
    def csr_pattern(self):
        rowptr = np.empty(self.N + 1, dtype=np.int32)
        colidx = np.empty(self.nnz, dtype=np.int32)
        call("fs_csr_pattern", self._h, ptr(rowptr), ptr(colidx))
        return rowptr, colidx

    def scatter_map(self):
        s = np.empty((self.T, 9), dtype=np.int32)
        call("fs_scatter_map", self._h, ptr(s))
        return s

    # -- assembly
    def stiffness_values(self):
        v = np.empty(self.nnz, dtype=np.float64)
        call("fs_assemble_stiffness", self._h, ptr(v))
        return v

    def stiffness(self):
        return self.matrix(self.stiffness_values())

    def matrix(self, vals):
        vals = as_f64(vals)
        h = C.c_void_p()
        call("fs_csr_from_mesh", self._h, ptr(vals, np.float64, (self.nnz,)), C.byref(h))
        return CsrMatrix(h, owner=self)

    def lumped_mass(self):
        m = np.empty(self.N, dtype=np.float64)
        call("fs_lumped_mass", self._h, ptr(m))
        return m

    def centroids(self, f32=False):
        cx = np.empty(self.T, dtype=np.float64)
        cy = np.empty(self.T, dtype=np.float64)
        call("fs_centroids", self._h, 1 if f32 else 0, ptr(cx), ptr(cy))
        return cx, cy

    def fem_system(self, g_source=1.0, f32=False):
        """(vals on the pattern, -b) of buildFemSystem, code/poisson.py:100-146."""
        g_arr, g_const = None, 0.0
        if callable(g_source):
            cx, cy = self.centroids(f32=f32)
            if f32:
                cx, cy = cx.astype(np.float32), cy.astype(np.float32)
            g_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(g_source(cx, cy), dtype=np.float64), (self.T,)))
        else:
            g_const = float(g_source)
        vals = np.empty(self.nnz, dtype=np.float64)
        b = np.empty(self.N, dtype=np.float64)
        call("fs_assemble_fem", self._h, 1 if f32 else 0, ptr(g_arr, np.float64) if g_arr is not None else None,
             g_const, ptr(vals), ptr(b))
        return vals, b

    # -- operators
    def divergence(self, u):
        u = as_f64(u)
        out = _empty_like_buf(u, self.N, np.float64)
        call("fs_divergence", self._h, ptr(u, np.float64, (self.N, 2), "u"), ptr(out))
        return out

    def gradient(self, p):
        p = as_f64(p)
        gx = _empty_like_buf(p, self.N, np.float64)
        gy = _empty_like_buf(p, self.N, np.float64)
        call("fs_gradient", self._h, ptr(p, np.float64, (self.N,), "p"), ptr(gx), ptr(gy))
        return gx, gy

    # -- boundary conditions
    def set_bc(self, wall, inner, pairs, interior):
        wall, inner, interior = as_i32(wall), as_i32(inner), as_i32(interior)
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        call("fs_bc_set", self._h, ptr(wall), len(wall), ptr(inner), len(inner), ptr(pr), len(pr),
             ptr(interior), len(interior))
        self._bc = dict(wall=wall, inner=inner, pairs=[(int(a), int(b)) for a, b in pr], interior=interior)

    def make_per_bcu(self, u):
        call("fs_make_per_bcu", self._h, ptr(u, np.float64, (self.N, 2), "u"))

    def make_dir_bcu(self, u, B1, B2):
        call("fs_make_dir_bcu", self._h, ptr(u, np.float64, (self.N, 2), "u"), float(B1), float(B2))

    def make_rot_bcu(self, u, omega, center=(0.5, 0.5)):
        """Rotating inner cylinder, scripts/stokes_report.py:1155-1171 (in place)."""
        call("fs_make_rot_bcu", self._h, ptr(u, np.float64, (self.N, 2), "u"), float(omega), float(center[0]), float(center[1]))

    def mass_convection(self, u):
        """build_mass_and_convection(nodes, triangles, u), code/StokesColor.py:286-312 -> (M, C) as CsrMatrix handles."""
        u = as_f64(u)
        mv = np.empty(self.nnz, dtype=np.float64)
        cv = np.empty(self.nnz, dtype=np.float64)
        call("fs_assemble_mass_convection", self._h, ptr(u, np.float64, (self.N, 2), "u"), ptr(mv), ptr(cv))
        return self.matrix(mv), self.matrix(cv)

    def dye_diffuse(self, c, DT, D):
        """scripts/good_visualization2.py:704-715: c <- clip(c + DT*D*(A_stiffness @ c), 0, 1), in place."""
        if getattr(self, "_K", None) is None:
            self._K = self.stiffness()
        call("fs_dye_diffuse", self._K._h, ptr(c, np.float64, (self.N,), "c"), float(DT), float(D))

    def reapply_scalar_bc(self, u, pairs_all, wall_value, inner_value):
        pr = np.ascontiguousarray(np.asarray(pairs_all, dtype=np.int32).reshape(-1, 2))
        call("fs_reapply_scalar_bc", self._h, ptr(u, np.float64, (self.N,), "u"), ptr(pr), len(pr),
             float(wall_value), float(inner_value))

    # -- tracers / dye
    def locate(self, pts):
        pts = as_f64(pts)
        P = pts.shape[0]
        ids = _empty_like_buf(pts, P, np.int32)
        call("fs_locate", self._h, ptr(pts, np.float64, (P, 2), "pts"), P, ptr(ids))
        return ids

    def locate_exact(self, pts, hint=None):
        pts = as_f64(pts)
        P = pts.shape[0]
        if hint is None:
            hint = _empty_like_buf(pts, P, np.int32)
            hint[:] = -1
        call("fs_locate_exact", self._h, ptr(pts, np.float64, (P, 2), "pts"), P, ptr(hint, np.int32, (P,), "hint"))
        return hint

    def advect_dye(self, c, u, DT, want_ids=False):
        ids = _empty_like_buf(c, self.N, np.int32) if want_ids else None
        call("fs_advect_dye", self._h, ptr(c, np.float64, (self.N,), "c"), ptr(as_f64(u), np.float64, (self.N, 2), "u"),
             float(DT), ptr(ids) if ids is not None else None)
        return ids

    def mixing_index(self, c, mass, mask=None):
        out = np.zeros(3, dtype=np.float64)
        mk = as_i32(mask) if mask is not None else None
        call("fs_mixing_index", self._h, ptr(as_f64(c), np.float64, (self.N,)), ptr(as_f64(mass), np.float64, (self.N,)),
             ptr(mk) if mk is not None else None, len(mk) if mk is not None else 0, ptr(out))
        return float(out[0]), float(out[1]), float(out[2])

    def tracer_step(self, pts, status, hint, u, DT, L=1.0, center=(0.5, 0.5), rcap=0.28):
        P = pts.shape[0]
        eaten = C.c_int64(0)
        call("fs_tracer_step", self._h, ptr(pts, np.float64, (P, 2), "pts"), ptr(status, np.int32, (P,), "status"),
             ptr(hint, np.int32, (P,), "hint"), P, ptr(as_f64(u), np.float64, (self.N, 2), "u"), float(DT), float(L),
             float(center[0]), float(center[1]), float(rcap), C.byref(eaten))
        return eaten.value


# ---------------------------------------------------------------------------------------
# function-style API with the reference's signatures.  Meshes are cached by content so
# repeated calls with the same (nodes, triangles) arrays reuse the device handle.
_mesh_cache: dict = {}


def _fingerprint(a):
    a = np.ascontiguousarray(a)
    return (a.shape, a.dtype.str, zlib.crc32(memoryview(a).cast("B")))


def mesh_for(nodes, triangles) -> Mesh:
    key = (_fingerprint(nodes), _fingerprint(triangles))
    m = _mesh_cache.get(key)
    if m is None:
        if len(_mesh_cache) >= 8:
            _mesh_cache.pop(next(iter(_mesh_cache)))
        m = Mesh(nodes, triangles)
        _mesh_cache[key] = m
    return m


def buildStiffnessMatrix(nodes, triangles, g_source=1.0):
    """-> (A, -B): A is a CsrMatrix, B the zero load vector.  code/StokesColor.py:98-128."""
    m = mesh_for(nodes, triangles)
    return m.stiffness(), -np.zeros(m.N, dtype=np.float64)


def buildFemSystem(nodes, triangles, g_source=1.0):
    """-> (A, -B).  code/poisson.py:100-146; float32 nodes reproduce the reference's
    float32 element arithmetic."""
    f32 = np.asarray(nodes).dtype == np.float32
    m = mesh_for(nodes, triangles)
    vals, b = m.fem_system(g_source, f32=f32)
    return m.matrix(vals), b


def buildLumpedMassMatrix(nodes_coords, triangles):
    """code/StokesColor.py:266-284."""
    return mesh_for(nodes_coords, triangles).lumped_mass()


def calculate_divergence(nodes, triangles, u_star):
    """code/StokesColor.py:130-165."""
    return mesh_for(nodes, triangles).divergence(u_star)


def calculate_gradiant(nodes, triangles, p_scalar):
    """code/StokesColor.py:224-263 (the reference's spelling)."""
    return mesh_for(nodes, triangles).gradient(p_scalar)


class PointLocator:
    """code/StokesColor.py:314-345."""

    def __init__(self, nodes, triangles):
        self.nodes = nodes
        self.triangles = triangles
        self._mesh = mesh_for(nodes, triangles)

    def find(self, x, y, k=10):
        if k != 10:
            raise ValueError("the device locator implements the reference's k=10")
        tid = int(self._mesh.locate(np.array([[x, y]], dtype=np.float64))[0])
        return None if tid < 0 else tid

    def find_many(self, pts):
        return self._mesh.locate(pts)


def build_mass_and_convection(nodes, triangles, u):
    """code/StokesColor.py:286-312 -> (M, C) CsrMatrix handles (the reference returns dense matrices)."""
    return mesh_for(nodes, triangles).mass_convection(u)


def mixing_index(c, mass, mask=None, mesh: Mesh | None = None):
    """code/StokesColor.py:391-403 -> (I, mu, var)."""
    if mesh is None:
        raise ValueError("mixing_index needs mesh= (the device handle that owns the reduction)")
    return mesh.mixing_index(c, mass, mask)
