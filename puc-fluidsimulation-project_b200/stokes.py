"""The operator-split Stokes step and its two drivers (dye mixing, food capture).

``StokesSolver`` is the reference's module body (code/StokesColor.py:437-498 setup,
:537-575 step) as an object: the module-level globals the reference functions
read (pairs, wall_node_indices, inner_boundary_indices, B1, B2, DT, v ...) are its
attributes, the functions that read them (makeDirBCU, makePerBCU,
advect_semilagrange) are its methods.  All arithmetic runs in libfluidsim.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import mesh as _mesh
from ._lib import StokesOpts, StokesStats, call, ptr, as_f64
from .core import CsrMatrix, Mesh


class StokesSolver:
    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, B1=-2.0, B2=0.0, DT=0.05, v=0.1,
                 L=1.0, H=1.0, tol=1e-6, rtol_pressure=1e-10, rtol_visc=1e-12, precond=3, warm_start=True,
                 maxit=200000, final_div=False, bc="squirmer", omega=0.0):
        self.nodes_coords = np.ascontiguousarray(nodes_coords, dtype=np.float64)
        self.nodes_boundary_markers = np.ascontiguousarray(nodes_boundary_markers, dtype=np.int32)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.N = self.nodes_coords.shape[0]
        self.B1, self.B2, self.DT, self.v, self.L, self.H, self.tol = B1, B2, DT, v, L, H, tol
        # code/StokesColor.py:442-464
        self.all_pairs = _mesh.find_boundary_pairs(self.nodes_coords, L=L, tol=tol)
        self.pairs = _mesh.filter_wall_pairs(self.nodes_coords, self.all_pairs, H=H, tol=tol)
        (self.wall_node_indices, self.inner_boundary_indices, self.dirichlet_node_indices,
         self.interior) = _mesh.index_sets(self.nodes_coords, self.nodes_boundary_markers, H=H, tol=tol)
        self.mesh = Mesh(self.nodes_coords, self.triangles, self.nodes_boundary_markers)
        self.mesh.set_bc(self.wall_node_indices, self.inner_boundary_indices, self.pairs, self.interior)
        h = C.c_void_p()
        call("fs_stokes_create", self.mesh._h, float(DT), float(v), C.byref(h))
        self._h = h
        self.opts = StokesOpts()
        call("fs_stokes_default_opts", C.byref(self.opts))
        self.opts.rtol_pressure = rtol_pressure
        self.opts.rtol_visc = rtol_visc
        self.opts.precond = precond
        self.opts.warm_start = 1 if warm_start else 0
        self.opts.maxit = maxit
        self.opts.final_div = 1 if final_div else 0
        # bc = "rotating": the rotating-cylinder variant of scripts/stokes_report.py:1155-1171; set self.omega before
        # every step (the script ramps it: target * (step + 1) / 200)
        if bc not in ("squirmer", "rotating"):
            raise ValueError("bc must be 'squirmer' or 'rotating'")
        self.bc, self.omega = bc, float(omega)
        self.stats = StokesStats()
        self.M_lumped_diag = self.mesh.lumped_mass()
        # code/StokesColor.py:482-483
        self.u = np.zeros((self.N, 2))
        self.makeDirBCU(self.u)

    def __del__(self):
        try:
            if self._h:
                _lib.lib.fs_stokes_destroy(self._h)
        except Exception:
            pass

    # -- the reference's global-reading helpers
    def makeDirBCU(self, u):
        """code/StokesColor.py:405-427 (in place); the rotating-cylinder data for bc="rotating"."""
        if self.bc == "rotating":
            self.mesh.make_rot_bcu(u, self.omega)
        else:
            self.mesh.make_dir_bcu(u, self.B1, self.B2)

    def makePerBCU(self, u):
        """code/StokesColor.py:429-431 (in place)."""
        self.mesh.make_per_bcu(u)

    # -- matrices, for inspection
    def matrices(self):
        av, kp = C.c_void_p(), C.c_void_p()
        dof = np.empty(self.N, dtype=np.int32)
        call("fs_stokes_matrices", self._h, C.byref(av), C.byref(kp), ptr(dof))
        return CsrMatrix(av, owner=self, borrowed=True), CsrMatrix(kp, owner=self, borrowed=True), dof

    # -- one step, code/StokesColor.py:540-575
    def step(self, u=None):
        """Advance ``u`` (default: self.u) in place; numpy (host, staged) or torch cuda tensor."""
        if u is None:
            u = self.u
        self.opts.bc_mode = 1 if self.bc == "rotating" else 0
        self.opts.omega = self.omega
        call("fs_stokes_step", self._h, ptr(u, np.float64, (self.N, 2), "u"), float(self.B1), float(self.B2),
             C.byref(self.opts), C.byref(self.stats))
        return self.stats

    def get_warm_state(self):
        """CG warm-start state of the two pressure solves: last three solutions of each + history depth
        (with ``u`` the full loop state)."""
        _, kp, _ = self.matrices()
        base = 6 * kp.n + 2
        need = C.c_int64(0)
        call("fs_stokes_recycle_state", self._h, None, 0, 0, C.byref(need))
        q = np.empty(base + need.value)
        call("fs_stokes_warm_state", self._h, ptr(q), 0)
        # large systems: + the A-orthonormal basis of previous solutions the next guess is projected onto
        call("fs_stokes_recycle_state", self._h, C.c_void_p(q.ctypes.data + 8 * base), need.value, 0, C.byref(need))
        return q

    def set_warm_state(self, q):
        _, kp, _ = self.matrices()
        q = np.ascontiguousarray(q, dtype=np.float64)
        base = 6 * kp.n + 2
        if q.size < base:
            raise ValueError(f"warm state has {q.size} entries, expected at least {base}")
        call("fs_stokes_warm_state", self._h, ptr(q), 1)
        call("fs_stokes_recycle_state", self._h, C.c_void_p(q.ctypes.data + 8 * base) if q.size > base else None,
             q.size - base, 1, None)

    # -- checkpoint / resume (the reference keeps its state in module globals)
    def _state_arrays(self):
        return {"u": self.u}

    @staticmethod
    def _npz_path(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"

    def save_state(self, path):
        """Write the complete loop state (velocity, CG warm-start vectors, the two pressure fields of the
        last step, and the dye / tracer arrays of the subclasses) to ``path`` (``.npz`` is appended when
        missing, by save and load alike)."""
        p, p2 = self.pressure()
        np.savez(self._npz_path(path), warm=self.get_warm_state(), p_full=p, p2_full=p2, **self._state_arrays())

    def load_state(self, path):
        """Restore a state written by save_state; the run continues bit for bit."""
        z = np.load(self._npz_path(path))
        for k, arr in self._state_arrays().items():
            arr[...] = z[k]
        self.set_warm_state(z["warm"])
        if "p_full" in z.files:
            self.set_pressure(z["p_full"], z["p2_full"])

    def set_pressure(self, p, p2):
        """Restore the nodal pressure fields that pressure() returns (checkpoint resume)."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        p2 = np.ascontiguousarray(p2, dtype=np.float64)
        call("fs_stokes_set_pressure", self._h, ptr(p, np.float64, (self.N,)), ptr(p2, np.float64, (self.N,)))

    def pressure(self):
        p = np.empty(self.N)
        p2 = np.empty(self.N)
        call("fs_stokes_pressure", self._h, ptr(p), ptr(p2))
        return p, p2


class StokesColor(StokesSolver):
    """code/StokesColor.py: Stokes step + semi-Lagrangian dye + mixing index."""

    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, B1=-2.0, B2=0.0, DT=0.05, v=0.1, **kw):
        super().__init__(nodes_coords, nodes_boundary_markers, triangles, B1=B1, B2=B2, DT=DT, v=v, **kw)
        # code/StokesColor.py:493-498
        self.c = np.zeros(self.N)
        self.c[self.nodes_coords[:, 0] < 0.5] = 1.0
        self.inner = np.where(self.nodes_boundary_markers == 0)[0].astype(np.int32)
        self.I0, self.mu0, self.var0 = self.mesh.mixing_index(self.c, self.M_lumped_diag, self.inner)
        self.progress = 0.0

    def _state_arrays(self):
        return {"u": self.u, "c": self.c}

    def advect_semilagrange(self, c, u, DT):
        """code/StokesColor.py:347-389 (in place on c)."""
        self.mesh.advect_dye(c, u, DT)

    dye_diffusivity = 0.0      # > 0: the explicit dye diffusion of scripts/good_visualization2.py:704-715 after the advection

    def step_all(self):
        """One pass of the reference loop body, code/StokesColor.py:538-585."""
        st = self.step()
        self.advect_semilagrange(self.c, self.u, self.DT)
        if self.dye_diffusivity > 0:
            self.mesh.dye_diffuse(self.c, self.DT, self.dye_diffusivity)
        I, mu, var = self.mesh.mixing_index(self.c, self.M_lumped_diag, self.inner)
        self.progress = 1.0 - var / (self.var0 + 1e-16)
        return st, self.progress


def food_tracer_grid(grid_density=25, L=1.0, H=1.0, radius=0.25, center=(0.5, 0.5)):
    """code/StokesFood.py:421-430: g x g grid on [0.05, L-0.05] x [0.05, H-0.05] minus the body."""
    xx = np.linspace(0.05, L - 0.05, grid_density)
    yy = np.linspace(0.05, H - 0.05, grid_density)
    gx, gy = np.meshgrid(xx, yy)
    pts = np.vstack([gx.ravel(), gy.ravel()]).T
    d = np.linalg.norm(pts - np.asarray(center), axis=1)
    return np.ascontiguousarray(pts[d > radius])


class StokesFood(StokesSolver):
    """code/StokesFood.py: Stokes step + passive tracers with sticky capture."""

    SQUIRMER_RADIUS = 0.25
    CAPTURE_RADIUS = 0.25 + 0.03
    SQUIRMER_CENTER = (0.5, 0.5)

    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, B1=-2.0, B2=0.0, DT=0.01, v=1.0,
                 grid_density=25, tracer_points=None, **kw):
        super().__init__(nodes_coords, nodes_boundary_markers, triangles, B1=B1, B2=B2, DT=DT, v=v, **kw)
        self.tracer_points = (food_tracer_grid(grid_density, self.L, self.H, self.SQUIRMER_RADIUS)
                              if tracer_points is None else np.ascontiguousarray(tracer_points, dtype=np.float64))
        self.num_tracers = self.tracer_points.shape[0]
        self.tracer_status = np.zeros(self.num_tracers, dtype=np.int32)
        self._hint = np.full(self.num_tracers, -1, dtype=np.int32)
        self.num_eaten = 0

    def _state_arrays(self):
        return {"u": self.u, "tracer_points": self.tracer_points, "tracer_status": self.tracer_status,
                "_hint": self._hint}

    def tracer_step(self):
        """code/StokesFood.py:482-503."""
        self.num_eaten = self.mesh.tracer_step(self.tracer_points, self.tracer_status, self._hint, self.u, self.DT,
                                               L=self.L, center=self.SQUIRMER_CENTER, rcap=self.CAPTURE_RADIUS)
        return self.num_eaten

    def step_all(self):
        st = self.step()
        return st, self.tracer_step()


class StokesSweep:
    """BASELINE config 4 on one GPU: B squirmer configurations (B1, B2) on one mesh advanced TOGETHER -- the operators
    are shared (B1, B2 enter only through makeDirBCU, code/StokesColor.py:419), so every kernel of the step runs once
    for all configurations (fs_stokes_step_batch) and the passive-tracer step of code/StokesFood.py:482-503 once for
    all tracer sets (fs_tracer_step_batch).  Arrays carry a leading configuration axis; numpy or torch cuda tensors."""

    def __init__(self, nodes_coords, nodes_boundary_markers, triangles, configs, DT=0.01, v=1.0, grid_density=25,
                 tracer_points=None, device_arrays=False, **kw):
        self.configs = [(float(a), float(b)) for a, b in configs]
        self.B = len(self.configs)
        self.base = StokesSolver(nodes_coords, nodes_boundary_markers, triangles, B1=self.configs[0][0], B2=self.configs[0][1],
                                 DT=DT, v=v, **kw)
        self.N, self.DT, self.mesh = self.base.N, DT, self.base.mesh
        self.b12 = np.ascontiguousarray(self.configs, dtype=np.float64)
        h = C.c_void_p()
        call("fs_stokes_batch_create", self.base._h, self.B, C.byref(h))
        self._h = h
        self.u = np.zeros((self.B, self.N, 2))
        for c, (B1, B2) in enumerate(self.configs):                   # code/StokesColor.py:482-483 per configuration
            self.mesh.make_dir_bcu(self.u[c], B1, B2)
        pts = (food_tracer_grid(grid_density, self.base.L, self.base.H, StokesFood.SQUIRMER_RADIUS)
               if tracer_points is None else np.ascontiguousarray(tracer_points, dtype=np.float64))
        self.num_tracers = pts.shape[0]
        if device_arrays:
            import torch
            self.u = torch.from_numpy(self.u).cuda()
            one = torch.from_numpy(pts).cuda()
            self.tracer_points = one.unsqueeze(0).repeat(self.B, 1, 1).contiguous()
            self.tracer_status = torch.zeros((self.B, self.num_tracers), dtype=torch.int32, device="cuda")
            self._hint = torch.full((self.B, self.num_tracers), -1, dtype=torch.int32, device="cuda")
        else:
            self.tracer_points = np.ascontiguousarray(np.broadcast_to(pts, (self.B,) + pts.shape))
            self.tracer_status = np.zeros((self.B, self.num_tracers), dtype=np.int32)
            self._hint = np.full((self.B, self.num_tracers), -1, dtype=np.int32)
        self.iters = np.zeros((self.B, 3), dtype=np.int32)
        self.num_eaten = np.zeros(self.B, dtype=np.int64)

    def __del__(self):
        try:
            if self._h:
                _lib.lib.fs_stokes_batch_destroy(self._h)
        except Exception:
            pass

    def step(self, want_iters=False):
        """One flow step of every configuration (code/StokesColor.py:540-575), in place on self.u."""
        call("fs_stokes_step_batch", self._h, ptr(self.u, np.float64, (self.B, self.N, 2), "u"), ptr(self.b12),
             C.byref(self.base.opts), ptr(self.iters) if want_iters else None)
        return self.iters if want_iters else None

    def tracer_step(self):
        """code/StokesFood.py:482-503 for every configuration's tracer set; returns the eaten counts (B,)."""
        s = StokesFood
        call("fs_tracer_step_batch", self.mesh._h, self.B, ptr(self.tracer_points, np.float64, (self.B, self.num_tracers, 2), "pts"),
             ptr(self.tracer_status, np.int32, (self.B, self.num_tracers), "status"),
             ptr(self._hint, np.int32, (self.B, self.num_tracers), "hint"), self.num_tracers,
             ptr(self.u, np.float64, (self.B, self.N, 2), "u"), float(self.DT), float(self.base.L), float(s.SQUIRMER_CENTER[0]),
             float(s.SQUIRMER_CENTER[1]), float(s.CAPTURE_RADIUS), ptr(self.num_eaten))
        return self.num_eaten

    def step_all(self):
        self.step()
        return self.tracer_step()
