"""Literal tier of the oracle: the reference's own functions, executed unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it.  This module additionally needs the reference
checkout at ``/root/reference`` and therefore only runs in the authoring
container (it produces the committed fixtures under ``tests/golden/`` via
``oracle/gen_golden.py``); it never runs on the GPU box.

It does NOT copy reference source: it parses the reference scripts with
``ast``, keeps only their ``def`` / ``class`` / ``import`` nodes (the scripts
have no ``__main__`` guard, importing them would run 6000 steps and open a
window) and ``exec``s those into a fresh namespace, with ``jax`` and
``matplotlib`` (absent here) replaced by ``MagicMock`` modules for the duration
of the exec.  Recipe: SURVEY.md Appendix C.

Reference entry points obtained this way:
  code/StokesColor.py:54-431  readNode, readEle, buildStiffnessMatrix,
      calculate_divergence, find_boundary_pairs, apply_periodic_bc (penalty),
      calculate_gradiant, buildLumpedMassMatrix, PointLocator,
      advect_semilagrange, mixing_index, makeDirBCU, makePerBCU
  code/poisson.py:27-213      readNode (fp32), readEle, readPoly,
      buildFemSystem, find_boundary_pairs, apply_periodic_bc (row merge)
  code/heatEq.py:282-301      reapply_dirchlect_u, reapply_periodic_u
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("FLUIDSIM_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "jax", "jax.numpy", "jax.experimental", "jax.experimental.sparse",
    "matplotlib", "matplotlib.tri", "matplotlib.pyplot", "matplotlib.colors",
    "matplotlib.animation",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "code", "StokesColor.py"))


def _install_stubs():
    saved = {k: sys.modules.get(k) for k in _STUBS}
    for name in _STUBS:
        sys.modules[name] = MagicMock(name=name)
    sys.modules["matplotlib.pyplot"].subplots = lambda *a, **k: (MagicMock(), MagicMock())
    return saved


def _remove_stubs(saved):
    # a MagicMock 'jax' left in sys.modules breaks scipy's array-API sniffing
    for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "matplotlib")]:
        if isinstance(sys.modules[k], MagicMock):
            del sys.modules[k]
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v


def load_functions(script: str = "StokesColor.py") -> dict:
    """Return a namespace holding the reference script's functions/classes."""
    path = os.path.join(REFERENCE_ROOT, "code", script)
    with open(path) as fh:
        tree = ast.parse(fh.read())
    keep = [n for n in tree.body
            if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.Import, ast.ImportFrom))]
    ns: dict = {"__name__": "reference_" + script.replace(".", "_")}
    saved = _install_stubs()
    try:
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    finally:
        _remove_stubs(saved)
    return ns


def mesh_paths(name: str):
    """name like 'mesh5.1' -> (.node, .ele) under the reference's resources/."""
    base = os.path.join(REFERENCE_ROOT, "resources", name)
    return base + ".node", base + ".ele"


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class LiteralStokes:
    """The reference's StokesColor/StokesFood module body, re-enacted with the
    reference's own functions (code/StokesColor.py:437-498 setup, :537-586 loop).

    The statements between the function calls are the reference's script body
    restated (it cannot be exec'd as a whole without plotting); every numeric
    operation is delegated to the reference functions loaded by load_functions.
    """

    def __init__(self, mesh: str, B1=-2.0, B2=0.0, DT=0.05, v=0.1, H=1.0, tol=1e-6,
                 solver=None):
        ns = load_functions("StokesColor.py")
        self.ns = ns
        self.solve = solver or np.linalg.solve
        node_path, ele_path = mesh_paths(mesh)
        with quiet():
            nodes, markers = ns["readNode"](node_path)
            tris = ns["readEle"](ele_path)
            all_pairs = ns["find_boundary_pairs"](nodes, L=1.0)
        N = nodes.shape[0]
        pairs = []
        for m, s in all_pairs:                                   # :449-457
            my = nodes[m, 1]
            if not (abs(my - 0.0) < tol or abs(my - H) < tol):
                pairs.append((int(m), int(s)))
        wall = np.where(np.isclose(nodes[:, 1], 0.0, atol=tol)
                        | np.isclose(nodes[:, 1], H, atol=tol))[0]  # :461
        inner_b = np.where(markers == 2)[0]                      # :462
        dirichlet = np.union1d(wall, inner_b)
        interior = np.setdiff1d(np.arange(N), dirichlet)
        with quiet():
            K, _ = ns["buildStiffnessMatrix"](nodes, tris, g_source=0.0)
            M = ns["buildLumpedMassMatrix"](nodes, tris)
        A_visc = np.eye(N) + DT * v * K                           # :471-475
        A_visc[dirichlet, :] = 0.0
        A_visc[:, dirichlet] = 0.0
        A_visc[dirichlet, dirichlet] = 1.0
        A_pressure = K / (M[:, None] + 1e-12)                     # :478-479
        ns["apply_periodic_bc"](A_pressure, pairs)
        ns.update(dict(
            wall_node_indices=wall, inner_boundary_indices=inner_b,
            nodes_coords=nodes, pairs=pairs, OUTER_BOUNDARY_VALUE=[0.0, 0.0],
            B1=B1, B2=B2, N=N, triangles=tris))
        self.nodes, self.markers, self.tris = nodes, markers, tris
        self.N, self.pairs, self.all_pairs = N, pairs, [(int(a), int(b)) for a, b in all_pairs]
        self.wall, self.inner_b, self.dirichlet, self.interior = wall, inner_b, dirichlet, interior
        self.K, self.M, self.A_visc, self.A_pressure = K, M, A_visc, A_pressure
        self.DT, self.v = DT, v
        self.u = np.zeros((N, 2))
        ns["makeDirBCU"](self.u)                                  # :482-483
        # dye (:493-498)
        self.c = np.zeros(N)
        self.c[nodes[:, 0] < 0.5] = 1.0
        self.inner_mask = np.where(markers == 0)[0]
        self.I0, self.mu0, self.var0 = ns["mixing_index"](self.c, M, mask=self.inner_mask)
        ns["point_locator"] = ns["PointLocator"](nodes, tris)
        self.p = np.zeros(N)
        self.p2 = np.zeros(N)

    def flow_step(self):
        """code/StokesColor.py:540-575."""
        ns, DT, nodes, tris = self.ns, self.DT, self.nodes, self.tris
        u = self.u
        u_star = np.zeros((self.N, 2))
        u_star[:, 0] = self.solve(self.A_visc, u[:, 0].copy())
        u_star[:, 1] = self.solve(self.A_visc, u[:, 1].copy())
        ns["makePerBCU"](u_star)
        ns["makeDirBCU"](u_star)
        div_u_star = ns["calculate_divergence"](nodes, tris, u_star)
        p = self.solve(self.A_pressure, -(1.0 / DT) * div_u_star)
        gx, gy = ns["calculate_gradiant"](nodes, tris, p)
        u[:, 0] = u_star[:, 0] - DT * gx
        u[:, 1] = u_star[:, 1] - DT * gy
        ns["makePerBCU"](u)
        ns["makeDirBCU"](u)
        div_u = ns["calculate_divergence"](nodes, tris, u)
        p2 = self.solve(self.A_pressure, -(1.0 / DT) * div_u)
        g2x, g2y = ns["calculate_gradiant"](nodes, tris, p2)
        u[self.interior, 0] -= DT * g2x[self.interior]
        u[self.interior, 1] -= DT * g2y[self.interior]
        self.p, self.p2 = p, p2
        self.div_u_star = div_u_star
        self.final_div = ns["calculate_divergence"](nodes, tris, u)

    def dye_step(self):
        """code/StokesColor.py:579-585."""
        ns = self.ns
        ns["advect_semilagrange"](self.c, self.u, self.DT)
        I, mu, var = ns["mixing_index"](self.c, self.M, mask=self.inner_mask)
        self.progress = 1.0 - var / (self.var0 + 1e-16)
        return self.progress
