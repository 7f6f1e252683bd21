"""The non-solver kernels of the path on the 4M-triangle mesh, a few launches each (dev tool for ncu:
assembly, divergence / gradient, dye advection (k=10 locator), exact locate + tracer step)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import fluidsim_b200 as fb
nodes, markers, tris = fb.square_with_hole(2048, 1024)
m = fb.Mesh(nodes, tris, markers)
N = m.N
rng = np.random.default_rng(0)
u = torch.from_numpy(rng.standard_normal((N, 2))).cuda()
p = torch.from_numpy(rng.standard_normal(N)).cuda()
c = torch.from_numpy((nodes[:, 0] < 0.5).astype(np.float64)).cuda()
vals = torch.empty(m.nnz, dtype=torch.float64, device="cuda")
from fluidsim_b200._lib import call, ptr
import ctypes as C
for rep in range(3):
    call("fs_assemble_stiffness", m._h, ptr(vals))
    m.divergence(u)
    m.gradient(p)
    m.advect_dye(c, u * 1e-3, 0.05)
P = 4_000_000
pts = torch.from_numpy(np.stack([rng.uniform(0.02, 0.98, P), rng.uniform(0.02, 0.98, P)], 1)).cuda()
status = torch.zeros(P, dtype=torch.int32, device="cuda")
hint = torch.full((P,), -1, dtype=torch.int32, device="cuda")
for rep in range(3):
    m.tracer_step(pts, status, hint, u * 1e-3, 0.01)
torch.cuda.synchronize()
print("ok", float(vals.abs().sum()) > 0)
