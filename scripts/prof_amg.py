"""Short AMG-PCG solve for ncu (dev tool): 4M-triangle pressure operator, a few iterations."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib
nodes, markers, tris = fb.square_with_hole(2048, 1024)
sim = fb.StokesSolver(nodes, markers, tris, B1=-2.0, B2=-5.0)
_, kp, _ = sim.matrices()
b = torch.randn(kp.n, dtype=torch.float64, device="cuda"); b -= b.mean()
xs = torch.zeros_like(b)
it, rr = C.c_int(), C.c_double()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for rep in range(2):
    xs.zero_()
    _lib.lib.fs_cg(kp._h, C.c_void_p(b.data_ptr()), C.c_void_p(xs.data_ptr()), 1, 1e-30, n, 2, 1, C.byref(it), C.byref(rr))
torch.cuda.synchronize()
print("ok", it.value)
