"""CPU prototype (oracle restatement, scipy): PCG iteration counts of the folded V-cycle with its operators rounded to fp32 / bf16 / fp16."""
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp
from oracle import restated as R
from oracle.amg_cpu import Hier, pcg
import importlib.util
spec = importlib.util.spec_from_file_location("hostmesh", "puc-fluidsimulation-project_b200/hostmesh.py")
hm = importlib.util.module_from_spec(spec); spec.loader.exec_module(hm)

def rnd(M, mode):
    M = M.tocsr().copy()
    if mode == "f64": return M
    d = M.data.astype(np.float32)
    if mode == "bf16":
        u = d.view(np.uint32)
        u = ((u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000).astype(np.uint32)   # RNE
        d = u.view(np.float32)
    elif mode == "f16":
        d = d.astype(np.float16).astype(np.float32)
    M.data = d.astype(np.float64)
    return M

class Folded:
    def __init__(self, H, mode, omega=2/3, big=100000):
        self.H = H; self.ops = []
        for l in range(len(H.A) - 1):
            A, P, D = H.A[l], H.P[l], H.D[l]
            Dinv = sp.diags(1.0 / D)
            Pt = (P - omega * Dinv @ (A @ P)).tocsr()
            G = (omega * Dinv @ (2 * sp.identity(A.shape[0]) - omega * A @ Dinv)).tocsr()
            m = mode if A.shape[0] > big else "f32"
            Rt = rnd(Pt.T.tocsr(), m); G = rnd(G, m); Pt = rnd(Pt, m)
            self.ops.append((Rt, G, Pt))
    def __call__(self, r):
        bs = [r]
        for (Rt, G, Pt) in self.ops: bs.append(Rt @ bs[-1])
        x = self.H.Cinv @ bs[-1]
        for l in range(len(self.ops) - 1, -1, -1):
            Rt, G, Pt = self.ops[l]
            x = G @ bs[l] + Pt @ x
        return x

nt, nr = int(sys.argv[1]), int(sys.argv[2])
nodes, markers, tris = hm.square_with_hole(nt, nr)
pairs = R.filter_wall_pairs(nodes, R.find_boundary_pairs(nodes, 1.0, 1e-6), 1.0, 1e-6)
ps = R.PressureSystem(nodes, tris, pairs)
A = ps.K.tocsr()
print("n", A.shape[0])
H = Hier(A)
print(H.sizes())
rng = np.random.default_rng(0)
b = rng.standard_normal(A.shape[0])
# smooth rhs as well
xy = nodes
bs = np.bincount(ps.dof, weights=np.sin(3*xy[:,0])*np.cos(2*xy[:,1]), minlength=ps.nd)
for mode in ["f64", "f32", "bf16", "f16"]:
    M = Folded(H, mode)
    print(mode, "iters random", pcg(A, b, M), "smooth", pcg(A, bs, M), flush=True)
