"""PCG iteration counts at a tight tolerance (rtol 1e-13, the 4M-triangle oracle test) for the preconditioner variants."""
import os, sys, subprocess, json
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, fluidsim_b200 as fb
    c, mk, t = fb.square_with_hole(2048, 1024)
    sim = fb.StokesSolver(c, mk, t, precond=fb.PRECOND_AMG, B1=-2.0, B2=-5.0, DT=0.05, v=0.1, rtol_pressure=float(sys.argv[2]), rtol_visc=1e-14)
    its = []
    for _ in range(3):
        st = sim.step()
        its.append((st.iters_p1, st.iters_p2, float(st.relres_p1)))
    print(json.dumps(its))
else:
    for rtol in ("1e-13", "1e-12"):
        for env in ({}, {"FS_PCG_R32": "0"}, {"FS_SELL_PACK": "0"}, {"FS_SELL_PACK": "0", "FS_PCG_R32": "0"}, {"FS_STOKES_RECYCLE": "0"}):
            e = dict(os.environ); e.update(env)
            r = subprocess.run([sys.executable, __file__, "child", rtol], env=e, capture_output=True, text=True)
            print(rtol, env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
