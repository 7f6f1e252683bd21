"""Build libfluidsim.so (sm_100a) in-tree with nvcc.

    python puc-fluidsimulation-project_b200/build.py [--force] [--verbose]

The library is the product: there is no CPU fallback, importing the Python
mirror fails loudly when the .so is missing or cannot be loaded.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfluidsim.so")
SOURCES = ["core.cu", "topology.cu", "assembly.cu", "solver.cu", "cg_persistent.cu", "spmv_warp.cu", "spmv_sell.cu", "amg.cu", "recycle.cu", "stokes.cu", "tracer.cu", "dist.cu", "pstokes.cu"]
HEADERS = [os.path.join(CSRC, "internal.cuh"), os.path.join(CSRC, "dist.cuh"), os.path.join(CSRC, "reduce.cuh"), os.path.join(CSRC, "recycle.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "fluidsim.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact assembly / div need the reference's unfused multiply-add order
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
    "-diag-suppress", "177,550",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libfluidsim cannot be built (no CPU fallback exists)")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, "-c", os.path.join(CSRC, s), "-o", obj] + NVCC_FLAGS
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
