"""ctypes wrapper of oracle/_ref/libcgport.so (oracle/cg_port.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libcgport.so")


def load():
    src = os.path.join(_HERE, "cg_port.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):     # missing or older than its source
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(_LIB)
    lib.cgport_threads.restype = C.c_int
    lib.cgport_spmv.argtypes = [C.c_int64] + [C.c_void_p] * 5
    lib.cgport_multidot.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.cgport_comb.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.cgport_cg.restype = C.c_int
    lib.cgport_cg.argtypes = [C.c_int64] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return lib


def threads():
    return load().cgport_threads()


def spmv(rowptr, colidx, vals, x):
    y = np.empty_like(x)
    load().cgport_spmv(len(rowptr) - 1, rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


def cg(rowptr, colidx, vals, b, x0=None, rtol=1e-10, maxit=100000, jacobi=True, project_mean=False):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
    rr = C.c_double(0)
    it = load().cgport_cg(len(rowptr) - 1, rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data, b.ctypes.data,
                          x.ctypes.data, rtol, maxit, 1 if jacobi else 0, 1 if project_mean else 0, C.byref(rr))
    return x, it, rr.value


def multidot(X, k, a):
    """X[:k] @ a with the OpenMP loops of the port (X: 2-D C-contiguous)."""
    out = np.zeros(max(k, 1))
    if k:
        load().cgport_multidot(X.shape[1], k, X.ctypes.data, X.shape[1], a.ctypes.data, out.ctypes.data)
    return out[:k]


def comb(X, k, coef, beta=0.0, vin=None, out=None):
    """beta * vin + coef @ X[:k]"""
    n = X.shape[1]
    if out is None:
        out = np.empty(n)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    load().cgport_comb(n, k, X.ctypes.data, n, coef.ctypes.data, float(beta), None if vin is None else vin.ctypes.data,
                       out.ctypes.data)
    return out
