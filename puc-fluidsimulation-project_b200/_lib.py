"""ctypes binding of libfluidsim.so (include/fluidsim.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc
(build.py); if that fails, or the library cannot be loaded, import fails loudly.
Compute calls on a machine without a CUDA device return FS_ERR_CUDA and raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

c_i64 = C.c_int64
c_i32 = C.c_int32
c_dbl = C.c_double
c_vp = C.c_void_p
P = C.POINTER


class FluidsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfluidsim error {code}: {msg}")
        self.code = code


class StokesOpts(C.Structure):
    _fields_ = [("rtol_visc", c_dbl), ("rtol_pressure", c_dbl), ("maxit", C.c_int),
                ("precond", C.c_int), ("warm_start", C.c_int), ("final_div", C.c_int),
                ("bc_mode", C.c_int), ("omega", c_dbl)]


class StokesStats(C.Structure):
    _fields_ = [("iters_visc", C.c_int), ("iters_p1", C.c_int), ("iters_p2", C.c_int),
                ("relres_visc", c_dbl), ("relres_p1", c_dbl), ("relres_p2", c_dbl),
                ("max_div_ustar", c_dbl), ("max_final_div", c_dbl)]


# name -> (restype, argtypes); every symbol declared in include/fluidsim.h
SIGNATURES = {
    "fs_version": (C.c_int, []),
    "fs_last_error": (C.c_char_p, []),
    "fs_device_count": (C.c_int, [P(C.c_int)]),
    "fs_set_device": (C.c_int, [C.c_int]),
    "fs_set_stream": (C.c_int, [c_vp]),
    "fs_sync": (C.c_int, []),
    "fs_stream_wait": (C.c_int, [c_vp]),
    "fs_launch_count": (c_i64, []),
    "fs_timer_start": (C.c_int, []),
    "fs_timer_stop": (C.c_int, [P(C.c_float)]),
    "fs_profile": (C.c_int, [C.c_int]),
    "fs_profile_read": (C.c_int, [c_vp, P(c_i64), P(c_i64)]),
    "fs_profile_read_top": (C.c_int, [P(C.c_double), P(c_i64), P(C.c_double)]),
    "fs_node_file_count": (C.c_int, [C.c_char_p, P(c_i64)]),
    "fs_read_node": (C.c_int, [C.c_char_p, c_vp, c_vp, c_i64]),
    "fs_ele_file_count": (C.c_int, [C.c_char_p, P(c_i64), P(c_i32)]),
    "fs_read_ele": (C.c_int, [C.c_char_p, c_vp, c_i64]),
    "fs_mesh_create": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, P(c_vp)]),
    "fs_mesh_destroy": (C.c_int, [c_vp]),
    "fs_mesh_sizes": (C.c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64)]),
    "fs_csr_pattern": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_scatter_map": (C.c_int, [c_vp, c_vp]),
    "fs_assemble_stiffness": (C.c_int, [c_vp, c_vp]),
    "fs_lumped_mass": (C.c_int, [c_vp, c_vp]),
    "fs_assemble_mass_convection": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "fs_make_rot_bcu": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl, c_dbl]),
    "fs_dye_diffuse": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl]),
    "fs_centroids": (C.c_int, [c_vp, C.c_int, c_vp, c_vp]),
    "fs_assemble_fem": (C.c_int, [c_vp, C.c_int, c_vp, c_dbl, c_vp, c_vp]),
    "fs_divergence": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_gradient": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "fs_bc_set": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64]),
    "fs_make_per_bcu": (C.c_int, [c_vp, c_vp]),
    "fs_make_dir_bcu": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl]),
    "fs_reapply_scalar_bc": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl]),
    "fs_csr_create": (C.c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, P(c_vp)]),
    "fs_csr_from_mesh": (C.c_int, [c_vp, c_vp, P(c_vp)]),
    "fs_csr_destroy": (C.c_int, [c_vp]),
    "fs_csr_sizes": (C.c_int, [c_vp, P(c_i64), P(c_i64)]),
    "fs_csr_get": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "fs_spmv": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_precond_apply": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_precond_bytes": (C.c_int, [c_vp, P(c_dbl)]),
    "fs_cg": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_dbl, C.c_int, C.c_int, C.c_int, P(C.c_int), P(c_dbl)]),
    "fs_bicgstab": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, C.c_int, C.c_int, P(C.c_int), P(c_dbl)]),
    "fs_stokes_default_opts": (C.c_int, [P(StokesOpts)]),
    "fs_stokes_create": (C.c_int, [c_vp, c_dbl, c_dbl, P(c_vp)]),
    "fs_stokes_destroy": (C.c_int, [c_vp]),
    "fs_stokes_step": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl, P(StokesOpts), P(StokesStats)]),
    "fs_stokes_pressure": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_stokes_set_pressure": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_stokes_matrices": (C.c_int, [c_vp, P(c_vp), P(c_vp), c_vp]),
    "fs_stokes_warm_state": (C.c_int, [c_vp, c_vp, C.c_int]),
    "fs_stokes_recycle_state": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, P(c_i64)]),
    "fs_dist_create": (C.c_int, [C.c_int, C.c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, P(c_vp)]),
    "fs_dist_destroy": (C.c_int, [c_vp]),
    "fs_dist_ipc_handle": (C.c_int, [c_vp, c_vp]),
    "fs_dist_connect": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64]),
    "fs_dist_cg_begin": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "fs_dist_cg_run": (C.c_int, [c_vp, c_dbl, c_dbl, c_vp, c_dbl, C.c_int, C.c_int, P(C.c_int), P(c_dbl), c_vp]),
    "fs_stokes_batch_create": (C.c_int, [c_vp, c_i32, P(c_vp)]),
    "fs_stokes_batch_destroy": (C.c_int, [c_vp]),
    "fs_stokes_step_batch": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "fs_tracer_step_batch": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp]),
    "fs_pstokes_create": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, P(c_vp)]),
    "fs_pstokes_destroy": (C.c_int, [c_vp]),
    "fs_pstokes_sizes": (C.c_int, [c_vp, P(c_i64), P(c_i64), P(c_i64), P(c_i64), P(c_i32)]),
    "fs_pstokes_ipc_handle": (C.c_int, [c_vp, c_vp]),
    "fs_pstokes_connect": (C.c_int, [c_vp, c_vp]),
    "fs_pstokes_step": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl, c_vp, c_vp]),
    "fs_pstokes_pressure": (C.c_int, [c_vp, c_vp, c_vp]),
    "fs_pstokes_state": (C.c_int, [c_vp, c_vp, C.c_int]),
    "fs_pstokes_recycle_state": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, P(c_i64)]),
    "fs_pstokes_profile_pcg": (C.c_int, [c_vp, C.c_int, P(c_dbl)]),
    "fs_pstokes_trace": (C.c_int, [c_vp, c_vp, c_i64, P(c_i64)]),
    "fs_locate": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "fs_advect_dye": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, c_vp]),
    "fs_mixing_index": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "fs_locate_exact": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "fs_tracer_step": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, P(c_i64)]),
    "fs_raster_field": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_vp]),
    "fs_raster_colormap": (C.c_int, [c_vp, c_i32, c_i32, c_dbl, c_dbl, c_vp, c_vp, c_vp]),
    "fs_raster_quiver": (C.c_int, [c_vp, c_i32, c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_vp]),
    "fs_raster_points": (C.c_int, [c_vp, c_i32, c_i32, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_i64, c_vp, c_i32, c_dbl]),
}

_NO_CHECK = {"fs_version", "fs_last_error", "fs_launch_count"}


def _load():
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise FluidsimError(rc, lib.fs_last_error().decode(errors="replace"))
    return rc


def call(name, *args):
    fn = getattr(lib, name)
    rc = fn(*args)
    if name not in _NO_CHECK:
        check(rc)
    return rc


# ---- buffers: numpy (host) or torch (device) -----------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Ptr(C.c_void_p):
    """c_void_p that keeps the array it points into alive for as long as the argument object lives
    (i.e. until the foreign call that receives it has returned)."""
    _keep = None


def _ptr_of(addr, owner):
    p = _Ptr(addr)
    p._keep = owner
    return p


def ptr(x, dtype=None, shape=None, name="array"):
    """Pointer argument for a C-contiguous numpy array or torch tensor (None -> NULL).  The returned
    object holds a reference to ``x``, so ``ptr(as_f64(v))`` is safe even when as_f64 had to copy.
    For a torch CUDA tensor the library stream is first ordered after torch's current stream
    (fs_stream_wait): the kernels that read the tensor cannot overtake the ones that produced it."""
    if x is None:
        return None
    if _is_torch(x):
        import torch
        if not x.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        if dtype is not None:
            want = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
            if x.dtype != want:
                raise TypeError(f"{name}: expected dtype {want}, got {x.dtype}")
        if shape is not None and tuple(x.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(x.shape)}")
        if x.is_cuda:
            call("fs_stream_wait", C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        return _ptr_of(x.data_ptr(), x)
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{name}: expected numpy.ndarray or torch.Tensor, got {type(x)}")
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name}: array must be C-contiguous")
    if dtype is not None and x.dtype != np.dtype(dtype):
        raise TypeError(f"{name}: expected dtype {np.dtype(dtype)}, got {x.dtype}")
    if shape is not None and tuple(x.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(x.shape)}")
    return _ptr_of(x.ctypes.data, x)


def as_f64(x, name="array"):
    """Host arrays are converted to contiguous float64 (copy only if needed)."""
    if _is_torch(x):
        return x
    return np.ascontiguousarray(x, dtype=np.float64)


def as_i32(x, name="array"):
    if _is_torch(x):
        return x
    return np.ascontiguousarray(x, dtype=np.int32)


def device_count() -> int:
    n = C.c_int(0)
    call("fs_device_count", C.byref(n))
    return n.value
