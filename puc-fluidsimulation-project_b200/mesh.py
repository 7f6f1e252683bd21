"""Mesh ingest and host-side setup logic (Triangle .node/.ele/.poly files, periodic
pairs, index sets) plus the synthetic square-with-hole generator used by the
benchmarks.  Mirrors the reference's loader functions (same names, same returns).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import call


# --------------------------------------------------------------------------- readers
def readNode(filepath, dtype=np.float64):
    """readNode(filepath) -> (nodes_coords (N,2), nodes_boundary_markers (N,) i32).
    code/StokesColor.py:54-78; ``dtype=np.float32`` gives the code/poisson.py:27-56
    variant (same parse, coordinates stored as float32)."""
    path = str(filepath).encode()
    n = C.c_int64(0)
    call("fs_node_file_count", path, C.byref(n))
    coords = np.zeros((n.value, 2), dtype=np.float64)
    markers = np.zeros(n.value, dtype=np.int32)
    call("fs_read_node", path, _lib.ptr(coords), _lib.ptr(markers), n.value)
    if np.dtype(dtype) != np.float64:
        coords = coords.astype(dtype)
    return coords, markers


def readEle(filepath):
    """readEle(filepath) -> (T,3) int32, 0-based.  code/StokesColor.py:82-95.
    Only the three corner ids are read (6-node .ele files keep their corners)."""
    path = str(filepath).encode()
    t = C.c_int64(0)
    k = C.c_int32(0)
    call("fs_ele_file_count", path, C.byref(t), C.byref(k))
    tris = np.zeros((t.value, 3), dtype=np.int32)
    call("fs_read_ele", path, _lib.ptr(tris), t.value)
    return tris


def readPoly(path):
    """readPoly(path) -> (segments (S,2) int, boundaryMarkers (S,) int).
    Same return as code/poisson.py:76-97 (never called by the reference's main bodies; host only).
    Parsed by the full Triangle .poly reader of meshgen (vertex and hole sections are skipped)."""
    from .meshgen import read_poly_full
    _, _, segs, smark, _ = read_poly_full(path, need_vertices=False)
    return np.asarray(segs, dtype=int), np.asarray(smark, dtype=int)


# numpy-only host helpers (boundary sets, synthetic / refined meshes, writers) live in hostmesh.py, a
# module without any dependency on the shared library, so that CPU-only tools can load it by path
from .hostmesh import (write_node, write_ele, find_boundary_pairs, filter_wall_pairs, index_sets,  # noqa: E402,F401
                       square_with_hole, refine_mesh)
