"""ctypes access to oracle/_ref/libfedd_ref_bm.so: the reference's own BlockMatrix::merge / BlockMap::merge
(core/LinearAlgebra/BlockMatrix_def.hpp:119-270, BlockMap_def.hpp:41-92) compiled where they lie against mock containers
(oracle/ref_shim/bm_driver.cpp).  Test infrastructure: pins oracle/csrops.py block_merge."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libfedd_ref_bm.so")
_L = None


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _L
    if _L is None:
        _L = C.CDLL(_SO)
        _L.ref_bm_last_error.restype = C.c_char_p
        _L.ref_bm_merge.restype = C.c_void_p
        _L.ref_bm_rows.restype = C.c_int64
        _L.ref_bm_rows.argtypes = [C.c_void_p]
        _L.ref_bm_nnz.restype = C.c_int64
        _L.ref_bm_nnz.argtypes = [C.c_void_p]
        _L.ref_bm_get.argtypes = [C.c_void_p] * 5
        _L.ref_bm_free.argtypes = [C.c_void_p]
    return _L


def merge(row_gids, blocks):
    """row_gids[i]: global ids of block row i (local row order); blocks[(i, j)] = (rowptr, colind_local, values, col_gids).
    Returns (rowptr, row gids of the merged map, column gids, values) of the reference's merged matrix."""
    nb = len(row_gids)
    keep = []
    n = nb * nb
    rg = (C.c_void_p * nb)()
    nr = np.array([len(g) for g in row_gids], dtype=np.int64)
    for i, g in enumerate(row_gids):
        a = np.ascontiguousarray(g, dtype=np.int64); keep.append(a); rg[i] = a.ctypes.data
    rp, ci, va, cg = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
    nc = np.zeros(n, dtype=np.int64)
    for (i, j), (rowptr, colind, values, col_gid) in blocks.items():
        k = i * nb + j
        a = [np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(colind, dtype=np.int32),
             np.ascontiguousarray(values, dtype=np.float64), np.ascontiguousarray(col_gid, dtype=np.int64)]
        keep.append(a)
        rp[k], ci[k], va[k], cg[k] = a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data
        nc[k] = a[3].size
    L = lib()
    h = L.ref_bm_merge(C.c_int(nb), C.c_void_p(nr.ctypes.data), rg, rp, ci, va, C.c_void_p(nc.ctypes.data), cg)
    if not h:
        raise RuntimeError(L.ref_bm_last_error().decode())
    try:
        rows, nnz = L.ref_bm_rows(h), L.ref_bm_nnz(h)
        o = [np.zeros(rows + 1, dtype=np.int64), np.zeros(rows, dtype=np.int64), np.zeros(max(nnz, 1), dtype=np.int64), np.zeros(max(nnz, 1))]
        L.ref_bm_get(h, *[C.c_void_p(x.ctypes.data) for x in o])
        return o[0], o[1], o[2][:nnz], o[3][:nnz]
    finally:
        L.ref_bm_free(h)
