"""ctypes access to oracle/_ref/libfedd_ref_bc.so: the reference's own BCBuilder members (setSystem -> setDirichletBC ->
setLocalRowOne / setLocalRowZero, setRHS; core/General/BCBuilder_def.hpp:93-200, 589-709) compiled where they lie against CSR
mocks (oracle/ref_shim/bc_driver.cpp).  Test infrastructure: pins oracle.set_dirichlet_rows / set_dirichlet_rhs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libfedd_ref_bc.so")
BC_FUNC = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double))


def available() -> bool:
    return os.path.exists(_SO)


_L = None


def lib():
    global _L
    if _L is None:
        _L = C.CDLL(_SO)
        _L.ref_bc_last_error.restype = C.c_char_p
    return _L


def _strs(types):
    arr = (C.c_char_p * len(types))(*[t.encode() for t in types])
    return arr


def set_system(dim, node_flags, node_gid, bcs, dofs_of_block, blocks):
    """bcs: list of (flag, block, type, dofs); blocks[(i, j)] = (rowptr int64, colind int32, values f64, col_gid int64) of a
    CSR on the node set (rows dofs_of_block[i] * node + d).  Returns {(i, j): new values}."""
    nb = len(dofs_of_block)
    node_flags = np.ascontiguousarray(node_flags, dtype=np.int32)
    node_gid = np.ascontiguousarray(node_gid, dtype=np.int64)
    n = nb * nb
    rp, ci, va, cg = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
    nc = np.zeros(n, dtype=np.int64)
    keep, out = [], {}
    for (i, j), (rowptr, colind, values, col_gid) in blocks.items():
        k = i * nb + j
        a = [np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(colind, dtype=np.int32),
             np.array(values, dtype=np.float64, copy=True), np.ascontiguousarray(col_gid, dtype=np.int64)]
        keep.append(a)
        rp[k], ci[k], va[k], cg[k] = a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data
        nc[k] = a[3].size
        out[(i, j)] = a[2]
    flag = np.array([b[0] for b in bcs], dtype=np.int32)
    block = np.array([b[1] for b in bcs], dtype=np.int32)
    dofs = np.array([b[3] for b in bcs], dtype=np.int32)
    dob = np.asarray(dofs_of_block, dtype=np.int32)
    rc = lib().ref_bc_set_system(C.c_int(dim), C.c_int64(node_flags.size), C.c_void_p(node_flags.ctypes.data), C.c_void_p(node_gid.ctypes.data),
                                 C.c_int(len(bcs)), C.c_void_p(flag.ctypes.data), C.c_void_p(block.ctypes.data), _strs([b[2] for b in bcs]),
                                 C.c_void_p(dofs.ctypes.data), C.c_int(nb), C.c_void_p(dob.ctypes.data), rp, ci, va, cg, C.c_void_p(nc.ctypes.data))
    if rc != 0:
        raise RuntimeError(lib().ref_bc_last_error().decode())
    return out


def set_rhs(dim, node_flags, points, node_gid, bcs, func, params, dofs_of_block, rhs_blocks, t=0.0):
    """BCBuilder::setRHS: rhs_blocks[i] (dofs_of_block[i] * n_nodes) -> new arrays; func(x, res, t, params) fills res."""
    nb = len(dofs_of_block)
    node_flags = np.ascontiguousarray(node_flags, dtype=np.int32)
    points = np.ascontiguousarray(points, dtype=np.float64)
    node_gid = np.ascontiguousarray(node_gid, dtype=np.int64)
    out = [np.array(r, dtype=np.float64, copy=True) for r in rhs_blocks]
    ptrs = (C.c_void_p * nb)(*[o.ctypes.data for o in out])
    flag = np.array([b[0] for b in bcs], dtype=np.int32)
    block = np.array([b[1] for b in bcs], dtype=np.int32)
    dofs = np.array([b[3] for b in bcs], dtype=np.int32)
    dob = np.asarray(dofs_of_block, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)

    def cb(x, res, tt, par):
        xv = np.array([x[d] for d in range(dim)])
        pv = np.array([par[k] for k in range(params.size)])
        r = func(xv, tt, pv)
        for d, v in enumerate(r):
            res[d] = v

    cfun = BC_FUNC(cb)
    rc = lib().ref_bc_set_rhs(C.c_int(dim), C.c_int64(node_flags.size), C.c_void_p(node_flags.ctypes.data), C.c_void_p(points.ctypes.data),
                              C.c_void_p(node_gid.ctypes.data), C.c_int(len(bcs)), C.c_void_p(flag.ctypes.data), C.c_void_p(block.ctypes.data),
                              _strs([b[2] for b in bcs]), C.c_void_p(dofs.ctypes.data), cfun, C.c_void_p(params.ctypes.data), C.c_int(params.size),
                              C.c_int(nb), C.c_void_p(dob.ctypes.data), ptrs, C.c_double(t))
    if rc != 0:
        raise RuntimeError(lib().ref_bc_last_error().decode())
    return out
