"""ctypes front-end of oracle/_ref/libfedd_ref.so: the REFERENCE's own hot-path routines (FEDDLib
feddlib/core/FE/FE_def.hpp, sliced at build time by oracle/ref_shim/extract.py and compiled against mock
Trilinos containers).  Test infrastructure; used to pin the restatement in oracle/fedd_oracle.c and as the
`cpu_baseline.kind = "reference"` timing leg.  The library is built only where /root/reference exists
(`make -C oracle ref`); the built .so travels with the repository snapshot."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libfedd_ref.so")
_lib = None
OPS = {"laplace": 0, "laplace_vec": 1, "linelas": 2, "advection": 3, "advection_in_u": 4, "div": 5, "div_fast": 6, "mass": 7, "mass_vec": 8, "bdstab": 9}


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        vp, i32p, i64p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        L.ref_assemble.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int64, vp, C.c_int, vp, C.c_int64, vp,
                                   vp, C.c_int, C.c_int64, vp, vp, C.c_double, C.c_double, vp, vp]
        L.ref_last_error.restype = C.c_char_p
        L.fo_matrix_new.restype = vp
        L.fo_matrix_new.argtypes = [C.c_int64, C.c_int32]
        L.fo_matrix_free.argtypes = [vp]
        L.fo_matrix_nnz.restype = C.c_int64
        L.fo_matrix_nnz.argtypes = [vp]
        L.fo_fill_complete.argtypes = [vp]
        L.fo_get_csr.argtypes = [vp, vp, vp, vp]
        L.ref_get_dphi.argtypes = [C.c_int, C.c_char_p, C.c_int, vp, vp]
        L.ref_get_phi.argtypes = [C.c_int, C.c_char_p, C.c_int, vp, vp]
        L.ref_determine_degree.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _csr(L, h, nrows):
    nnz = L.fo_matrix_nnz(h)
    rp = np.empty(nrows + 1, dtype=np.int64); ci = np.empty(nnz, dtype=np.int64); v = np.empty(nnz)
    L.fo_get_csr(h, _p(rp), _p(ci), _p(v))
    return rp, ci, v


def assemble_stress(dim, fe, conn, coords, func, gid=None, nrows_global=None):
    """FE::assemblyStress of the reference with the Python callback func(xyz) -> float; returns CSR (rowptr, col gid, values)."""
    L = lib()
    conn = np.ascontiguousarray(conn, dtype=np.int32); coords = np.ascontiguousarray(coords, dtype=np.float64)
    nn = coords.shape[0]
    gid = np.arange(nn, dtype=np.int64) if gid is None else np.ascontiguousarray(gid, dtype=np.int64)
    nglob = int(nrows_global if nrows_global is not None else gid.max() + 1)
    cb = O.COEFF_FUNC(lambda x, _u: float(func(np.array([x[d] for d in range(dim)]))))
    L.ref_assemble_stress.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                      O.COEFF_FUNC, C.c_void_p, C.c_void_p]
    h = L.fo_matrix_new(dim * nglob, 64)
    try:
        rc = L.ref_assemble_stress(dim, fe.encode(), conn.shape[0], _p(conn), conn.shape[1], _p(coords), nn, _p(gid), cb, None, h)
        if rc != 0:
            raise RuntimeError(L.ref_last_error().decode())
        L.fo_fill_complete(h)
        return _csr(L, h, dim * nglob)
    finally:
        L.fo_matrix_free(h)


def assemble(op, dim, fe, conn, coords, gid=None, u=None, lam=0.0, mu=0.0, fe2=None, conn2=None, gid2=None,
             nrows_global=None):
    """Run the reference routine; returns CSR (rowptr, col gid, values) -- for div ops a pair (B, BT)."""
    L = lib()
    conn = np.ascontiguousarray(conn, dtype=np.int32); coords = np.ascontiguousarray(coords, dtype=np.float64)
    nn = coords.shape[0]
    gid = np.arange(nn, dtype=np.int64) if gid is None else np.ascontiguousarray(gid, dtype=np.int64)
    nglob = int(nrows_global if nrows_global is not None else gid.max() + 1)
    dofs = 1 if op in ("laplace", "mass", "bdstab") else dim
    is_div = op in ("div", "div_fast")
    if is_div:
        conn2 = np.ascontiguousarray(conn2, dtype=np.int32)
        nn2 = int(conn2.max()) + 1
        gid2 = np.arange(nn2, dtype=np.int64) if gid2 is None else np.ascontiguousarray(gid2, dtype=np.int64)
        nrowsA, nrowsB = int(gid2.max()) + 1, dim * nglob
    else:
        nn2, nrowsA, nrowsB = 0, dofs * nglob, 1
    hA, hB = L.fo_matrix_new(nrowsA, 64), L.fo_matrix_new(nrowsB, 64)
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    rc = L.ref_assemble(OPS[op], dim, fe.encode(), (fe2 or fe).encode(), conn.shape[0], _p(conn), conn.shape[1],
                        _p(coords), nn, _p(gid), _p(conn2) if is_div else None, conn2.shape[1] if is_div else 0, nn2,
                        _p(gid2) if is_div else None, _p(uu), float(lam), float(mu), hA, hB)
    try:
        if rc != 0:
            raise ValueError("reference: " + L.ref_last_error().decode())
        out = _csr(L, hA, nrowsA)
        return (out, _csr(L, hB, nrowsB)) if is_div else out
    finally:
        L.fo_matrix_free(hA); L.fo_matrix_free(hB)


def assemble_rhs(dim, fe, conn, coords, value_func, deg_func=0, vec_field=False):
    """Reference FE::assemblyRHS (constant source) -> repeated vector."""
    L = lib()
    L.ref_assemble_rhs.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]
    conn = np.ascontiguousarray(conn, dtype=np.int32); coords = np.ascontiguousarray(coords, dtype=np.float64)
    f = np.ascontiguousarray(value_func, dtype=np.float64)
    out = np.zeros(coords.shape[0] * (dim if vec_field else 1))
    rc = L.ref_assemble_rhs(dim, fe.encode(), conn.shape[0], _p(conn), conn.shape[1], _p(coords), coords.shape[0],
                            int(bool(vec_field)), int(deg_func), _p(f), _p(out))
    if rc != 0:
        raise ValueError("reference: " + L.ref_last_error().decode())
    return out


def get_dphi(dim, fe, deg):
    n = O.lib().fo_nloc(dim, fe.encode())
    d = np.zeros(30 * 10 * 3); w = np.zeros(30)
    nq = lib().ref_get_dphi(dim, fe.encode(), deg, _p(d), _p(w))
    return d[: nq * n * dim].reshape(nq, n, dim).copy(), w[:nq].copy()


def get_phi(dim, fe, deg):
    n = O.lib().fo_nloc(dim, fe.encode())
    d = np.zeros(30 * 10); w = np.zeros(30)
    nq = lib().ref_get_phi(dim, fe.encode(), deg, _p(d), _p(w))
    return d[: nq * n].reshape(nq, n).copy(), w[:nq].copy()


def determine_degree(dim, fe1, fe2, t1, t2, extra=0):
    return lib().ref_determine_degree(dim, fe1.encode(), fe2.encode(), t1, t2, extra)
