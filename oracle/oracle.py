"""ctypes front-end of the CPU ORACLE (oracle/fedd_oracle.c).

Test infrastructure only: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The product package (feddlib_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfedd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "fedd_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_I32P = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_I64P = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_F64P = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
COEFF_FUNC = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.fo_matrix_new.restype = C.c_void_p
        L.fo_matrix_new.argtypes = [C.c_int64, C.c_int32]
        L.fo_matrix_free.argtypes = [C.c_void_p]
        L.fo_fill_complete.restype = C.c_int64
        L.fo_fill_complete.argtypes = [C.c_void_p]
        L.fo_matrix_nnz.restype = C.c_int64
        L.fo_matrix_nnz.argtypes = [C.c_void_p]
        L.fo_get_csr.argtypes = [C.c_void_p, _I64P, _I64P, _F64P]
        L.fo_matrix_scale.argtypes = [C.c_void_p, C.c_double]
        L.fo_insert.argtypes = [C.c_void_p, C.c_int64, C.c_int32, _I64P, _F64P]
        common = [C.c_int, C.c_char_p, C.c_int64, _I32P, _F64P, _I64P]
        L.fo_assembly_laplace.argtypes = common + [C.c_void_p]
        L.fo_assembly_laplace_vecfield.argtypes = common + [C.c_void_p]
        L.fo_assembly_mass.argtypes = common + [C.c_int, C.c_void_p]
        L.fo_assembly_bdstab.argtypes = common + [C.c_void_p]
        L.fo_assembly_stress.argtypes = common + [COEFF_FUNC, C.c_void_p, C.c_void_p]
        L.fo_assembly_rhs.argtypes = [C.c_int, C.c_char_p, C.c_int64, _I32P, _F64P, C.c_int, C.c_int, _F64P, _F64P]
        L.fo_assembly_linelas.argtypes = common + [C.c_double, C.c_double, C.c_void_p]
        L.fo_assembly_advection.argtypes = common + [_F64P, C.c_void_p]
        L.fo_assembly_advection_in_u.argtypes = common + [_F64P, C.c_void_p]
        L.fo_assembly_div_divT.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int64, _I32P, _F64P, _I64P,
                                           _I32P, _I64P, C.c_void_p, C.c_void_p, C.c_int]
        L.fo_quadrature.argtypes = [C.c_int, C.c_int, _F64P, _F64P]
        L.fo_get_phi.argtypes = [C.c_int, C.c_char_p, C.c_int, _F64P, _F64P]
        L.fo_get_dphi.argtypes = [C.c_int, C.c_char_p, C.c_int, _F64P, _F64P]
        L.fo_determine_degree2.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.fo_determine_degree1.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.fo_nloc.argtypes = [C.c_int, C.c_char_p]
        _lib = L
    return _lib


STD, GRAD = 0, 1


class Matrix:
    """Emulation of FEDD::Matrix(map, numEntries) on a dense global row-id range [0, nrows)."""

    def __init__(self, nrows: int, cap_hint: int = 32):
        self.nrows = int(nrows)
        self._h = lib().fo_matrix_new(self.nrows, int(cap_hint))
        self.filled = False

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fo_matrix_free(self._h)
            self._h = None

    def insertGlobalValues(self, row, cols, vals):
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        if lib().fo_insert(self._h, int(row), cols.size, cols, vals) != 0:
            raise IndexError("row out of range")

    def fillComplete(self):
        lib().fo_fill_complete(self._h)
        self.filled = True

    def scale(self, s: float):
        lib().fo_matrix_scale(self._h, float(s))

    def csr(self):
        """(rowptr int64[nrows+1], col_gid int64[nnz], values float64[nnz])"""
        if not self.filled:
            self.fillComplete()
        nnz = lib().fo_matrix_nnz(self._h)
        rowptr = np.empty(self.nrows + 1, dtype=np.int64)
        col = np.empty(nnz, dtype=np.int64)
        val = np.empty(nnz, dtype=np.float64)
        lib().fo_get_csr(self._h, rowptr, col, val)
        return rowptr, col, val

    def scipy(self, ncols=None):
        import scipy.sparse as sp
        rowptr, col, val = self.csr()
        return sp.csr_matrix((val, col, rowptr), shape=(self.nrows, ncols or self.nrows))


def _chk(rc, what):
    if rc != 0:
        raise ValueError(f"oracle: {what} failed (unsupported dim / FE type?)")


def _prep(conn, coords, gid):
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    gid = np.ascontiguousarray(gid, dtype=np.int64)
    return conn, coords, gid


def assembly_laplace(dim, fe, conn, coords, gid, A: Matrix):
    conn, coords, gid = _prep(conn, coords, gid)
    _chk(lib().fo_assembly_laplace(dim, fe.encode(), conn.shape[0], conn, coords, gid, A._h), "assemblyLaplace")


def assembly_rhs(dim, fe, conn, coords, value_func, deg_func=0, vec_field=False, rhs=None):
    """FE::assemblyRHS with a constant source (FE_def.hpp:4694-4766): adds into (and returns) the repeated vector."""
    conn = np.ascontiguousarray(conn, dtype=np.int32); coords = np.ascontiguousarray(coords, dtype=np.float64)
    f = np.ascontiguousarray(value_func, dtype=np.float64)
    n = coords.shape[0] * (dim if vec_field else 1)
    rhs = np.zeros(n) if rhs is None else rhs
    _chk(lib().fo_assembly_rhs(dim, fe.encode(), conn.shape[0], conn, coords, int(bool(vec_field)), int(deg_func), f, rhs), "assemblyRHS")
    return rhs


def assembly_stress(dim, fe, conn, coords, gid, func, A: Matrix):
    """FE::assemblyStress (FE_def.hpp:2407-2735); func(xyz: np.ndarray[dim]) -> float is the CoeffFunc_Type callback."""
    conn, coords, gid = _prep(conn, coords, gid)
    cb = COEFF_FUNC(lambda x, _u: float(func(np.array([x[d] for d in range(dim)]))))
    _chk(lib().fo_assembly_stress(dim, fe.encode(), conn.shape[0], conn, coords, gid, cb, None, A._h), "assemblyStress")


def assembly_bdstab(dim, fe, conn, coords, gid, A: Matrix):
    """FE::assemblyBDStabilization (FE_def.hpp:2151-2220), P1 only."""
    conn, coords, gid = _prep(conn, coords, gid)
    _chk(lib().fo_assembly_bdstab(dim, fe.encode(), conn.shape[0], conn, coords, gid, A._h), "assemblyBDStabilization")


def assembly_mass(dim, fe, conn, coords, gid, A: Matrix, vec_field=False):
    """FE::assemblyMass, fieldType "Scalar" / "Vector" (FE_def.hpp:454-521)."""
    conn, coords, gid = _prep(conn, coords, gid)
    _chk(lib().fo_assembly_mass(dim, fe.encode(), conn.shape[0], conn, coords, gid, int(bool(vec_field)), A._h), "assemblyMass")


def assembly_laplace_vecfield(dim, fe, conn, coords, gid, A: Matrix):
    conn, coords, gid = _prep(conn, coords, gid)
    _chk(lib().fo_assembly_laplace_vecfield(dim, fe.encode(), conn.shape[0], conn, coords, gid, A._h),
         "assemblyLaplaceVecField")


def assembly_linelas(dim, fe, conn, coords, gid, lam, mu, A: Matrix):
    conn, coords, gid = _prep(conn, coords, gid)
    _chk(lib().fo_assembly_linelas(dim, fe.encode(), conn.shape[0], conn, coords, gid, lam, mu, A._h),
         "assemblyLinElasXDim")


def assembly_advection(dim, fe, conn, coords, gid, u, A: Matrix):
    conn, coords, gid = _prep(conn, coords, gid)
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.size == dim * coords.shape[0]
    _chk(lib().fo_assembly_advection(dim, fe.encode(), conn.shape[0], conn, coords, gid, u, A._h),
         "assemblyAdvectionVecField")


def assembly_advection_in_u(dim, fe, conn, coords, gid, u, A: Matrix):
    conn, coords, gid = _prep(conn, coords, gid)
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.size == dim * coords.shape[0]
    _chk(lib().fo_assembly_advection_in_u(dim, fe.encode(), conn.shape[0], conn, coords, gid, u, A._h),
         "assemblyAdvectionInUVecField")


def assembly_div_divT(dim, fe1, fe2, conn1, coords1, gid1, conn2, gid2, B: Matrix, BT: Matrix, fast=False):
    conn1, coords1, gid1 = _prep(conn1, coords1, gid1)
    conn2 = np.ascontiguousarray(conn2, dtype=np.int32)
    gid2 = np.ascontiguousarray(gid2, dtype=np.int64)
    assert conn1.shape[0] == conn2.shape[0]
    _chk(lib().fo_assembly_div_divT(dim, fe1.encode(), fe2.encode(), conn1.shape[0], conn1, coords1, gid1,
                                    conn2, gid2, B._h, BT._h, int(bool(fast))), "assemblyDivAndDivT")


def quadrature(dim, deg):
    pts = np.zeros(16 * 3)
    w = np.zeros(16)
    n = lib().fo_quadrature(dim, deg, pts, w)
    if n < 0:
        raise ValueError("no such rule on this path")
    return pts[: n * dim].reshape(n, dim).copy(), w[:n].copy()


def get_phi(dim, fe, deg):
    n = lib().fo_nloc(dim, fe.encode())
    phi = np.zeros(16 * 10)
    w = np.zeros(16)
    nq = lib().fo_get_phi(dim, fe.encode(), deg, phi, w)
    return phi[: nq * n].reshape(nq, n).copy(), w[:nq].copy()


def get_dphi(dim, fe, deg):
    n = lib().fo_nloc(dim, fe.encode())
    dphi = np.zeros(16 * 10 * 3)
    w = np.zeros(16)
    nq = lib().fo_get_dphi(dim, fe.encode(), deg, dphi, w)
    return dphi[: nq * n * dim].reshape(nq, n, dim).copy(), w[:nq].copy()


def determine_degree(dim, fe1, fe2, t1, t2, extra=0):
    return lib().fo_determine_degree2(dim, fe1.encode(), fe2.encode(), t1, t2, extra)


def determine_degree1(dim, fe, t):
    return lib().fo_determine_degree1(dim, fe.encode(), t)


def set_dirichlet_rows(rowptr, colgid, values, row_gid_of_dof, dof_is_dirichlet, diagonal_block=True):
    """BCBuilder::setLocalRowOne / setLocalRowZero (core/General/BCBuilder_def.hpp:653-709) on a dof-level CSR:
    every Dirichlet dof row is replaced by zeros (:672, :704) and, on a diagonal block, the entry whose column gid
    equals the row gid by one (:674-679).  Pinned: tests/test_bc_vs_ref.py runs the reference's own members (compiled where
    they lie against CSR mocks, oracle/ref_bc.py) on the same systems and compares bitwise; golden vectors in
    tests/golden/bc_vectors.npz.  Returns a new values array."""
    out = np.array(values, dtype=np.float64, copy=True)
    for r in np.nonzero(dof_is_dirichlet)[0]:
        a, b = rowptr[r], rowptr[r + 1]
        out[a:b] = 0.0
        if diagonal_block:
            hit = np.nonzero(colgid[a:b] == row_gid_of_dof[r])[0]
            if hit.size:
                out[a + hit[0]] = 1.0
    return out


DIRICHLET_MASKS = {"Dirichlet": 0b111, "Dirichlet_X": 0b001, "Dirichlet_Y": 0b010, "Dirichlet_Z": 0b100, "Dirichlet_X_Y": 0b011,
                   "Dirichlet_X_Z": 0b101, "Dirichlet_Y_Z": 0b110}


def dirichlet_dof_masks(node_flags, bcs, block, dofs):
    """Per node, the bit mask of the dofs of block `block` that carry a Dirichlet condition: BCBuilder::findFlag picks the FIRST
    entry of the BC table with the node's flag and this block (BCBuilder_def.hpp:561-586); its type selects the dofs
    (:662-669).  bcs: list of (flag, block, type, dofs) in addBC order.  Also returns the index of the matching entry (-1: none)."""
    node_flags = np.asarray(node_flags)
    mask = np.zeros(node_flags.size, dtype=np.uint8)
    which = -np.ones(node_flags.size, dtype=np.int64)
    for i, fl in enumerate(node_flags):
        for k, (flag, blk, typ, _d) in enumerate(bcs):
            if flag == fl and blk == block:
                if typ in DIRICHLET_MASKS:
                    mask[i] = DIRICHLET_MASKS[typ] & ((1 << dofs) - 1)
                    which[i] = k
                break
    return mask, which


def set_dirichlet_rhs(rhs, node_flags, points, bcs, block, dofs, func, params, t=0.0):
    """BCBuilder::setRHS, Dirichlet part (core/General/BCBuilder_def.hpp:93-166): at every node whose flag has a Dirichlet entry
    for this block, result starts as the node's coordinates (:131-133), the boundary function overwrites it (:134) and the dofs
    selected by the type are written into the right-hand side (:138-163).  Returns a new vector."""
    out = np.array(rhs, dtype=np.float64, copy=True)
    mask, which = dirichlet_dof_masks(node_flags, bcs, block, dofs)
    dim = points.shape[1]
    for i in np.nonzero(which >= 0)[0]:
        res = [points[i][d] if d < dim else 0.0 for d in range(dofs)]
        got = func(np.array(points[i]), t, np.asarray(params))
        for d, v in enumerate(got):
            res[d] = v
        for d in range(dofs):
            if (mask[i] >> d) & 1:
                out[dofs * i + d] = res[d]
    return out
