"""ctypes binding of the C++ multi-rank host plan (include/feddb200_halo.h, feddlib_b200/csrc/halo.cpp).

`NativeHaloPlan` has the attributes of `dist.HaloPlan` (the numpy implementation the CPU tests keep as the executable
specification) but is built by the C++ code a FEDDLib host links: ownership, Tpetra column map, ghost-row exchange plan,
host-side globalAssemble (`export_add`) and the unique -> repeated vector import.  Communication and the node-pattern
builder are Python callables wrapped as C callbacks, so the same plan runs over torch.distributed (gloo / nccl), over the
in-process thread communicator of the tests, or -- from C++ -- over MPI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, check

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_A2A = C.CFUNCTYPE(C.c_int, C.c_void_p, _i64p, _i64p, C.POINTER(_i64p), _i64p)
_PAT = C.CFUNCTYPE(C.c_int, C.c_void_p, _i32p, C.c_int64, C.c_int64, _i32p, C.c_int64, _i32p, _i32p, C.c_int64,
                   C.POINTER(_i64p), C.POINTER(_i32p))


class _CComm(C.Structure):
    _fields_ = [("user", C.c_void_p), ("rank", C.c_int), ("size", C.c_int), ("alltoallv64", _A2A)]


HALO_SIGNATURES = {
    "feddb200_halo_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(_CComm), C.c_int64, C.c_void_p, C.c_void_p, _PAT, C.c_void_p]),
    "feddb200_halo_free": (None, [C.c_void_p]),
    "feddb200_halo_sizes": (C.c_int, [C.c_void_p] + [_i64p] * 9),
    "feddb200_halo_array": (C.c_void_p, [C.c_void_p, C.c_int, _i64p]),
    "feddb200_halo_recv_slots": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "feddb200_halo_split_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "feddb200_halo_export_add": (C.c_int, [C.c_void_p, C.POINTER(_CComm), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "feddb200_halo_import_vector": (C.c_int, [C.c_void_p, C.POINTER(_CComm), C.c_int, C.c_void_p, C.c_void_p]),
}
_ARRAYS = {"row_lid": (0, np.int32), "col_lid": (1, np.int32), "extra_row": (2, np.int32), "extra_col": (3, np.int32),
           "colmap_gids": (4, np.int64), "unique_gids": (5, np.int64), "ghost_row_gids": (6, np.int64), "ghost_row_owner": (7, np.int64),
           "rowptr": (8, np.int64), "colind": (9, np.int32), "send_counts_nodes": (10, np.int64), "recv_counts_nodes": (11, np.int64),
           "recv_row": (12, np.int64), "recv_pos": (13, np.int64), "recv_len_sender": (14, np.int64), "recv_q": (15, np.int64),
           "import_send_rows": (16, np.int64), "import_send_counts": (17, np.int64), "rep_of_row": (18, np.int64)}


def _load():
    L = _lib.load()
    if not getattr(L, "_halo_typed", False):
        for name, (res, args) in HALO_SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L._halo_typed = True
    return L


class NativeHaloPlan:
    """comm: object with .rank, .size and .alltoallv(list of int64 arrays) -> list of int64 arrays (dist.Comm or a test
    communicator); pattern_fn as in dist.HaloPlan."""

    def __init__(self, comm, gid_rep, owner, pattern_fn):
        self._L = _load()
        self.comm = comm
        self._keep = {}
        gid_rep = np.ascontiguousarray(gid_rep, dtype=np.int64)
        owner32 = np.ascontiguousarray(owner, dtype=np.int32)
        self.gid_rep, self.owner = gid_rep, owner32.astype(np.int64)
        size = comm.size

        def a2a(_user, send, scounts, recv_pp, rcounts):
            try:
                sc = np.ctypeslib.as_array(scounts, shape=(size,)).copy()
                total = int(sc.sum())
                buf = np.ctypeslib.as_array(send, shape=(max(total, 1),))[:total].copy() if total else np.zeros(0, dtype=np.int64)
                off = np.concatenate([[0], np.cumsum(sc)])
                got = comm.alltoallv([buf[off[d]:off[d + 1]] for d in range(size)])
                flat = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.int64) for g in got]) if got else np.zeros(0, dtype=np.int64))
                if flat.size == 0:
                    flat = np.zeros(1, dtype=np.int64)
                self._keep["recv"] = flat
                recv_pp[0] = flat.ctypes.data_as(_i64p)
                for s in range(size):
                    rcounts[s] = len(got[s])
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        def pat(_user, row_lid, n_rows, n_owned, col_lid, n_cols, er, ec, n_extra, rp_pp, ci_pp):
            try:
                nn = gid_rep.size
                rl = np.ctypeslib.as_array(row_lid, shape=(max(nn, 1),))[:nn].copy()
                cl = np.ctypeslib.as_array(col_lid, shape=(max(nn, 1),))[:nn].copy() if col_lid else None
                xr = np.ctypeslib.as_array(er, shape=(n_extra,)).copy() if n_extra else None
                xc = np.ctypeslib.as_array(ec, shape=(n_extra,)).copy() if n_extra else None
                rp, ci = pattern_fn(rl, int(n_rows), int(n_owned), cl, int(n_cols), xr, xc)
                rp = np.ascontiguousarray(rp, dtype=np.int64)
                ci = np.ascontiguousarray(ci, dtype=np.int32) if len(ci) else np.zeros(1, dtype=np.int32)
                self._keep["rp"], self._keep["ci"] = rp, ci
                rp_pp[0] = rp.ctypes.data_as(_i64p)
                ci_pp[0] = ci.ctypes.data_as(_i32p)
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return -2

        self._a2a, self._pat = _A2A(a2a), _PAT(pat)
        self._ccomm = _CComm(None, comm.rank, comm.size, self._a2a)
        h = C.c_void_p()
        check(self._L.feddb200_halo_create(C.byref(h), C.byref(self._ccomm), gid_rep.size, gid_rep.ctypes.data, owner32.ctypes.data,
                                          self._pat, None))
        self._h = h
        v = [C.c_int64() for _ in range(9)]
        check(self._L.feddb200_halo_sizes(self._h, *[C.byref(x) for x in v]))
        (self.n_owned, self.n_ghost, self.n_rows, self.n_colmap, self.n_cols, self.n_extra, self.nnz_owned_nodes, self.nnz_nodes,
         self.n_recv) = [int(x.value) for x in v]
        # the final node pattern stays what the callback returned (the plan references those arrays)
        self.rowptr, self.colind = self._keep["rp"], self._keep["ci"][: self.nnz_nodes]
        self._slots = {}

    def __getattr__(self, name):
        # plan arrays are fetched on first use (several are as large as the mesh)
        if name in _ARRAYS and name not in ("rowptr", "colind") and self.__dict__.get("_h"):
            which, dt = _ARRAYS[name]
            n = C.c_int64()
            ptr = self._L.feddb200_halo_array(self._h, which, C.byref(n))
            arr = (np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n.value,)).copy()
                   if n.value else np.zeros(0, dtype=dt))
            self.__dict__[name] = arr
            return arr
        raise AttributeError(name)

    def close(self):
        if getattr(self, "_h", None):
            self._L.feddb200_halo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @staticmethod
    def _factor(rd, cd, mode):
        return 1 if mode == BLOCK_SCALAR else (rd if mode == BLOCK_DIAG else rd * cd)

    def split_sizes(self, rd, cd, mode):
        s, r = np.zeros(self.comm.size, dtype=np.int64), np.zeros(self.comm.size, dtype=np.int64)
        check(self._L.feddb200_halo_split_sizes(self._h, rd, cd, mode, s.ctypes.data, r.ctypes.data))
        return s.tolist(), r.tolist()

    def recv_slots(self, rd, cd, mode):
        key = (rd, cd, mode)
        if key not in self._slots:
            out = np.zeros(max(1, self._factor(rd, cd, mode) * self.n_recv), dtype=np.int64)
            check(self._L.feddb200_halo_recv_slots(self._h, rd, cd, mode, out.ctypes.data))
            self._slots[key] = out[: self._factor(rd, cd, mode) * self.n_recv]
        return self._slots[key]

    def vec_split_sizes(self, dofs):
        send = np.array([(self.ghost_row_owner == d).sum() for d in range(self.comm.size)], dtype=np.int64) * dofs
        first = self.recv_q == 0
        off = np.concatenate([[0], np.cumsum(self.recv_counts_nodes)])
        recv = np.array([first[off[s]:off[s + 1]].sum() for s in range(self.comm.size)], dtype=np.int64) * dofs
        return send.tolist(), recv.tolist()

    def vec_recv_slots(self, dofs):
        I = self.recv_row[self.recv_q == 0]
        return (dofs * I[:, None] + np.arange(dofs)[None, :]).ravel().astype(np.int64)

    def export_add(self, values, rd, cd, mode, nnz_owned_values):
        """Host-side globalAssemble (Matrix::fillComplete of the reference): ghost part of `values` to the owners, added."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        check(self._L.feddb200_halo_export_add(self._h, C.byref(self._ccomm), rd, cd, mode, int(nnz_owned_values), values.ctypes.data))
        return values

    def import_vector(self, u_unique, dofs):
        """unique -> repeated import of a node-wise interleaved vector (MultiVector::importFromVector)."""
        u_unique = np.ascontiguousarray(u_unique, dtype=np.float64)
        out = np.zeros(dofs * self.gid_rep.size, dtype=np.float64)
        check(self._L.feddb200_halo_import_vector(self._h, C.byref(self._ccomm), dofs, u_unique.ctypes.data, out.ctypes.data))
        return out
