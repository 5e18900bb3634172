"""Thin object layer over the C ABI (include/feddb200.h): Context, Mesh, Pattern.

PyTorch appears only as plumbing (device buffers, streams); every computation is in
libfeddb200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, SCATTER_ATOMIC, SCATTER_COLOURED, SCATTER_GATHER, LogicError, check,
                   ptr)

_MODES = {"atomic": SCATTER_ATOMIC, "coloured": SCATTER_COLOURED, "colored": SCATTER_COLOURED,
          "gather": SCATTER_GATHER}


class Context:
    """One engine context per GPU / rank (feddb200_create)."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.feddb200_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        if use_torch_stream:
            self.bind_torch_stream()

    def bind_torch_stream(self):
        """Run the engine's kernels on torch's current CUDA stream so torch events time them."""
        import torch
        self.torch_stream = torch.cuda.current_stream(self.device)   # the stream the engine launches on, as a torch object
        check(self._L.feddb200_set_stream(self._h, C.c_void_p(self.torch_stream.cuda_stream)))

    def set_stream(self, stream):
        """Launch the following calls on `stream` (a torch.cuda.Stream); bind_torch_stream's stream stays the main one."""
        check(self._L.feddb200_set_stream(self._h, C.c_void_p(stream.cuda_stream)))

    def bind_host_numa(self) -> int:
        """Bind this thread to the CPUs next to the GPU (feddb200_bind_host_numa): page-locked buffers allocated afterwards
        are NUMA-local.  Returns the node or -1."""
        node = C.c_int(-1)
        check(self._L.feddb200_bind_host_numa(self._h, C.byref(node)))
        return int(node.value)

    def close(self):
        if getattr(self, "_h", None):
            self._L.feddb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scatter_mode(self, mode):
        check(self._L.feddb200_set_scatter_mode(self._h, _MODES[mode] if isinstance(mode, str) else int(mode)))

    @property
    def scatter_mode(self) -> int:
        return self._L.feddb200_get_scatter_mode(self._h)

    def set_row_phase(self, phase: int):
        """0 all rows, 1 geometry + ghost rows, 2 owned rows, 3 geometry only, 4 ghost rows only (feddb200_set_row_phase)."""
        check(self._L.feddb200_set_row_phase(self._h, int(phase)))

    def synchronize(self):
        check(self._L.feddb200_synchronize(self._h))

    @property
    def launches(self) -> int:
        return int(self._L.feddb200_launch_count(self._h))

    def empty_values(self, n: int):
        import torch
        return torch.empty(int(n), dtype=torch.float64, device=f"cuda:{self.device}")

    def scale_d(self, values, alpha: float):
        check(self._L.feddb200_scale_d(self._h, ptr(values), values.numel(), float(alpha)))

    def unpack_add_d(self, values, recv, slots, n=None):
        """values[slots[k]] += recv[k]; `recv` is a tensor or a raw device pointer (then `n` values)."""
        check(self._L.feddb200_unpack_add_d(self._h, ptr(values), ptr(recv), ptr(slots), recv.numel() if n is None else int(n)))

    # ---- peer memory (CUDA IPC) and ghost-row targets: the fused ghost-row exchange --------------------
    def ipc_alloc(self, nbytes: int):
        """Device buffer that other processes can map: returns (device pointer, 64-byte handle)."""
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        check(self._L.feddb200_ipc_alloc(self._h, int(nbytes), C.byref(p), C.cast(h, C.c_void_p)))
        return int(p.value), bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        check(self._L.feddb200_ipc_open(self._h, C.cast(h, C.c_void_p), C.byref(p)))
        return int(p.value)

    def ipc_close(self, p: int):
        check(self._L.feddb200_ipc_close(self._h, C.c_void_p(p)))

    def free_d(self, p: int):
        check(self._L.feddb200_dev_free(self._h, C.c_void_p(p)))

    def set_ghost_targets(self, seg_begin=None, seg_ptr=None):
        """Ghost rows of the next gather assemblies go to seg_ptr[o] + (offset - seg_begin[o]) (feddb200_set_ghost_targets);
        no arguments: back to the values array."""
        if seg_begin is None:
            check(self._L.feddb200_set_ghost_targets(self._h, 0, None, None))
            return
        n = len(seg_ptr)
        b = (C.c_int64 * (n + 1))(*[int(x) for x in seg_begin])
        q = (C.c_void_p * n)(*[C.c_void_p(int(x)) if x else C.c_void_p(None) for x in seg_ptr])
        check(self._L.feddb200_set_ghost_targets(self._h, n, C.cast(b, C.c_void_p), C.cast(q, C.c_void_p)))


class DeviceCsr:
    """A dof-level CSR matrix resident on the GPU: int64 rowptr, int32 colind (ascending per row), f64 values (torch tensors)."""

    def __init__(self, rowptr, colind, values, n_cols):
        self.rowptr, self.colind, self.values, self.n_cols = rowptr, colind, values, int(n_cols)

    @property
    def n_rows(self):
        return self.rowptr.numel() - 1

    @classmethod
    def from_host(cls, ctx, rowptr, colind, values, n_cols):
        import torch
        dev = f"cuda:{ctx.device}"
        return cls(torch.from_numpy(np.ascontiguousarray(rowptr, dtype=np.int64)).to(dev),
                   torch.from_numpy(np.ascontiguousarray(colind, dtype=np.int32)).to(dev),
                   values if hasattr(values, "data_ptr") else torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(dev),
                   n_cols)

    def to_host(self):
        return self.rowptr.cpu().numpy(), self.colind.cpu().numpy(), self.values.cpu().numpy()


def csr_add(ctx: "Context", alpha: float, A: DeviceCsr, beta: float, B: DeviceCsr) -> DeviceCsr:
    """Matrix::addMatrix / TwoMatrixAdd (Matrix_def.hpp:281-287): alpha*A + beta*B on the union pattern, on the device."""
    import torch
    if A.n_rows != B.n_rows or A.n_cols != B.n_cols:
        raise LogicError("addMatrix: the matrices live on different maps")
    dev = A.rowptr.device
    rpC = torch.empty(A.n_rows + 1, dtype=torch.int64, device=dev)
    nnz = C.c_int64(0)
    check(ctx._L.feddb200_csr_add_symbolic_d(ctx._h, A.n_rows, ptr(A.rowptr), ptr(A.colind), ptr(B.rowptr), ptr(B.colind),
                                             ptr(rpC), C.byref(nnz)))
    ciC = torch.empty(nnz.value, dtype=torch.int32, device=dev)
    vC = torch.empty(nnz.value, dtype=torch.float64, device=dev)
    check(ctx._L.feddb200_csr_add_numeric_d(ctx._h, A.n_rows, float(alpha), ptr(A.rowptr), ptr(A.colind), ptr(A.values), float(beta),
                                            ptr(B.rowptr), ptr(B.colind), ptr(B.values), ptr(rpC), ptr(ciC), ptr(vC)))
    return DeviceCsr(rpC, ciC, vC, A.n_cols)


def block_merge(ctx: "Context", blocks) -> DeviceCsr:
    """BlockMatrix::merge (BlockMatrix_def.hpp:119-289): blocks[i][j] is a DeviceCsr or None; returns the monolithic matrix."""
    import torch
    nb = len(blocks)
    if any(len(row) != nb for row in blocks):
        raise LogicError("merge: the block matrix must be square")
    n_rows, n_cols = [None] * nb, [None] * nb
    for i in range(nb):
        for j in range(nb):
            b = blocks[i][j]
            if b is None:
                continue
            if n_rows[i] not in (None, b.n_rows) or n_cols[j] not in (None, b.n_cols):
                raise LogicError("merge: blocks of one block row / column have different sizes")
            n_rows[i], n_cols[j] = b.n_rows, b.n_cols
    if any(x is None for x in n_rows) or any(x is None for x in n_cols):
        raise LogicError("merge: a block row or column has no block")
    some = next(b for row in blocks for b in row if b is not None)
    dev = some.rowptr.device
    nr = (C.c_int64 * nb)(*n_rows)
    ncl = (C.c_int32 * nb)(*n_cols)

    def parr(attr):
        return (C.c_void_p * (nb * nb))(*[C.c_void_p(getattr(blocks[i][j], attr).data_ptr()) if blocks[i][j] is not None else C.c_void_p(None)
                                          for i in range(nb) for j in range(nb)])

    rp, ci, v = parr("rowptr"), parr("colind"), parr("values")
    total = int(sum(n_rows))
    rpM = torch.empty(total + 1, dtype=torch.int64, device=dev)
    nnz = C.c_int64(0)
    vp = lambda a: C.cast(a, C.c_void_p)
    check(ctx._L.feddb200_block_merge_symbolic_d(ctx._h, nb, vp(nr), vp(ncl), vp(rp), ptr(rpM), C.byref(nnz)))
    ciM = torch.empty(nnz.value, dtype=torch.int32, device=dev)
    vM = torch.empty(nnz.value, dtype=torch.float64, device=dev)
    check(ctx._L.feddb200_block_merge_numeric_d(ctx._h, nb, vp(nr), vp(ncl), vp(rp), vp(ci), vp(v), ptr(rpM), ptr(ciM), ptr(vM)))
    return DeviceCsr(rpM, ciM, vM, int(sum(n_cols)))


class Mesh:
    """Uploaded connectivity + repeated points (feddb200_mesh_upload)."""

    def __init__(self, ctx: Context, dim: int, conn: np.ndarray, coords: np.ndarray):
        self.ctx = ctx
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if conn.ndim != 2 or coords.ndim != 2 or coords.shape[1] != dim:
            raise _lib.LogicError("Mesh: conn must be [ne, nloc] and coords [nn, dim]")
        self.dim, self.nloc = int(dim), int(conn.shape[1])
        self.ne, self.nn = int(conn.shape[0]), int(coords.shape[0])
        h = C.c_void_p()
        check(ctx._L.feddb200_mesh_upload(ctx._h, C.byref(h), self.dim, self.nloc, self.ne, ptr(conn), self.nn,
                                          ptr(coords)))
        self._h = h

    def update_coords(self, coords: np.ndarray):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        assert coords.shape == (self.nn, self.dim)
        check(self.ctx._L.feddb200_mesh_update_coords(self.ctx._h, self._h, ptr(coords)))

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._L.feddb200_mesh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pattern:
    """Node-level CSR pattern + scatter/gather maps (feddb200_pattern_build)."""

    def __init__(self, ctx: Context, row_mesh: Mesh, col_mesh: Mesh | None = None, row_lid=None, n_rows=0,
                 n_owned_rows=0, col_lid=None, n_cols=0, extra_row=None, extra_col=None):
        self.ctx = ctx
        self.row_mesh = row_mesh
        self.col_mesh = col_mesh or row_mesh
        as32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        row_lid, col_lid, extra_row, extra_col = as32(row_lid), as32(col_lid), as32(extra_row), as32(extra_col)
        n_extra = 0 if extra_row is None else int(extra_row.size)
        h = C.c_void_p()
        check(ctx._L.feddb200_pattern_build(ctx._h, C.byref(h), row_mesh._h, self.col_mesh._h, int(n_rows),
                                            int(n_owned_rows), ptr(row_lid), int(n_cols), ptr(col_lid), n_extra,
                                            ptr(extra_row), ptr(extra_col)))
        self._h = h
        v = [C.c_int64() for _ in range(5)]
        ml = C.c_int32()
        check(ctx._L.feddb200_pattern_info(h, *[C.byref(x) for x in v], C.byref(ml), None))
        self.n_rows, self.n_owned_rows, self.n_cols, self.nnz_nodes, self.nnz_owned_nodes = [int(x.value) for x in v]
        self.max_row_len = int(ml.value)
        self.dim = row_mesh.dim

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._L.feddb200_pat_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_colours(self) -> int:
        nc = C.c_int32()
        check(self.ctx._L.feddb200_pattern_info(self._h, None, None, None, None, None, None, C.byref(nc)))
        return int(nc.value)

    def nodes(self):
        rowptr = np.empty(self.n_rows + 1, dtype=np.int64)
        colind = np.empty(self.nnz_nodes, dtype=np.int32)
        check(self.ctx._L.feddb200_pattern_get_nodes(self.ctx._h, self._h, ptr(rowptr), ptr(colind)))
        return rowptr, colind

    def nnz(self, row_dofs=1, col_dofs=1, mode=BLOCK_SCALAR) -> int:
        n = int(self.ctx._L.feddb200_pattern_nnz(self._h, row_dofs, col_dofs, mode))
        if n < 0:
            raise _lib.LogicError("unsupported dof layout")
        return n

    def nnz_owned(self, row_dofs=1, col_dofs=1, mode=BLOCK_SCALAR) -> int:
        return int(self.ctx._L.feddb200_pattern_nnz_owned(self._h, row_dofs, col_dofs, mode))

    def expand(self, row_dofs=1, col_dofs=1, mode=BLOCK_SCALAR):
        """dof-level CSR (rowptr int64, colind int32) of owned + ghost rows."""
        nnz = self.nnz(row_dofs, col_dofs, mode)
        rowptr = np.empty(self.n_rows * row_dofs + 1, dtype=np.int64)
        colind = np.empty(nnz, dtype=np.int32)
        check(self.ctx._L.feddb200_pattern_expand(self.ctx._h, self._h, row_dofs, col_dofs, mode, ptr(rowptr),
                                                  ptr(colind)))
        return rowptr, colind

    # ---- device-resident assembly (values: CUDA float64 tensor, fully overwritten) ----
    def assemble_laplace_d(self, values, vec_field=False):
        check(self.ctx._L.feddb200_assemble_laplace_d(self.ctx._h, self._h, int(bool(vec_field)), ptr(values)))

    def assemble_mass_d(self, values, vec_field=False):
        check(self.ctx._L.feddb200_assemble_mass_d(self.ctx._h, self._h, int(bool(vec_field)), ptr(values)))

    def stress_points(self, conn, coords):
        """Physical quadrature points x_k = B q_k + p_1 of FE::assemblyStress for every element, [ne][nq][dim], summed in
        the reference's order (FE_def.hpp:2479-2484, 2599-2608); conn / coords are the host arrays of the mesh."""
        dim, nloc = self.dim, conn.shape[1]
        nq = C.c_int(0)
        check(self.ctx._L.feddb200_stress_quadrature(dim, nloc, C.byref(nq), None, None))
        q = np.empty((nq.value, dim), dtype=np.float64)
        check(self.ctx._L.feddb200_stress_quadrature(dim, nloc, C.byref(nq), ptr(q), None))
        p = coords[conn[:, :dim + 1]]                                   # [ne][dim+1][dim]
        B = np.stack([p[:, j + 1, :] - p[:, 0, :] for j in range(dim)], axis=2)   # B[e][i][j] = x_{j+1}[i] - x_0[i]
        xyz = np.zeros((conn.shape[0], nq.value, dim))
        for r in range(dim):
            xyz += B[:, None, :, r] * q[None, :, r, None]
        xyz += p[:, None, 0, :]
        return xyz

    def assemble_stress_d(self, values, coef=1.0):
        """FE::assemblyStress; coef = the constant value of the coefficient function, or a CUDA tensor [ne * nq]."""
        if hasattr(coef, "data_ptr"):
            check(self.ctx._L.feddb200_assemble_stress_d(self.ctx._h, self._h, 1.0, ptr(coef), ptr(values)))
        else:
            check(self.ctx._L.feddb200_assemble_stress_d(self.ctx._h, self._h, float(coef), None, ptr(values)))

    def assemble_stress(self, coef=1.0):
        """Host variant: coef = constant or array [ne][nq] of func at the points of stress_points()."""
        d = self.dim
        out = np.empty(self.nnz(d, d, BLOCK_FULL), dtype=np.float64)
        if np.ndim(coef) == 0:
            check(self.ctx._L.feddb200_assemble_stress(self.ctx._h, self._h, float(coef), None, 0, ptr(out)))
        else:
            cf = np.ascontiguousarray(coef, dtype=np.float64).ravel()
            check(self.ctx._L.feddb200_assemble_stress(self.ctx._h, self._h, 1.0, ptr(cf), cf.size, ptr(out)))
        return out

    def assemble_bdstab_d(self, values):
        check(self.ctx._L.feddb200_assemble_bdstab_d(self.ctx._h, self._h, ptr(values)))

    def assemble_bdstab(self):
        out = np.empty(self.nnz(), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_bdstab(self.ctx._h, self._h, ptr(out)))
        return out

    def assemble_linelas_d(self, values, lam, mu):
        check(self.ctx._L.feddb200_assemble_linelas_d(self.ctx._h, self._h, float(lam), float(mu), ptr(values)))

    def assemble_advection_d(self, values, u):
        check(self.ctx._L.feddb200_assemble_advection_d(self.ctx._h, self._h, ptr(u), ptr(values)))

    def assemble_advection_in_u_d(self, values, u):
        check(self.ctx._L.feddb200_assemble_advection_in_u_d(self.ctx._h, self._h, ptr(u), ptr(values)))

    def assemble_ns_jacobian_d(self, values, u, rho, nu, newton=True):
        check(self.ctx._L.feddb200_assemble_ns_jacobian_d(self.ctx._h, self._h, float(rho), float(nu), ptr(u),
                                                          int(bool(newton)), ptr(values)))

    def assemble_rhs(self, value_func, deg_func=0, vec_field=False):
        """FE::assemblyRHS with a constant source (FE_def.hpp:4694-4766): load vector in pattern-row order (host array)."""
        f = np.zeros(3)
        vf = np.atleast_1d(np.asarray(value_func, dtype=np.float64))
        f[:vf.size] = vf
        out = np.empty(self.n_rows * (self.dim if vec_field else 1), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_rhs(self.ctx._h, self._h, int(bool(vec_field)), int(deg_func), ptr(f), ptr(out)))
        return out

    def set_dirichlet_rows_d(self, values, node_mask, row_dofs=1, col_dofs=1, mode=BLOCK_SCALAR, diagonal_block=True):
        """BCBuilder::setDirichletBC on resident values (BCBuilder_def.hpp:618-709): node_mask = uint8 CUDA tensor, one
        byte per owned row node, bit a set when dof a is a Dirichlet dof."""
        check(self.ctx._L.feddb200_set_dirichlet_rows_d(self.ctx._h, self._h, int(row_dofs), int(col_dofs), int(mode),
                                                       ptr(node_mask), int(bool(diagonal_block)), ptr(values)))

    def set_dirichlet_rhs_d(self, rhs, node_mask, bc_values, dofs=1):
        """BCBuilder::setRHS, Dirichlet part, on a resident right-hand side (BCBuilder_def.hpp:93-166): rhs[dofs*I + a] =
        bc_values[dofs*I + a] where bit a of node_mask[I] is set (CUDA tensors; bc_values = the boundary function at the nodes)."""
        check(self.ctx._L.feddb200_set_dirichlet_rhs_d(self.ctx._h, int(node_mask.numel()), int(dofs), ptr(node_mask), ptr(bc_values), ptr(rhs)))

    # ---- host-buffer forms (H2D / D2H inside the call) ----
    def assemble_laplace(self, vec_field=False):
        out = np.empty(self.nnz(self.dim, self.dim, BLOCK_DIAG) if vec_field else self.nnz(), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_laplace(self.ctx._h, self._h, int(bool(vec_field)), ptr(out)))
        return out

    def assemble_mass(self, vec_field=False):
        out = np.empty(self.nnz(self.dim, self.dim, BLOCK_DIAG) if vec_field else self.nnz(), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_mass(self.ctx._h, self._h, int(bool(vec_field)), ptr(out)))
        return out

    def assemble_linelas(self, lam, mu, out=None):
        if out is None:
            out = np.empty(self.nnz(self.dim, self.dim, BLOCK_FULL), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_linelas(self.ctx._h, self._h, float(lam), float(mu), ptr(out)))
        return out

    def assemble_advection(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(self.nnz(self.dim, self.dim, BLOCK_DIAG), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_advection(self.ctx._h, self._h, ptr(u), ptr(out)))
        return out

    def assemble_advection_in_u(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(self.nnz(self.dim, self.dim, BLOCK_FULL), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_advection_in_u(self.ctx._h, self._h, ptr(u), ptr(out)))
        return out

    def assemble_ns_jacobian(self, u, rho, nu, newton=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(self.nnz(self.dim, self.dim, BLOCK_FULL), dtype=np.float64)
        check(self.ctx._L.feddb200_assemble_ns_jacobian(self.ctx._h, self._h, float(rho), float(nu), ptr(u),
                                                        int(bool(newton)), ptr(out)))
        return out


def assemble_div_divT_d(ctx: Context, patB: Pattern | None, patBT: Pattern | None, valuesB=None, valuesBT=None):
    check(ctx._L.feddb200_assemble_div_divT_d(ctx._h, patB._h if patB else None, patBT._h if patBT else None,
                                              ptr(valuesB), ptr(valuesBT)))


def assemble_div_divT(ctx: Context, patB: Pattern, patBT: Pattern):
    dim = patB.dim
    vB = np.empty(patB.nnz(1, dim, BLOCK_FULL), dtype=np.float64)
    vBT = np.empty(patBT.nnz(dim, 1, BLOCK_FULL), dtype=np.float64)
    check(ctx._L.feddb200_assemble_div_divT(ctx._h, patB._h, patBT._h, ptr(vB), ptr(vBT)))
    return vB, vBT
