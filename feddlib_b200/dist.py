"""Multi-GPU element partition: ownership, column maps and the ghost-row exchange plan.

What this replaces in the reference: every rank assembles its own elements; contributions to rows
owned by another rank are shipped to the owner and ADDed inside `Matrix::fillComplete`
(core/LinearAlgebra/Matrix_def.hpp:192-199 -> Tpetra globalAssemble), and the
column map is rebuilt as "owned GIDs in domain-map order, then remote GIDs grouped by owning rank,
ascending GID".  Here the same happens with one process per GPU:

  one-time (host, torch.distributed for the plumbing; NCCL on GPUs, gloo in the CPU tests)
    1. rows: owned nodes in repeated order (= Map::buildUniqueMap order, Map_def.hpp:201-206), then
       ghost nodes grouped by owner -> the ghost values of one destination are one contiguous segment
    2. preliminary node pattern -> ghost-row structure (row gid, col gid, col owner) -> owners
    3. column index space = Tpetra column map (owned, then remotes by (owner, gid)) followed by
       ghost-only columns that never reach the owned rows
    4. final device pattern with the received entries as extra entries
    5. senders publish their ghost rows in final CSR order; receivers resolve them to value slots
  per assembly (device)
    values[ghost part] --all_to_all_single (NCCL over NVLink)--> owners --unpack-add kernel--> CSR

The plan builder takes the node-pattern builder as a callable so the identical host logic runs in
the CPU tests (world_size 2, gloo) with a numpy pattern builder.
"""
from __future__ import annotations

import os

import numpy as np

from ._lib import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR


# ------------------------------------------------------------------------------------------------
# small collective helpers on numpy int64 / float64 arrays
# ------------------------------------------------------------------------------------------------
class Comm:
    """torch.distributed wrapper; device=None -> CPU tensors (gloo), else CUDA tensors (nccl)."""

    def __init__(self, rank: int, size: int, device=None):
        self.rank, self.size, self.device = rank, size, device

    def _t(self, a, dtype):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        return t.to(self.device) if self.device is not None else t

    def alltoallv(self, chunks, np_dtype=np.int64):
        """chunks[d] = 1-D array for rank d.  Returns list of arrays received from every rank."""
        import torch
        import torch.distributed as dist
        tdt = torch.int64 if np_dtype == np.int64 else torch.float64
        if self.size == 1:
            return [np.ascontiguousarray(chunks[0], dtype=np_dtype)]
        counts = torch.tensor([len(c) for c in chunks], dtype=torch.int64)
        if self.device is not None:
            counts = counts.to(self.device)
        rcounts = torch.empty_like(counts)
        dist.all_to_all_single(rcounts, counts)
        rc = rcounts.cpu().tolist()
        send = self._t(np.concatenate([np.asarray(c, dtype=np_dtype) for c in chunks]) if sum(len(c) for c in chunks)
                       else np.zeros(0, dtype=np_dtype), tdt)
        recv = torch.empty(int(sum(rc)), dtype=tdt, device=send.device)
        dist.all_to_all_single(recv, send, rc, [len(c) for c in chunks])
        out = recv.cpu().numpy()
        offs = np.concatenate([[0], np.cumsum(rc)]).astype(np.int64)
        return [out[offs[i]:offs[i + 1]] for i in range(self.size)]


def numpy_node_pattern(conn_r, conn_c, row_lid, n_rows, col_lid, extra_row=None, extra_col=None):
    """Reference (host) node-pattern builder with the semantics of feddb200_pattern_build."""
    nr, nc = conn_r.shape[1], conn_c.shape[1]
    r = row_lid[conn_r] if row_lid is not None else conn_r
    c = col_lid[conn_c] if col_lid is not None else conn_c
    rows = np.repeat(r, nc, axis=1).ravel().astype(np.int64)
    cols = np.tile(c, (1, nr)).ravel().astype(np.int64)
    keep = rows >= 0
    rows, cols = rows[keep], cols[keep]
    if extra_row is not None and len(extra_row):
        rows = np.concatenate([rows, np.asarray(extra_row, dtype=np.int64)])
        cols = np.concatenate([cols, np.asarray(extra_col, dtype=np.int64)])
    key = np.unique(rows << 32 | cols)
    rows_u, cols_u = key >> 32, key & 0xffffffff
    rowptr = np.searchsorted(rows_u, np.arange(n_rows + 1)).astype(np.int64)
    return rowptr, cols_u.astype(np.int32)


# ------------------------------------------------------------------------------------------------
# the plan
# ------------------------------------------------------------------------------------------------
class HaloPlan:
    """Ownership, index spaces and ghost-row exchange plan of one rank (square node pattern)."""

    def __init__(self, comm: Comm, gid_rep, owner, pattern_fn):
        """
        gid_rep [nn]  global id of every repeated node;  owner [nn]  owning rank of every repeated node
        pattern_fn(row_lid, n_rows, n_owned, col_lid, n_cols, extra_row, extra_col) -> (rowptr, colind)
                      builds the node pattern of this rank's elements (device builder or numpy_node_pattern)
        """
        self.comm = comm
        rank, size = comm.rank, comm.size
        gid_rep = np.ascontiguousarray(gid_rep, dtype=np.int64)
        owner = np.ascontiguousarray(owner, dtype=np.int64)
        nn = gid_rep.size
        self.gid_rep, self.owner = gid_rep, owner

        # 1. rows
        owned = owner == rank
        owned_ids = np.flatnonzero(owned)                       # repeated order -> unique-map order
        ghost_ids = np.flatnonzero(~owned)
        ghost_ids = ghost_ids[np.lexsort((gid_rep[ghost_ids], owner[ghost_ids]))]   # by (owner, gid)
        self.n_owned, self.n_ghost = owned_ids.size, ghost_ids.size
        self.n_rows = self.n_owned + self.n_ghost
        row_lid = np.empty(nn, dtype=np.int32)
        row_lid[owned_ids] = np.arange(self.n_owned, dtype=np.int32)
        row_lid[ghost_ids] = self.n_owned + np.arange(self.n_ghost, dtype=np.int32)
        self.row_lid = row_lid
        self.unique_gids = gid_rep[owned_ids]
        self.ghost_row_gids = gid_rep[ghost_ids]
        self.ghost_row_owner = owner[ghost_ids]
        rep_of_row = np.empty(self.n_rows, dtype=np.int64)
        rep_of_row[row_lid] = np.arange(nn)

        # 2. preliminary pattern (columns = repeated local ids) -> ghost-row structure to the owners
        rp0, ci0 = pattern_fn(row_lid, self.n_rows, self.n_owned, None, nn, None, None)
        g0, g1 = rp0[self.n_owned], rp0[self.n_rows]
        grow = np.repeat(np.arange(self.n_owned, self.n_rows), np.diff(rp0[self.n_owned:]))  # row lid per ghost entry
        gcol_rep = ci0[g0:g1].astype(np.int64)
        dest = owner[rep_of_row[grow]]
        send = []
        for d in range(size):
            m = dest == d
            send.append(np.stack([gid_rep[rep_of_row[grow[m]]], gid_rep[gcol_rep[m]], owner[gcol_rep[m]]], axis=1).ravel()
                        if m.any() else np.zeros(0, dtype=np.int64))
        recv = [x.reshape(-1, 3) for x in comm.alltoallv(send)]
        rec = np.concatenate(recv, axis=0) if len(recv) else np.zeros((0, 3), dtype=np.int64)

        # 3. column index space
        rep_sorted = np.argsort(gid_rep, kind="stable")
        gid_sorted = gid_rep[rep_sorted]

        def rep_of_gid(g):      # repeated local id of a gid, -1 if this rank does not have the node
            pos = np.searchsorted(gid_sorted, g)
            pos = np.minimum(pos, max(nn - 1, 0))
            hit = (gid_sorted[pos] == g) if nn else np.zeros(g.shape, dtype=bool)
            return np.where(hit, rep_sorted[pos], -1)

        owned_cols_in_owned_rows = np.zeros(nn, dtype=bool)
        owned_cols_in_owned_rows[ci0[:g0]] = True               # repeated nodes appearing in owned rows
        remote_gid = gid_rep[~owned & owned_cols_in_owned_rows]
        remote_own = owner[~owned & owned_cols_in_owned_rows]
        if rec.size:
            m = rec[:, 2] != rank                               # received columns owned elsewhere
            remote_gid = np.concatenate([remote_gid, rec[m, 1]])
            remote_own = np.concatenate([remote_own, rec[m, 2]])
        if remote_gid.size:
            rg, idx = np.unique(remote_gid, return_index=True)
            ro = remote_own[idx]
            order = np.lexsort((rg, ro))
            remote_gid, remote_own = rg[order], ro[order]
        self.colmap_gids = np.concatenate([self.unique_gids, remote_gid])      # the Tpetra column map
        self.n_colmap = self.colmap_gids.size
        # ghost-only columns (appear only in ghost rows): appended behind the column map
        col_lid = np.full(nn, -1, dtype=np.int64)
        col_lid[owned_ids] = np.arange(self.n_owned)
        if remote_gid.size:
            r_rep = rep_of_gid(remote_gid)
            has = r_rep >= 0
            col_lid[r_rep[has]] = self.n_owned + np.flatnonzero(has)
        rest = np.flatnonzero(col_lid < 0)
        col_lid[rest] = self.n_colmap + np.arange(rest.size)
        self.n_cols = self.n_colmap + rest.size
        self.col_lid = col_lid.astype(np.int32)

        # 4. extra entries in owned rows
        if rec.size:
            colmap_sorted = np.argsort(self.colmap_gids, kind="stable")
            cm_gid_sorted = self.colmap_gids[colmap_sorted]
            er = row_lid[rep_of_gid(rec[:, 0])]
            ec = colmap_sorted[np.searchsorted(cm_gid_sorted, rec[:, 1])]
            assert np.all(er >= 0) and np.all(er < self.n_owned), "ghost row sent to a rank that does not own it"
            self.extra_row, self.extra_col = er.astype(np.int32), ec.astype(np.int32)
        else:
            self.extra_row = self.extra_col = np.zeros(0, dtype=np.int32)
        self.rowptr, self.colind = pattern_fn(row_lid, self.n_rows, self.n_owned, self.col_lid, self.n_cols,
                                              self.extra_row, self.extra_col)
        rp, ci = self.rowptr, self.colind
        self.nnz_owned_nodes, self.nnz_nodes = int(rp[self.n_owned]), int(rp[self.n_rows])

        # 5. ghost rows in final CSR order -> owners resolve them to node-level (row, position)
        gid_of_col = np.empty(self.n_cols, dtype=np.int64)
        gid_of_col[self.col_lid] = gid_rep
        gid_of_col[: self.n_colmap] = self.colmap_gids
        glen = np.diff(rp[self.n_owned:])
        grow = np.repeat(np.arange(self.n_owned, self.n_rows), glen)
        gq = np.arange(rp[self.n_owned], rp[self.n_rows]) - np.repeat(rp[self.n_owned:-1], glen)  # position in ghost row
        gcol_gid = gid_of_col[ci[rp[self.n_owned]:rp[self.n_rows]]]
        dest = owner[rep_of_row[grow]]
        self.send_counts_nodes = np.array([(dest == d).sum() for d in range(size)], dtype=np.int64)
        assert np.all(np.diff(dest) >= 0), "ghost rows are not grouped by owner"
        send = []
        for d in range(size):
            m = dest == d
            send.append(np.stack([gid_rep[rep_of_row[grow[m]]], gcol_gid[m], np.repeat(glen, glen)[m], gq[m]], axis=1).ravel()
                        if m.any() else np.zeros(0, dtype=np.int64))
        recv = [x.reshape(-1, 4) for x in comm.alltoallv(send)]
        self.recv_counts_nodes = np.array([x.shape[0] for x in recv], dtype=np.int64)
        rec = np.concatenate(recv, axis=0) if len(recv) else np.zeros((0, 4), dtype=np.int64)
        if rec.size:
            I = row_lid[rep_of_gid(rec[:, 0])].astype(np.int64)
            colmap_sorted = np.argsort(self.colmap_gids, kind="stable")
            c = colmap_sorted[np.searchsorted(self.colmap_gids[colmap_sorted], rec[:, 1])]
            # position of column c in owned row I (rows are sorted by column index)
            key_all = (np.repeat(np.arange(self.n_rows), np.diff(rp)).astype(np.int64) << 32) | ci.astype(np.int64)
            slot = np.searchsorted(key_all, (I << 32) | c)
            assert np.array_equal(key_all[slot], (I << 32) | c), "received entry missing from the owner's pattern"
            self.recv_row, self.recv_pos = I, slot - rp[I]
            self.recv_len_sender, self.recv_q = rec[:, 2], rec[:, 3]
        else:
            z = np.zeros(0, dtype=np.int64)
            self.recv_row = self.recv_pos = self.recv_len_sender = self.recv_q = z
        self._slots = {}

    # -- dof-level helpers ---------------------------------------------------------------------
    @staticmethod
    def _factor(rd, cd, mode):
        return 1 if mode == BLOCK_SCALAR else (rd if mode == BLOCK_DIAG else rd * cd)

    def split_sizes(self, rd, cd, mode):
        f = self._factor(rd, cd, mode)
        return (self.send_counts_nodes * f).tolist(), (self.recv_counts_nodes * f).tolist()

    def recv_slots(self, rd, cd, mode):
        """Value slots (into this rank's dof-level CSR values) of every received value, in arrival order.

        A sender ships each ghost node row as one block of f*len values in its own CSR order
        (a, q, b); the receiver adds element (a, q, b) to rd*cd*base_I + a*cd*L_I + cd*p + b."""
        key = (rd, cd, mode)
        if key in self._slots:
            return self._slots[key]
        f = self._factor(rd, cd, mode)
        rp = self.rowptr
        I, p = self.recv_row, self.recv_pos
        n = I.size
        per = 1 if mode != BLOCK_FULL else cd
        nrow_dofs = 1 if mode == BLOCK_SCALAR else rd
        base, L = rp[I], rp[I + 1] - rp[I]
        # arrival order within one sender: ghost rows in order, within a row (a, q, b).  Build the arrival index
        # of every (entry, a, b) and scatter.
        out = np.empty(n * f, dtype=np.int64)
        # start offset (in values) of each received entry's row block within the stream
        # entries of one ghost row are consecutive (q = 0..len-1); block start = cumulative f*len of earlier rows
        first = self.recv_q == 0
        row_block_start = np.cumsum(np.where(first, self.recv_len_sender * f, 0)) - np.where(first, self.recv_len_sender * f, 0)
        row_block_start = np.maximum.accumulate(np.where(first, row_block_start, 0))
        Ls, q = self.recv_len_sender, self.recv_q
        for a in range(nrow_dofs):
            for b in range(per):
                arrival = row_block_start + a * per * Ls + per * q + b
                out[arrival] = f * base + a * per * L + per * p + b
        self._slots[key] = out
        return out


    # -- vectors on the pattern rows (FE::assemblyRHS + exportFromVector(..., "Add"), Problem_def.hpp:184-216) -------
    def vec_split_sizes(self, dofs):
        """(send, recv) counts per rank of the ghost part of a row vector with `dofs` entries per node."""
        send = np.array([(self.ghost_row_owner == d).sum() for d in range(self.comm.size)], dtype=np.int64) * dofs
        first = self.recv_q == 0
        off = np.concatenate([[0], np.cumsum(self.recv_counts_nodes)])
        recv = np.array([first[off[s]:off[s + 1]].sum() for s in range(self.comm.size)], dtype=np.int64) * dofs
        return send.tolist(), recv.tolist()

    def vec_recv_slots(self, dofs):
        """Entry (into this rank's row vector, owned part) of every received vector value, in arrival order: a sender
        ships its ghost rows in order, node-wise interleaved dofs; the first pattern entry of each received ghost row
        names the owner's row."""
        I = self.recv_row[self.recv_q == 0]
        return (dofs * I[:, None] + np.arange(dofs)[None, :]).ravel().astype(np.int64)


class DistributedMatrixAssembler:
    """Device-side driver: per-rank mesh + pattern from a HaloPlan, assembly + ghost exchange."""

    def __init__(self, ctx, dim, conn, coords, gid_rep, owner, rank, size):
        import os
        import time
        import torch
        from .engine import Mesh, Pattern
        self.ctx, self.dim, self.rank, self.size = ctx, dim, rank, size
        self.conn, self.coords = conn, coords
        t0 = time.perf_counter()
        self.mesh = Mesh(ctx, dim, conn, coords)
        self._pats = []
        self.timing = {"mesh_upload_s": time.perf_counter() - t0, "pattern_calls_s": 0.0}

        def pattern_fn(row_lid, n_rows, n_owned, col_lid, n_cols, er, ec):
            t1 = time.perf_counter()
            for p in self._pats:            # the preliminary pattern is not needed once the next one is built
                p.close()
            pat = Pattern(ctx, self.mesh, None, row_lid, n_rows, n_owned, col_lid, n_cols, er, ec)
            self._pats.append(pat)
            out = pat.nodes()
            self.timing["pattern_calls_s"] += time.perf_counter() - t1
            return out

        dev = torch.device(f"cuda:{ctx.device}")
        # the plan is built by the C++ host code a FEDDLib build links (halo.cpp); HaloPlan (numpy) is its executable
        # specification in the CPU tests
        from .halo import NativeHaloPlan
        t0 = time.perf_counter()
        self.plan = NativeHaloPlan(Comm(rank, size, dev), gid_rep, owner, pattern_fn)
        self.timing["plan_total_s"] = time.perf_counter() - t0
        self.pat = self._pats[-1]
        for p in self._pats[:-1]:
            p.close()
        self._dev = dev
        self._slot_t = {}
        self._recv_buf = {}
        if os.environ.get("FEDDB200_TIMING"):
            import sys
            print(f"[feddb200 timing] rank {rank}: {self.timing}", file=sys.stderr, flush=True)

    def _main_stream(self):
        """The stream the engine launches on (bound in Context.__init__): the exchange is ordered against IT, not against whatever
        torch's current stream happens to be when the call is made."""
        import torch
        s = getattr(self.ctx, "torch_stream", None)
        return s if s is not None else torch.cuda.current_stream(self._dev)

    def import_vector(self, u_unique, dofs):
        """MultiVector::importFromVector (core/LinearAlgebra/MultiVector_def.hpp:258-294), unique -> repeated, on the device: the
        velocity of the advection operators lives on the unique map between Newton steps and is read on the repeated map
        (problems/specific/NavierStokes_def.hpp:294).  u_unique: [dofs * n_owned] device tensor in unique-map order; returns the
        [dofs * nn] repeated vector.  One all-to-all-v of dofs * 8 bytes per ghost node."""
        import torch
        import torch.distributed as dist
        plan = self.plan
        nn = plan.gid_rep.size
        if not hasattr(self, "_imp"):
            rep = torch.from_numpy(np.asarray(plan.rep_of_row, dtype=np.int64)).to(self._dev)
            self._imp = {"rep": rep, "send_rows": torch.from_numpy(np.asarray(plan.import_send_rows, dtype=np.int64)).to(self._dev)}
        u2 = u_unique.reshape(-1, dofs)
        out = torch.empty((nn, dofs), dtype=torch.float64, device=self._dev)
        out[self._imp["rep"][: plan.n_owned]] = u2
        if self.size > 1:
            send = u2[self._imp["send_rows"]].contiguous()
            ssz = [int(x) for x in plan.import_send_counts]
            rsz = [int((plan.ghost_row_owner == d).sum()) for d in range(self.size)]
            recv = torch.empty((plan.n_ghost, dofs), dtype=torch.float64, device=self._dev)
            dist.all_to_all_single(recv, send, rsz, ssz)
            out[self._imp["rep"][plan.n_owned:]] = recv
        return out.reshape(-1)

    def _exchange_buffers(self, rd, cd, mode):
        import torch
        key = (rd, cd, mode)
        if key not in self._slot_t:
            self._slot_t[key] = torch.from_numpy(self.plan.recv_slots(rd, cd, mode)).to(self._dev)
            self._recv_buf[key] = torch.empty(self._slot_t[key].numel(), dtype=torch.float64, device=self._dev)
        return key

    def assemble_overlapped(self, values, rd, cd, mode, assemble):
        """`assemble()` launches one assembly of this pattern into `values`.  Ghost rows are assembled first and
        shipped (NCCL, side stream) while the owned rows are assembled; the received values are added last."""
        import torch
        import torch.distributed as dist
        if self.size == 1:
            assemble()
            return
        key = self._exchange_buffers(rd, cd, mode)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self._dev)
        ssz, rsz = self.plan.split_sizes(rd, cd, mode)
        send, recv = values[self.pat.nnz_owned(rd, cd, mode):], self._recv_buf[key]
        main = self._main_stream()
        self.ctx.set_row_phase(1)
        try:
            assemble()                      # geometry pre-pass + ghost rows
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                dist.all_to_all_single(recv, send, rsz, ssz)
            self.ctx.set_row_phase(2)
            assemble()                      # owned rows, concurrent with the exchange
        finally:
            self.ctx.set_row_phase(0)
        main.wait_stream(self._side)
        self._unpack(values, key, rsz)

    # ---- fused exchange over NVLink peer memory ------------------------------------------------------------
    def _peer_setup(self, key, rd, cd, mode):
        """Two receive buffers per rank (step parity), mapped into every sender (CUDA IPC).  A sender's segment of the
        receiver's buffer starts at the receiver's own arrival offset of that sender, so unpack_add is unchanged."""
        import torch.distributed as dist
        ssz, rsz = self.plan.split_sizes(rd, cd, mode)
        mine = [self.ctx.ipc_alloc(8 * max(1, sum(rsz))) for _ in range(2)]
        info = [None] * self.size
        dist.all_gather_object(info, ([h for _, h in mine], rsz))
        seg_ptr = [[0] * self.size for _ in range(2)]
        opened = []
        for o in range(self.size):
            if o == self.rank or ssz[o] == 0:
                continue
            handles, rsz_o = info[o]
            off = 8 * int(sum(rsz_o[:self.rank]))
            assert rsz_o[self.rank] == ssz[o]
            for par in range(2):
                base = self.ctx.ipc_open(handles[par])
                opened.append(base)
                seg_ptr[par][o] = base + off
        n_owned_vals = self.pat.nnz_owned(rd, cd, mode)
        seg_begin = [n_owned_vals + int(sum(ssz[:o])) for o in range(self.size + 1)]
        return {"recv": [p for p, _ in mine], "seg_ptr": seg_ptr, "seg_begin": seg_begin, "opened": opened, "step": 0}

    def assemble_fused(self, values, rd, cd, mode, assemble):
        """Like assemble_overlapped, but no collective moves data: the ghost-row kernels store their rows straight into
        the owners' receive buffers over NVLink (feddb200_set_ghost_targets) as they are computed.  A one-element
        all-reduce on the side stream is the cross-rank barrier between those stores and the owners' unpack_add; it
        runs under the assembly of the owned rows.  Receive buffers alternate with the step parity, so one barrier
        per assembly orders everything.  Gather mode, Laplace / elasticity operators, at most 8 ranks."""
        import torch
        import torch.distributed as dist
        if self.size == 1:
            assemble()
            return
        key = self._exchange_buffers(rd, cd, mode)
        if not hasattr(self, "_peer"):
            self._peer = {}
        if key not in self._peer:
            self._peer[key] = self._peer_setup(key, rd, cd, mode)
            self._flag = torch.zeros(1, dtype=torch.int32, device=self._dev)
        P = self._peer[key]
        par = P["step"] & 1
        P["step"] += 1
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self._dev)
        _, rsz = self.plan.split_sizes(rd, cd, mode)
        main = self._main_stream()
        # optional per-phase device times (self.phase_timing = True; read with phase_ms() after a synchronize)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if getattr(self, "phase_timing", False) else None
        if ev: ev[0].record(main)
        self.ctx.set_ghost_targets(P["seg_begin"], P["seg_ptr"][par])
        # (measured at 2 GPUs, M = 70: 2.75 ms with the ghost rows next to the owned rows, 2.62 ms with them in front -- the
        # concurrent launches take SMs from the owned rows for longer than they save; off by default)
        concurrent = getattr(self, "concurrent_ghost", os.environ.get("FEDDB200_CONCURRENT_GHOST", "0") != "0")
        try:
            if concurrent:
                # only the geometry pre-pass is a dependency of both row sets: the (small, under-filled) ghost-row launches and
                # the barrier behind them run on the side stream NEXT TO the owned rows instead of in front of them
                self.ctx.set_row_phase(3)
                assemble()                  # geometry pre-pass
                self._side.wait_stream(main)
                self.ctx.set_stream(self._side)
                self.ctx.set_row_phase(4)
                try:
                    assemble()              # ghost rows -> peer memory, side stream
                finally:
                    self.ctx.set_stream(main)
                self.ctx.set_ghost_targets()
                if ev: ev[1].record(self._side)
                with torch.cuda.stream(self._side):
                    dist.all_reduce(self._flag)  # barrier: every rank's ghost rows have been stored
                    if ev: ev[4].record(self._side)
            else:
                self.ctx.set_row_phase(1)
                assemble()                  # geometry pre-pass + ghost rows -> peer memory
                self.ctx.set_ghost_targets()
                if ev: ev[1].record(main)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    dist.all_reduce(self._flag)  # barrier: every rank's ghost rows have been stored
                    if ev: ev[4].record(self._side)
            self.ctx.set_row_phase(2)
            assemble()                      # owned rows
        finally:
            self.ctx.set_ghost_targets()
            self.ctx.set_row_phase(0)
        if ev: ev[2].record(main)
        main.wait_stream(self._side)
        slots = self._slot_t[key]
        off = 0
        for n in rsz:                       # one launch per sender, fixed order: deterministic sums
            if n:
                self.ctx.unpack_add_d(values, P["recv"][par] + 8 * off, slots[off:off + n], n)
            off += n
        if ev:
            ev[3].record(main)
            self._phase_ev = ev

    def phase_ms(self):
        """Device times of the last assemble_fused call (phase_timing = True): geometry + ghost rows, owned rows, the
        cross-rank barrier measured from the end of the ghost rows (it runs under the owned rows), the wait for it after the
        owned rows plus the unpack-add launches, and the whole call."""
        ev = getattr(self, "_phase_ev", None)
        if ev is None:
            return None
        return {"geometry_and_ghost_rows": ev[0].elapsed_time(ev[1]), "owned_rows": ev[1].elapsed_time(ev[2]),
                "barrier_after_ghost_rows": ev[1].elapsed_time(ev[4]), "barrier_wait_and_unpack_add": ev[2].elapsed_time(ev[3]),
                "total": ev[0].elapsed_time(ev[3])}

    def close_peer(self):
        for P in getattr(self, "_peer", {}).values():
            for b in P["opened"]:
                self.ctx.ipc_close(b)
            for b in P["recv"]:
                self.ctx.free_d(b)
        self._peer = {}

    def export_add_vector(self, vec, dofs):
        """Reference: MultiVector::exportFromVector(repeated, ..., "Add") after FE::assemblyRHS (Problem_def.hpp:213):
        the entries of rows owned elsewhere (behind the owned ones in pattern-row order) are shipped to the owners and
        added; afterwards vec[: dofs * n_owned] is the vector on the unique map."""
        import torch
        import torch.distributed as dist
        if self.size == 1:
            return vec
        key = ("vec", dofs)
        if key not in self._slot_t:
            self._slot_t[key] = torch.from_numpy(self.plan.vec_recv_slots(dofs)).to(self._dev)
            self._recv_buf[key] = torch.empty(self._slot_t[key].numel(), dtype=torch.float64, device=self._dev)
        ssz, rsz = self.plan.vec_split_sizes(dofs)
        send, recv = vec[dofs * self.plan.n_owned:], self._recv_buf[key]
        dist.all_to_all_single(recv, send, rsz, ssz)
        off = 0
        for n in rsz:                       # one launch per sender: deterministic sums
            if n:
                self.ctx.unpack_add_d(vec, recv[off:off + n], self._slot_t[key][off:off + n])
            off += n
        return vec

    def assemble_rhs(self, value_func, deg_func=0, vec_field=False):
        """FE::assemblyRHS on this rank's elements + the export/ADD to the owners; returns the device vector in
        pattern-row order (owned rows first = the unique-map vector)."""
        import torch
        dofs = self.dim if vec_field else 1
        vec = torch.from_numpy(self.pat.assemble_rhs(value_func, deg_func, vec_field)).to(self._dev)
        return self.export_add_vector(vec, dofs)

    def _unpack(self, values, key, rsz):
        recv, slots = self._recv_buf[key], self._slot_t[key]
        off = 0
        for n in rsz:
            if n:
                self.ctx.unpack_add_d(values, recv[off:off + n], slots[off:off + n])
            off += n

    def exchange(self, values, rd, cd, mode):
        """Ship ghost-row values to their owners (NCCL all-to-all-v) and add them into the owned CSR."""
        import torch
        import torch.distributed as dist
        if self.size == 1:
            return
        key = self._exchange_buffers(rd, cd, mode)
        ssz, rsz = self.plan.split_sizes(rd, cd, mode)
        n_owned_vals = self.pat.nnz_owned(rd, cd, mode)
        send = values[n_owned_vals:]
        recv = self._recv_buf[key]
        dist.all_to_all_single(recv, send, rsz, ssz)
        # one unpack launch per sender: slots are distinct within a sender, launches are stream ordered,
        # so the summation order is fixed (deterministic)
        off = 0
        slots = self._slot_t[key]
        for n in rsz:
            if n:
                self.ctx.unpack_add_d(values, recv[off:off + n], slots[off:off + n])
            off += n


def box_dims(world: int):
    """Sub-domain grid for `world` ranks: cubes first (reference: N^3 ranks), else near-cubic boxes."""
    best = (world, 1, 1)
    for nx in range(1, world + 1):
        for ny in range(1, world // nx + 1):
            if world % (nx * ny) == 0:
                nz = world // (nx * ny)
                cand = tuple(sorted((nx, ny, nz), reverse=True))
                if max(cand) - min(cand) < max(best) - min(best):
                    best = cand
    return best


class DistributedElasticity(DistributedMatrixAssembler):
    """bench.py's multi-GPU workload: P2 elasticity on a box of `world` structured sub-cubes (H/h = M each)."""

    def __init__(self, ctx, dim, fe, M, rank, world):
        from . import mesh as PM
        import time
        dims = box_dims(world)
        t0 = time.perf_counter()
        conn, coords, gid, owner = PM.build_structured_box(dim, fe, dims, M, rank)
        t_mesh = time.perf_counter() - t0
        super().__init__(ctx, dim, conn, coords, gid, owner, rank, world)
        self.timing["mesh_generation_s"] = t_mesh
        self.dims = dims

    def exchange(self, values):
        super().exchange(values, self.dim, self.dim, BLOCK_FULL)

    def assemble_linelas_fused(self, values, lam, mu):
        self.assemble_fused(values, self.dim, self.dim, BLOCK_FULL, lambda: self.pat.assemble_linelas_d(values, lam, mu))

    def assemble_linelas_overlapped(self, values, lam, mu):
        self.assemble_overlapped(values, self.dim, self.dim, BLOCK_FULL, lambda: self.pat.assemble_linelas_d(values, lam, mu))
