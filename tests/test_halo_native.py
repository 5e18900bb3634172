"""The C++ multi-rank host plan (include/feddb200_halo.h) against the numpy HaloPlan (feddlib_b200/dist.py, the executable
specification that tests/test_dist_cpu.py checks against the oracle's global matrix): identical index spaces, column maps,
exchange plans and value slots on structured boxes AND on an irregular (random) element partition; host-side
globalAssemble (export_add) and the unique -> repeated vector import reproduce the serial result.  All ranks run as
threads of this process over an in-process communicator -- the same callbacks an MPI build would supply."""
import threading

import numpy as np
import pytest

from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR
from feddlib_b200 import mesh as PM
from feddlib_b200.dist import HaloPlan, box_dims, numpy_node_pattern
from feddlib_b200.halo import NativeHaloPlan


class ThreadComm:
    """alltoallv between the threads of one process (one thread per rank)."""

    class _Shared:
        def __init__(self, size):
            self.size, self.mail, self.barrier = size, [[None] * size for _ in range(size)], threading.Barrier(size)

    def __init__(self, shared, rank):
        self.s, self.rank, self.size = shared, rank, shared.size

    def alltoallv(self, chunks, np_dtype=np.int64):
        for d in range(self.size):
            self.s.mail[self.rank][d] = np.array(chunks[d], dtype=np_dtype, copy=True)
        self.s.barrier.wait()
        got = [self.s.mail[src][self.rank] for src in range(self.size)]
        self.s.barrier.wait()
        return got


def run_ranks(size, fn):
    out, err = [None] * size, [None] * size

    def work(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            try:
                shared_abort[0].barrier.abort()
            except Exception:  # noqa: BLE001
                pass

    ts = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    for e in err:
        if e is not None:
            raise e
    return out


shared_abort = [None]


def partitions(kind, size):
    """Per-rank (conn, gid, owner) of one global mesh."""
    if kind == "box":
        dims = box_dims(size)
        parts = [PM.build_structured_box(3, "P2", dims, 2, r) for r in range(size)]
        return [(c, g, o) for c, _, g, o in parts]
    # irregular: the elements of one P1 cube dealt to the ranks at random (what a METIS epart file amounts to:
    # MeshPartitioner_def.hpp:349-355); a node is owned by the lowest rank holding it (Map_def.hpp:194-199 leaves the
    # choice to Tpetra's directory)
    conn, coords, _ = PM.build_structured(3, "P1" if kind == "random-P1" else "P2", 1, 3)
    rng = np.random.default_rng(42)
    epart = rng.integers(0, size, conn.shape[0])
    holder = np.full(coords.shape[0], size, dtype=np.int64)
    for r in range(size):
        np.minimum.at(holder, np.unique(conn[epart == r]), r)
    out = []
    for r in range(size):
        ce = conn[epart == r]
        gids = np.unique(ce)
        loc = -np.ones(coords.shape[0], dtype=np.int64)
        loc[gids] = np.arange(gids.size)
        out.append((loc[ce].astype(np.int32), gids.astype(np.int64), holder[gids].astype(np.int32)))
    return out


@pytest.mark.parametrize("kind,size", [("box", 2), ("box", 4), ("random-P1", 3), ("random-P2", 4)])
def test_native_plan_equals_numpy_plan(kind, size):
    parts = partitions(kind, size)
    results = {}
    for impl in ("numpy", "native"):
        shared = ThreadComm._Shared(size)
        shared_abort[0] = shared

        def build(r, impl=impl, shared=shared):
            conn, gid, owner = parts[r]

            def pattern_fn(row_lid, n_rows, n_owned, col_lid, n_cols, er, ec):
                return numpy_node_pattern(conn, conn, row_lid, n_rows, col_lid, er, ec)

            comm = ThreadComm(shared, r)
            return (HaloPlan if impl == "numpy" else NativeHaloPlan)(comm, gid, owner, pattern_fn)

        results[impl] = run_ranks(size, build)
    for r in range(size):
        a, b = results["numpy"][r], results["native"][r]
        for name in ("n_owned", "n_ghost", "n_rows", "n_colmap", "n_cols", "nnz_owned_nodes", "nnz_nodes"):
            assert getattr(a, name) == getattr(b, name), (r, name)
        for name in ("row_lid", "col_lid", "extra_row", "extra_col", "colmap_gids", "unique_gids", "ghost_row_gids", "ghost_row_owner",
                     "rowptr", "colind", "send_counts_nodes", "recv_counts_nodes", "recv_row", "recv_pos", "recv_len_sender", "recv_q"):
            assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))), (r, name)
        for layout in ((3, 3, BLOCK_FULL), (3, 3, BLOCK_DIAG), (1, 1, BLOCK_SCALAR), (1, 3, BLOCK_FULL)):
            assert np.array_equal(a.recv_slots(*layout), b.recv_slots(*layout)), (r, layout)
            assert a.split_sizes(*layout) == b.split_sizes(*layout)


@pytest.mark.parametrize("kind,size", [("box", 2), ("random-P2", 3)])
def test_export_add_and_import_vector(kind, size):
    """export_add = the numpy scatter of the received ghost values; import_vector gives every repeated node its owner's value."""
    parts = partitions(kind, size)
    shared = ThreadComm._Shared(size)
    shared_abort[0] = shared
    rd = cd = 3

    def work(r):
        conn, gid, owner = parts[r]

        def pattern_fn(row_lid, n_rows, n_owned, col_lid, n_cols, er, ec):
            return numpy_node_pattern(conn, conn, row_lid, n_rows, col_lid, er, ec)

        comm = ThreadComm(shared, r)
        plan = NativeHaloPlan(comm, gid, owner, pattern_fn)
        rng = np.random.default_rng(100 + r)
        values = rng.standard_normal(rd * cd * plan.nnz_nodes)
        n_owned_vals = rd * cd * plan.nnz_owned_nodes
        # reference: ship the ghost part with the communicator, scatter-add with the plan's slots
        ssz, rsz = plan.split_sizes(rd, cd, BLOCK_FULL)
        off = np.concatenate([[0], np.cumsum(ssz)])
        ghost = values[n_owned_vals:]
        got = comm.alltoallv([ghost[off[d]:off[d + 1]].view(np.int64) for d in range(size)])
        want = values.copy()
        np.add.at(want, plan.recv_slots(rd, cd, BLOCK_FULL), np.concatenate(got).view(np.float64))
        have = plan.export_add(values.copy(), rd, cd, BLOCK_FULL, n_owned_vals)
        assert np.array_equal(have[:n_owned_vals], want[:n_owned_vals])
        # import: u(gid) = (gid, 2 gid + 1, -gid) on the unique map -> the same function of gid on the repeated map
        ug = plan.unique_gids.astype(np.float64)
        u_unique = np.stack([ug, 2 * ug + 1, -ug], axis=1).ravel()
        u_rep = plan.import_vector(u_unique, 3)
        g = gid.astype(np.float64)
        assert np.array_equal(u_rep, np.stack([g, 2 * g + 1, -g], axis=1).ravel())
        return True

    assert all(run_ranks(size, work))


def test_halo_header_symbols_are_exported():
    import ctypes
    import os
    import re
    from feddlib_b200 import _lib, halo
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "feddb200_halo.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(feddb200_halo_[a-z0-9_]+)\s*\(", src)))
    L = ctypes.CDLL(_lib.LIB_PATH)
    assert syms and all(hasattr(L, s) for s in syms)
    assert sorted(halo.HALO_SIGNATURES) == syms
