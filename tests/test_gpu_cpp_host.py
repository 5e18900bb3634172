"""C++ host side of the boundary (feddlib_b200/csrc/host/FE_b200.hpp) against FEDDLib's own FE routines.

tests/cpp/fe_b200_driver.cpp builds ONE mock Domain and runs it through (1) the reference's FE_def.hpp routines
(oracle/_ref/libfedd_ref.so, compiled from /root/reference where it lies) and (2) FEDD::FE_b200 -> C ABI -> CUDA;
patterns must be identical, values within 1e-12 relative Frobenius error (the tolerance of BASELINE.json's
north_star).  The driver also checks the std::logic_error behaviour of the boundary."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import util as U

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "cpp", "_build", "fe_b200_driver")
REF_SO = os.path.join(HERE, "..", "oracle", "_ref", "libfedd_ref.so")

pytestmark = pytest.mark.gpu


def _write_mesh(path, dim, conn1, coords, gid1, conn2, gid2, u):
    with open(path, "wb") as f:
        np.array([dim, conn1.shape[1], conn1.shape[0], coords.shape[0], conn2.shape[1], gid2.size], dtype=np.int64).tofile(f)
        np.ascontiguousarray(conn1, dtype=np.int32).tofile(f)
        np.ascontiguousarray(coords, dtype=np.float64).tofile(f)
        np.ascontiguousarray(gid1, dtype=np.int64).tofile(f)
        np.ascontiguousarray(conn2, dtype=np.int32).tofile(f)
        np.ascontiguousarray(gid2, dtype=np.int64).tofile(f)
        np.ascontiguousarray(u, dtype=np.float64).tofile(f)


def _p1_of_p2(dim, conn):
    """Pressure space of a Taylor-Hood pair: the vertex nodes of the P2 mesh, renumbered densely."""
    nv = dim + 1
    verts = np.unique(conn[:, :nv])
    lid = -np.ones(int(conn.max()) + 1, dtype=np.int64)
    lid[verts] = np.arange(verts.size)
    return lid[conn[:, :nv]].astype(np.int32), verts.size


@pytest.mark.parametrize("dim,fe,M,shuffle,permute_gids", [
    (2, "P2", 6, False, False), (2, "P1", 7, True, False),
    (3, "P2", 3, False, False), (3, "P2", 3, True, True), (3, "P1", 4, True, False),
])
def test_cpp_host_matches_reference(tmp_path, dim, fe, M, shuffle, permute_gids):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libfedd_ref.so is not built (needs the reference tree at build time)")
    assert os.path.exists(DRIVER), "tests/cpp/_build/fe_b200_driver missing: run __graft_entry__.build()"
    conn, coords = U.mesh_structured(dim, fe, M, warp=True, shuffle=shuffle, seed=3)
    nn = coords.shape[0]
    rng = np.random.default_rng(11)
    gid1 = rng.permutation(nn).astype(np.int64) if permute_gids else np.arange(nn, dtype=np.int64)
    if fe == "P2":
        conn2, nn2 = _p1_of_p2(dim, conn)
    else:
        conn2, nn2 = conn, nn
    gid2 = rng.permutation(nn2).astype(np.int64) if permute_gids and fe == "P2" else np.arange(nn2, dtype=np.int64)
    u = U.random_u(dim, nn)
    path = os.path.join(tmp_path, "mesh.bin")
    _write_mesh(path, dim, conn, coords, gid1, conn2, gid2, u)
    r = subprocess.run([DRIVER, path], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


MR_DRIVER = os.path.join(HERE, "cpp", "_build", "fe_b200_mr_driver")


def _partitions(kind, size):
    """Per-rank (conn, coords, gid, owner) of one global P2 tetrahedral mesh."""
    from feddlib_b200 import mesh as PM
    from feddlib_b200.dist import box_dims
    if kind == "box":
        return [PM.build_structured_box(3, "P2", box_dims(size), 2, r) for r in range(size)]
    conn, coords = U.mesh_structured(3, "P2" if kind == "random-P2" else "P1", 3, warp=True, shuffle=False, seed=3)
    epart = np.random.default_rng(42).integers(0, size, conn.shape[0])    # what a METIS epart file amounts to
    holder = np.full(coords.shape[0], size, dtype=np.int64)
    for r in range(size):
        np.minimum.at(holder, np.unique(conn[epart == r]), r)             # lowest rank holding a node owns it
    out = []
    for r in range(size):
        ce = conn[epart == r]
        gids = np.unique(ce)
        loc = -np.ones(coords.shape[0], dtype=np.int64)
        loc[gids] = np.arange(gids.size)
        out.append((loc[ce].astype(np.int32), coords[gids], gids.astype(np.int64), holder[gids].astype(np.int32)))
    return out


@pytest.mark.parametrize("kind,size", [("box", 2), ("random-P2", 3), ("random-P1", 4)])
def test_cpp_host_multi_rank_matches_reference(tmp_path, kind, size):
    """FE_b200 on several ranks (threads, in-process communicator callbacks) against the reference's multi-rank insertion into one
    global matrix: structured boxes and an irregular (random) element partition."""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libfedd_ref.so is not built (needs the reference tree at build time)")
    assert os.path.exists(MR_DRIVER), "tests/cpp/_build/fe_b200_mr_driver missing: run __graft_entry__.build()"
    parts = _partitions(kind, size)
    path = os.path.join(tmp_path, "parts.bin")
    with open(path, "wb") as f:
        np.array([3, parts[0][0].shape[1], size], dtype=np.int64).tofile(f)
        for conn, coords, gid, owner in parts:
            np.array([conn.shape[0], coords.shape[0]], dtype=np.int64).tofile(f)
            np.ascontiguousarray(conn, dtype=np.int32).tofile(f)
            np.ascontiguousarray(coords, dtype=np.float64).tofile(f)
            np.ascontiguousarray(gid, dtype=np.int64).tofile(f)
            np.ascontiguousarray(owner, dtype=np.int32).tofile(f)
            x = coords
            u = np.stack([np.sin(2 * x[:, 1]) + x[:, 2], -x[:, 0] ** 2, 0.3 + x[:, 0] * x[:, 1]], axis=1)   # a field, the same on every rank
            np.ascontiguousarray(u, dtype=np.float64).tofile(f)
    r = subprocess.run([MR_DRIVER, path], capture_output=True, text=True, timeout=900)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr
