"""Host-side multi-rank logic on CPU: ownership, column map, ghost-row plan and the value exchange,
run as a real 2-rank (and 4-rank) torch.distributed job over gloo.  Per-rank element matrices come from
the oracle; the distributed result must equal the oracle's global matrix."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import torch, torch.distributed as dist
from feddlib_b200 import mesh as PM
from feddlib_b200.dist import Comm, HaloPlan, numpy_node_pattern, box_dims
from feddlib_b200 import BLOCK_FULL, BLOCK_SCALAR, BLOCK_DIAG
from oracle import oracle as O
import scipy.sparse as sp

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dim, fe, M = int(os.environ["T_DIM"]), os.environ["T_FE"], int(os.environ["T_M"])
dims = box_dims(world)
conn, coords, gid, owner = PM.build_structured_box(dim, fe, dims, M, rank)
nn = coords.shape[0]

def pattern_fn(row_lid, n_rows, n_owned, col_lid, n_cols, er, ec):
    return numpy_node_pattern(conn, conn, row_lid, n_rows, col_lid, er, ec)

plan = HaloPlan(Comm(rank, world), gid, owner, pattern_fn)

# --- per-rank local values from the oracle, laid out in the plan's CSR (owned + ghost rows) ---
lam, mu = 8e6, 2e6
A = O.Matrix(dim * nn, 64)
O.assembly_linelas(dim, fe, conn, coords, np.arange(nn), lam, mu, A)      # rows/cols = repeated local dofs
S = A.scipy().tocoo()
node_r, a = S.row // dim, S.row % dim
node_c, b = S.col // dim, S.col % dim
I = plan.row_lid[node_r].astype(np.int64); J = plan.col_lid[node_c].astype(np.int64)
rp, ci = plan.rowptr, plan.colind
key_all = (np.repeat(np.arange(plan.n_rows), np.diff(rp)).astype(np.int64) << 32) | ci.astype(np.int64)
slot = np.searchsorted(key_all, (I << 32) | J)
assert np.array_equal(key_all[slot], (I << 32) | J)
p = slot - rp[I]; L = rp[I + 1] - rp[I]
values = np.zeros(dim * dim * plan.nnz_nodes)
values[dim * dim * rp[I] + a * dim * L + dim * p + b] = S.data

# --- exchange (what DistributedMatrixAssembler.exchange does on the device) ---
ssz, rsz = plan.split_sizes(dim, dim, BLOCK_FULL)
n_owned_vals = dim * dim * plan.nnz_owned_nodes
send = torch.from_numpy(values[n_owned_vals:].copy())
recv = torch.empty(int(sum(rsz)), dtype=torch.float64)
dist.all_to_all_single(recv, send, rsz, ssz)
slots = plan.recv_slots(dim, dim, BLOCK_FULL)
assert slots.size == recv.numel() and (slots < n_owned_vals).all()
np.add.at(values, slots, recv.numpy())

# --- compare the owned rows with the oracle's global matrix ---
nglob = 1
for d in range(dim):
    k = 2 if fe == "P2" else 1
    nglob *= dims[d] * k * M + 1
G = O.Matrix(dim * nglob, 64)
for r in range(world):
    c2, x2, g2, _ = PM.build_structured_box(dim, fe, dims, M, r)
    O.assembly_linelas(dim, fe, c2, x2, g2, lam, mu, G)
Gs = G.scipy().tocsr()
# dof-level CSR of the owned rows in the plan's index space
rows_gid = plan.unique_gids
cols_gid = plan.colmap_gids
err_num = 0.0; ref_num = 0.0; nnz_checked = 0
for Irow in range(plan.n_owned):
    for aa in range(dim):
        grow = dim * rows_gid[Irow] + aa
        seg = Gs.indices[Gs.indptr[grow]:Gs.indptr[grow + 1]]
        segv = Gs.data[Gs.indptr[grow]:Gs.indptr[grow + 1]]
        Lr = rp[Irow + 1] - rp[Irow]
        mine_cols = (dim * cols_gid[ci[rp[Irow]:rp[Irow + 1]]][:, None] + np.arange(dim)[None, :]).ravel()
        mine_vals = values[dim * dim * rp[Irow] + aa * dim * Lr: dim * dim * rp[Irow] + (aa + 1) * dim * Lr]
        order = np.argsort(mine_cols)
        assert np.array_equal(mine_cols[order], seg), (rank, Irow, aa)       # same pattern as the global matrix
        err_num += ((mine_vals[order] - segv) ** 2).sum(); ref_num += (segv ** 2).sum(); nnz_checked += seg.size
assert np.sqrt(err_num / ref_num) < 1e-13, np.sqrt(err_num / ref_num)
# column map rule: owned gids first (unique-map order), then remotes grouped by owner, ascending gid
assert np.array_equal(plan.colmap_gids[: plan.n_owned], plan.unique_gids)
rem = plan.colmap_gids[plan.n_owned:]
if rem.size:
    all_owner = np.zeros(nglob, dtype=np.int64)
    for r in range(world):
        c2, x2, g2, o2 = PM.build_structured_box(dim, fe, dims, M, r)
        all_owner[g2] = o2
    ro = all_owner[rem]
    assert np.all(ro != rank)
    assert np.all((np.diff(ro) > 0) | ((np.diff(ro) == 0) & (np.diff(rem) > 0)))
# every node is owned exactly once
cnt = torch.zeros(nglob, dtype=torch.int64); cnt[torch.from_numpy(plan.unique_gids)] = 1
dist.all_reduce(cnt)
assert int(cnt.min()) == 1 and int(cnt.max()) == 1
tot = torch.tensor([nnz_checked]); dist.all_reduce(tot)
assert int(tot) == Gs.nnz
# --- load vector: FE::assemblyRHS on the repeated map + exportFromVector(..., "Add") (Problem_def.hpp:184-216) ---
for vec_field in (False, True):
    dofs = dim if vec_field else 1
    f = np.array([1.5, -2.0, 0.25])[:dim]
    loc = O.assembly_rhs(dim, fe, conn, coords, f, 1, vec_field)                 # repeated-node order
    vec = np.zeros(dofs * plan.n_rows)
    rows = plan.row_lid.astype(np.int64)
    vec[(dofs * rows[:, None] + np.arange(dofs)[None, :]).ravel()] = loc
    ssz, rsz = plan.vec_split_sizes(dofs)
    send = torch.from_numpy(vec[dofs * plan.n_owned:].copy())
    recv = torch.empty(int(sum(rsz)), dtype=torch.float64)
    dist.all_to_all_single(recv, send, rsz, ssz)
    vs = plan.vec_recv_slots(dofs)
    assert vs.size == recv.numel() and (vs < dofs * plan.n_owned).all()
    np.add.at(vec, vs, recv.numpy())
    glob = np.zeros(dofs * nglob)
    for r in range(world):
        c2, x2, g2, _ = PM.build_structured_box(dim, fe, dims, M, r)
        part = O.assembly_rhs(dim, fe, c2, x2, f, 1, vec_field)
        np.add.at(glob, (dofs * g2[:, None] + np.arange(dofs)[None, :]).ravel(), part)
    want = glob[(dofs * plan.unique_gids[:, None] + np.arange(dofs)[None, :]).ravel()]
    assert np.abs(vec[: dofs * plan.n_owned] - want).max() <= 1e-13 * np.abs(want).max(), "load vector export/ADD"
print(f"rank {rank}/{world} OK: owned {plan.n_owned} ghost {plan.n_ghost} extra {plan.extra_row.size} recv {slots.size}")
dist.destroy_process_group()
'''


def run_world(world, dim, fe, M, port):
    env = dict(os.environ, REPO_ROOT=ROOT, T_DIM=str(dim), T_FE=fe, T_M=str(M), OMP_NUM_THREADS="1")
    path = os.path.join(ROOT, "tests", "_dist_worker.py")
    with open(path, "w") as f:
        f.write(WORKER)
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                              "--master-addr", "127.0.0.1", "--master-port", str(port), path],
                             env=env, capture_output=True, text=True, timeout=600)
    finally:
        os.remove(path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" OK: ") == world, out.stdout[-2000:]


def test_ghost_exchange_world2_p2_tets():
    run_world(2, 3, "P2", 2, 29541)


def test_ghost_exchange_world4_p1_triangles():
    run_world(4, 2, "P1", 3, 29542)


def test_ghost_exchange_world8_cube_p1_tets():
    run_world(8, 3, "P1", 2, 29543)


def test_box_mesh_matches_reference_cube_decomposition():
    """For N^3 ranks the box generator reproduces the reference's sub-cube meshes (up to the domain scaling)."""
    from feddlib_b200 import mesh as PM
    from oracle import mesh as OM
    for rank in range(8):
        c, x, g, o = PM.build_structured_box(3, "P2", (2, 2, 2), 2, rank)
        c0, x0, g0 = OM.structured(3, "P2", 2, 2, rank)
        assert np.array_equal(c, c0) and np.array_equal(g, g0)
        assert np.allclose(x, 2.0 * x0)
    owners = OM.lowest_rank_owner([PM.build_structured_box(3, "P2", (2, 2, 2), 2, r)[2] for r in range(8)])
    for rank in range(8):
        assert np.array_equal(PM.build_structured_box(3, "P2", (2, 2, 2), 2, rank)[3], owners[rank])


def test_box_dims():
    from feddlib_b200.dist import box_dims
    assert box_dims(1) == (1, 1, 1) and box_dims(2) == (2, 1, 1) and box_dims(4) == (2, 2, 1) and box_dims(8) == (2, 2, 2)
