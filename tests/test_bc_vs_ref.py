"""Pins the Dirichlet boundary-condition restatements (oracle.set_dirichlet_rows / dirichlet_dof_masks / set_dirichlet_rhs)
against the reference's own BCBuilder members, compiled where they lie (oracle/_ref/libfedd_ref_bc.so; skipped where the
reference tree was not available at build time), and against committed golden vectors of the same routines."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import ref_bc as RB
from util import GOLDEN, mesh_structured

BCS = [(1, 0, "Dirichlet", 3), (2, 0, "Dirichlet_X_Z", 3), (3, 0, "Dirichlet_Y", 3), (4, 0, "Robin", 3), (2, 1, "Dirichlet", 1),
       (1, 0, "Dirichlet_Z", 3)]   # "Robin": a non-Dirichlet entry (ignored by both routines); the last entry is shadowed by the first one with flag 1 (findFlag returns the first match)


def system(seed=0):
    """2 x 2 block system on the nodes of a small P1 cube: velocity (3 dofs) and pressure (1 dof) blocks with random values,
    random node gids, boundary flags 0..4 by position."""
    conn, coords = mesh_structured(3, "P1", 3, warp=True)
    nn = coords.shape[0]
    rng = np.random.default_rng(seed)
    gid = rng.permutation(nn).astype(np.int64) + 5
    flags = np.zeros(nn, dtype=np.int32)
    flags[np.abs(coords[:, 0]) < 1e-12] = 1
    flags[np.abs(coords[:, 1] - 1) < 1e-12] = 2
    flags[np.abs(coords[:, 2]) < 1e-12] = 3
    flags[np.abs(coords[:, 2] - 1) < 1e-12] = 4
    P = sp.csr_matrix((np.ones(conn.size * 4), (np.repeat(conn, 4, axis=1).ravel(), np.tile(conn, (1, 4)).ravel())), shape=(nn, nn))
    P.sum_duplicates()
    dofs = [3, 1]
    blocks = {}
    for i in range(2):
        for j in range(2):
            if (i, j) == (1, 1) and seed % 2:      # block absent (Stokes without stabilisation)
                continue
            B = sp.kron(P, np.ones((dofs[i], dofs[j]))).tocsr()
            B.sort_indices()
            perm = rng.permutation(nn * dofs[j])                     # column map in an arbitrary local order
            col_gid = np.empty(nn * dofs[j], dtype=np.int64)
            col_gid[perm] = (dofs[j] * gid[:, None] + np.arange(dofs[j])[None, :]).ravel()
            colind = perm[B.indices].astype(np.int32)
            blocks[(i, j)] = (B.indptr.astype(np.int64), colind, rng.uniform(-1, 1, B.nnz), col_gid)
    return coords, gid, flags, dofs, blocks


def restated(gid, flags, dofs, blocks):
    out = {}
    for (i, j), (rp, ci, va, cg) in blocks.items():
        mask, _ = O.dirichlet_dof_masks(flags, BCS, i, dofs[i])
        dofmask = np.zeros(flags.size * dofs[i], dtype=bool)
        for a in range(dofs[i]):
            dofmask[a::dofs[i]] = (mask >> a) & 1
        row_gid = (dofs[i] * gid[:, None] + np.arange(dofs[i])[None, :]).ravel()
        out[(i, j)] = O.set_dirichlet_rows(rp, cg[ci], va, row_gid, dofmask, diagonal_block=(i == j))
    return out


def bc_func(x, t, par):
    return [par[0] * x[0] + t, par[1] - x[1] * x[2], 2.0 + x[2]]


def bc_func1(x, t, par):
    return [par[0] * x[0] + t - x[2]]


def rhs_cases():
    """setRHS block by block: the boundary function writes as many components as the block has dofs (BCBuilder_def.hpp:108-134)."""
    return [(0, [b for b in BCS if b[1] == 0], bc_func), (1, [b for b in BCS if b[1] == 1], bc_func1)]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_set_system_restatement_equals_reference(seed):
    if not RB.available():
        pytest.skip("oracle/_ref/libfedd_ref_bc.so is not built (needs the reference tree at build time)")
    coords, gid, flags, dofs, blocks = system(seed)
    ref = RB.set_system(3, flags, gid, BCS, dofs, blocks)
    mine = restated(gid, flags, dofs, blocks)
    for k in blocks:
        assert np.array_equal(ref[k], mine[k]), k
        assert not np.array_equal(ref[k], blocks[k][2])


@pytest.mark.parametrize("seed", [0, 1])
def test_set_rhs_restatement_equals_reference(seed):
    if not RB.available():
        pytest.skip("oracle/_ref/libfedd_ref_bc.so is not built (needs the reference tree at build time)")
    coords, gid, flags, dofs, _ = system(seed)
    rng = np.random.default_rng(seed + 10)
    rhs = [rng.uniform(-1, 1, flags.size * d) for d in dofs]
    par = np.array([1.5, -0.25])
    for b, bcs, f in rhs_cases():
        ref = RB.set_rhs(3, flags, coords, gid, bcs, f, par, dofs, rhs, t=0.75)
        mine = O.set_dirichlet_rhs(rhs[b], flags, coords, bcs, b, dofs[b], f, par, t=0.75)
        assert np.array_equal(ref[b], mine), b
        assert not np.array_equal(ref[b], rhs[b])
        assert np.array_equal(ref[1 - b], rhs[1 - b])


def test_golden_vectors():
    """Outputs of the reference's own BCBuilder on system(0) (tests/golden/make_bc_vectors.py), checked where neither the
    reference tree nor oracle/_ref exists."""
    z = np.load(os.path.join(GOLDEN, "bc_vectors.npz"))
    coords, gid, flags, dofs, blocks = system(0)
    mine = restated(gid, flags, dofs, blocks)
    for (i, j) in blocks:
        assert np.array_equal(z[f"sys_{i}{j}"], mine[(i, j)])
    rng = np.random.default_rng(10)
    rhs = [rng.uniform(-1, 1, flags.size * d) for d in dofs]
    for b, bcs, f in rhs_cases():
        assert np.array_equal(z[f"rhs_{b}"], O.set_dirichlet_rhs(rhs[b], flags, coords, bcs, b, dofs[b], f, np.array([1.5, -0.25]), t=0.75))
