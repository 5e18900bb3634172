"""Pins oracle/csrops.py block_merge against the reference's own BlockMatrix::merge / BlockMap::merge, compiled where they lie
(oracle/_ref/libfedd_ref_bm.so; skipped where the reference tree was not available at build time), and against committed golden
vectors.  (Matrix::addMatrix is a two-line wrapper around Xpetra's TwoMatrixAdd, core/LinearAlgebra/Matrix_def.hpp:281-287: its
arithmetic lives in Trilinos, which /root/reference does not contain -- add_matrix stays a reviewed restatement.)"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import csrops as OC
from oracle import ref_bm as RB
from util import GOLDEN


def system(seed, with_c=True):
    """Stokes-like 2 x 2 block system [[A, BT], [B, C]]: random sparsity, random values, every block with its own column map
    (a permutation of the block column's gids, the way Tpetra orders owned / remote columns)."""
    rng = np.random.default_rng(seed)
    n = [40 + 3 * seed, 17]
    row_gids = [np.arange(n[0], dtype=np.int64), np.arange(n[1], dtype=np.int64)]
    blocks, blocks_glob = {}, [[None, None], [None, None]]
    for i in range(2):
        for j in range(2):
            if (i, j) == (1, 1) and not with_c:
                continue
            M = sp.random(n[i], n[j], density=0.2, random_state=int(rng.integers(1 << 30)), format="csr")
            M.data[:] = rng.uniform(-1, 1, M.nnz)
            M.sort_indices()
            perm = rng.permutation(n[j])                  # local column id of gid g is perm[g]
            col_gid = np.empty(n[j], dtype=np.int64)
            col_gid[perm] = np.arange(n[j])
            blocks[(i, j)] = (M.indptr.astype(np.int64), perm[M.indices].astype(np.int32), M.data.copy(), col_gid)
            blocks_glob[i][j] = (M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data.copy())
    return n, row_gids, blocks, blocks_glob


@pytest.mark.parametrize("seed,with_c", [(0, True), (1, False), (2, True)])
def test_block_merge_restatement_equals_reference(seed, with_c):
    if not RB.available():
        pytest.skip("oracle/_ref/libfedd_ref_bm.so is not built (needs the reference tree at build time)")
    n, row_gids, blocks, blocks_glob = system(seed, with_c)
    rp, rg, cg, va = RB.merge(row_gids, blocks)
    rp2, ci2, v2 = OC.block_merge(blocks_glob, n, n)
    assert np.array_equal(rg, np.concatenate([row_gids[0], row_gids[1] + n[0]]))       # merged map: offsets = max gid + 1
    assert np.array_equal(rp, rp2) and np.array_equal(cg, ci2) and np.array_equal(va, v2)


def test_golden_vectors():
    z = np.load(os.path.join(GOLDEN, "bm_vectors.npz"))
    n, row_gids, blocks, blocks_glob = system(0, True)
    rp2, ci2, v2 = OC.block_merge(blocks_glob, n, n)
    assert np.array_equal(z["rowptr"], rp2) and np.array_equal(z["colgid"], ci2) and np.array_equal(z["values"], v2)
