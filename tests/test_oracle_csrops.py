"""The numpy restatement of Matrix::addMatrix (TwoMatrixAdd) and BlockMatrix::merge (oracle/csrops.py) against scipy on
random matrices, including explicit zeros, empty rows and absent blocks.  CPU only; the device kernels are checked
against this restatement in tests/test_gpu_csrops.py."""
import numpy as np
import scipy.sparse as sp

from oracle import csrops as OC


def _rand(n, m, density, seed, explicit_zeros=0):
    A = sp.random(n, m, density, random_state=seed, format="csr")
    A.sort_indices()
    rp, ci, v = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.copy()
    v[:explicit_zeros] = 0.0           # stored zeros must survive (Tpetra keeps every inserted entry)
    return (rp, ci, v), sp.csr_matrix((v, ci, rp), shape=(n, m))


def test_add_matrix_is_the_union_with_explicit_zeros():
    (A, As), (B, Bs) = _rand(40, 50, 0.15, 1, explicit_zeros=5), _rand(40, 50, 0.2, 2, explicit_zeros=3)
    rp, ci, v = OC.add_matrix(2.0, A, -1.5, B)
    S = sp.csr_matrix((v, ci, rp), shape=(40, 50))
    assert abs(S - (2.0 * As - 1.5 * Bs)).max() < 1e-15
    union = ((abs(As) + abs(Bs)) != 0).tocsr()
    pa = sp.csr_matrix((np.ones(A[1].size), A[1], A[0]), shape=(40, 50))
    pb = sp.csr_matrix((np.ones(B[1].size), B[1], B[0]), shape=(40, 50))
    pat = (pa + pb).tocsr()
    pat.sort_indices()
    assert np.array_equal(rp, pat.indptr) and np.array_equal(ci, pat.indices)     # pattern union, zeros included
    assert pat.nnz >= union.nnz
    for r in range(40):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)                           # ascending, no duplicates


def test_block_merge_matches_bmat_and_keeps_absent_blocks_empty():
    (A, As), (B, Bs), (Cm, Cs) = _rand(30, 30, 0.2, 3), _rand(12, 30, 0.3, 4), _rand(30, 12, 0.3, 5)
    rp, ci, v = OC.block_merge([[A, Cm], [B, None]], [30, 12], [30, 12])
    S = sp.csr_matrix((v, ci, rp), shape=(42, 42))
    assert abs(S - sp.bmat([[As, Cs], [Bs, None]]).tocsr()).max() == 0.0
    assert S[30:, 30:].nnz == 0
    assert rp[-1] == A[1].size + B[1].size + Cm[1].size
