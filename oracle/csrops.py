"""Test infrastructure: numpy/scipy restatement of the two matrix passes SURVEY.md 8(f) rank 3 names.  Their arithmetic
lives in Trilinos (Tpetra insertGlobalValues / fillComplete, Xpetra TwoMatrixAdd), which is absent from /root/reference.
block_merge is PINNED: tests/test_csrops_vs_ref.py runs the reference's own BlockMatrix::merge / BlockMap::merge (compiled where
they lie against mock containers, oracle/ref_bm.py) and compares bitwise, golden vectors in tests/golden/bm_vectors.npz.
add_matrix is a reviewed restatement ("parity unpinned"): the reference routine is a two-line wrapper around Xpetra's TwoMatrixAdd.

  add_matrix   Matrix::addMatrix(alpha, B, beta): B := alpha*A + beta*B via TwoMatrixAdd on a resumed-fill B
               (core/LinearAlgebra/Matrix_def.hpp:281-287): every entry of A is summed into B; entries new to B are
               inserted, so the result lives on the UNION of the two patterns, explicit zeros included.
  block_merge  BlockMatrix::merge -> mergeBlockNew (core/LinearAlgebra/BlockMatrix_def.hpp:119-147, 252-270): every row
               of block (i, j) is inserted at row + rowOffset_i with columns + colOffset_j; offsets are the running sums
               of the block sizes (determineGlobalOffsets, :211-249); fillComplete sorts the columns of each row.
"""
import numpy as np


def _rows(rowptr):
    return np.repeat(np.arange(rowptr.size - 1, dtype=np.int64), np.diff(rowptr))


def _assemble(n_rows, rows, cols, vals):
    """insert + fillComplete: sort by (row, col), sum duplicates in insertion order, keep explicit zeros."""
    order = np.lexsort((cols, rows))            # stable: equal (row, col) keep their insertion order
    rows, cols, vals = rows[order], cols[order], vals[order]
    first = np.ones(rows.size, dtype=bool)
    first[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    idx = np.cumsum(first) - 1
    out = np.zeros(int(first.sum()))
    for k in range(rows.size):                  # sequential sum, like the CPU path
        out[idx[k]] += vals[k]
    r, c = rows[first], cols[first]
    rowptr = np.searchsorted(r, np.arange(n_rows + 1)).astype(np.int64)
    return rowptr, c.astype(np.int64), out


def add_matrix(alpha, A, beta, B):
    """A, B = (rowptr, colind, values).  Returns alpha*A + beta*B on the union pattern (B's entries first, then A's)."""
    n = A[0].size - 1
    rows = np.concatenate([_rows(B[0]), _rows(A[0])])
    cols = np.concatenate([B[1], A[1]]).astype(np.int64)
    vals = np.concatenate([beta * B[2], alpha * A[2]])
    return _assemble(n, rows, cols, vals)


def block_merge(blocks, n_rows, n_cols):
    """blocks[i][j] = (rowptr, colind, values) or None."""
    nb = len(blocks)
    roff = np.concatenate([[0], np.cumsum(n_rows)]).astype(np.int64)
    coff = np.concatenate([[0], np.cumsum(n_cols)]).astype(np.int64)
    R, Cc, V = [], [], []
    for i in range(nb):
        for j in range(nb):
            if blocks[i][j] is None:
                continue
            rp, ci, v = blocks[i][j]
            R.append(_rows(rp) + roff[i]); Cc.append(ci.astype(np.int64) + coff[j]); V.append(v)
    return _assemble(int(roff[-1]), np.concatenate(R), np.concatenate(Cc), np.concatenate(V))
