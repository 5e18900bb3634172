"""Device CSR algebra: Matrix::addMatrix (TwoMatrixAdd) and BlockMatrix::merge on resident
matrices, against the numpy restatement in oracle/csrops.py.  Patterns bit-exact, values bit-exact (one addition per
entry, no re-association)."""
import numpy as np
import pytest

from oracle import csrops as OC
from util import mesh_structured, random_u

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def system(engine_ctx):
    from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, Mesh, Pattern
    from feddlib_b200.engine import DeviceCsr, assemble_div_divT
    dim = 3
    conn, coords = mesh_structured(dim, "P2", 3)
    conn_p, coords_p = mesh_structured(dim, "P1", 3)
    npre = coords_p.shape[0]
    ctx = engine_ctx
    ctx.set_scatter_mode("gather")
    mv, mp = Mesh(ctx, dim, conn, coords), Mesh(ctx, dim, conn_p, coords_p)
    pat = Pattern(ctx, mv)
    patB, patBT = Pattern(ctx, mp, mv), Pattern(ctx, mv, mp)
    u = random_u(dim, coords.shape[0])
    nn = coords.shape[0]
    A = DeviceCsr.from_host(ctx, *pat.expand(dim, dim, BLOCK_DIAG), pat.assemble_laplace(True), dim * nn)
    W = DeviceCsr.from_host(ctx, *pat.expand(dim, dim, BLOCK_FULL), pat.assemble_advection_in_u(u), dim * nn)
    vB, vBT = assemble_div_divT(ctx, patB, patBT)
    B = DeviceCsr.from_host(ctx, *patB.expand(1, dim, BLOCK_FULL), vB, dim * nn)
    BT = DeviceCsr.from_host(ctx, *patBT.expand(dim, 1, BLOCK_FULL), vBT, npre)
    return dict(ctx=ctx, A=A, W=W, B=B, BT=BT, dim=dim, nn=nn, npre=npre, keep=(mv, mp, pat, patB, patBT))


def _host(M):
    rp, ci, v = M.to_host()
    return rp, ci.astype(np.int64), v


def test_add_matrix_union_pattern(system):
    """NavierStokes::reAssemble: A->addMatrix(1., ANW, 0.); W->addMatrix(1., ANW, 1.) (NavierStokes_def.hpp:303-313):
    block-diagonal + full pattern -> the full pattern, values alpha*A + beta*W."""
    from feddlib_b200.engine import csr_add
    s = system
    for alpha, beta in ((1.0, 1.0), (1e-3, 2.5), (1.0, 0.0)):
        C = csr_add(s["ctx"], alpha, s["A"], beta, s["W"])
        rp, ci, v = _host(C)
        rpo, cio, vo = OC.add_matrix(alpha, _host(s["A"]), beta, _host(s["W"]))
        assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
        assert np.array_equal(v, vo)
    # the union is symmetric in its arguments; a matrix added to itself keeps its pattern
    C2 = csr_add(s["ctx"], 2.5, s["W"], 1e-3, s["A"])
    C1 = csr_add(s["ctx"], 1e-3, s["A"], 2.5, s["W"])
    assert np.array_equal(_host(C1)[1], _host(C2)[1]) and np.array_equal(_host(C1)[2], _host(C2)[2])
    D = csr_add(s["ctx"], 1.0, s["A"], 1.0, s["A"])
    assert np.array_equal(_host(D)[0], _host(s["A"])[0]) and np.array_equal(_host(D)[2], 2.0 * _host(s["A"])[2])


def test_add_matrix_disjoint_and_empty_rows(system):
    from feddlib_b200.engine import DeviceCsr, csr_add
    ctx = system["ctx"]
    rpA = np.array([0, 2, 2, 5, 5]); ciA = np.array([0, 4, 1, 3, 5]); vA = np.arange(1.0, 6.0)
    rpB = np.array([0, 1, 1, 4, 6]); ciB = np.array([2, 0, 3, 4, 0, 5]); vB = -np.arange(1.0, 7.0)
    A, B = DeviceCsr.from_host(ctx, rpA, ciA, vA, 6), DeviceCsr.from_host(ctx, rpB, ciB, vB, 6)
    rp, ci, v = _host(csr_add(ctx, 2.0, A, 3.0, B))
    rpo, cio, vo = OC.add_matrix(2.0, (rpA, ciA, vA), 3.0, (rpB, ciB, vB))
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio) and np.array_equal(v, vo)
    assert rp.tolist() == [0, 3, 3, 8, 10]


def test_block_merge_saddle_point(system):
    """[F B^T; B 0] of the Stokes / Navier-Stokes drivers merged into one matrix (BlockMatrix_def.hpp:119-289)."""
    from feddlib_b200.engine import block_merge, csr_add
    s = system
    F = csr_add(s["ctx"], 1.0, s["A"], 1.0, s["W"])
    M = block_merge(s["ctx"], [[F, s["BT"]], [s["B"], None]])
    rp, ci, v = _host(M)
    nv, npre = s["dim"] * s["nn"], s["npre"]
    rpo, cio, vo = OC.block_merge([[_host(F), _host(s["BT"])], [_host(s["B"]), None]], [nv, npre], [nv, npre])
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio) and np.array_equal(v, vo)
    assert M.n_rows == nv + npre and M.n_cols == nv + npre
    # transpose structure of the merged saddle-point matrix: the (1,0) block is the transpose of the (0,1) block
    import scipy.sparse as sp
    S = sp.csr_matrix((v, ci, rp), shape=(nv + npre, nv + npre))
    assert abs(S[nv:, :nv] - S[:nv, nv:].T).max() < 1e-12 * abs(S).max()
    assert S[nv:, nv:].nnz == 0


def test_block_merge_errors(system):
    from feddlib_b200 import LogicError
    from feddlib_b200.engine import block_merge
    s = system
    with pytest.raises(LogicError):
        block_merge(s["ctx"], [[s["A"], None], [None, None]])          # block row without a block
    with pytest.raises(LogicError):
        block_merge(s["ctx"], [[s["A"], s["B"]], [s["B"], None]])      # inconsistent sizes
