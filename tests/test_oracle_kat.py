"""Known-answer tests that pin the CPU oracle.  The reference ships no golden
vectors for this path, so these analytic properties -- plus oracle/_ref where it can be built -- are
what the restatement is anchored on."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mesh as OM
from oracle import oracle as O

from util import mesh_dfg, oracle_csr, random_u


def to_sp(csr, ncols=None):
    rp, c, v = csr
    n = rp.size - 1
    return sp.csr_matrix((v, c, rp), shape=(n, ncols or n))


def test_reference_triangle_laplace():
    conn = np.array([[0, 1, 2]], dtype=np.int32)
    co = np.array([[0, 0], [1, 0], [0, 1.0]])
    K = to_sp(oracle_csr("laplace", 2, "P1", conn, co)).toarray()
    np.testing.assert_allclose(K, 0.5 * np.array([[2, -1, -1], [-1, 1, 0], [-1, 0, 1.0]]), atol=1e-15)


@pytest.mark.parametrize("M", [1, 2, 3, 4])
def test_nnz_polynomials_cube(M):
    for fe, poly in (("P1", 15 * M**3 + 21 * M**2 + 9 * M + 1), ("P2", 230 * M**3 + 138 * M**2 + 24 * M + 1)):
        conn, co, _ = OM.structured(3, fe, 1, M)
        rp, c, v = oracle_csr("laplace", 3, fe, conn, co)
        assert rp[-1] == poly
        assert np.all(np.diff(c)[np.setdiff1d(np.arange(c.size - 1), rp[1:-1] - 1)] > 0)  # ascending per row
    conn2, co2, _ = OM.structured(3, "P2", 1, M)
    c1 = OM.p1_of_p2(conn2, 3)
    rp, c, v = oracle_csr("div", 3, "P2", conn2, co2, fe2="P1", conn2=c1)
    assert rp[-1] == 3 * (65 * M**3 + 57 * M**2 + 15 * M + 1)


def test_nnz_square_cfg1():
    M = 224
    conn, co, _ = OM.structured(2, "P1", 1, M)
    rp, c, v = oracle_csr("laplace", 2, "P1", conn, co)
    assert rp[-1] == (M + 1) ** 2 + 2 * (3 * M * M + 2 * M) == 352577
    assert conn.shape[0] == 100352 and co.shape[0] == 50625


@pytest.mark.parametrize("dim,fe,maxlen", [(2, "P1", 7), (2, "P2", 19), (3, "P1", 15), (3, "P2", 65)])
def test_max_row_length(dim, fe, maxlen):
    conn, co, _ = OM.structured(dim, fe, 1, 4)
    rp, _, _ = oracle_csr("laplace", dim, fe, conn, co)
    assert np.diff(rp).max() == maxlen


@pytest.mark.parametrize("dim,fe", [(2, "P1"), (2, "P2"), (3, "P1"), (3, "P2")])
def test_laplace_properties(dim, fe):
    conn, co, _ = OM.structured(dim, fe, 1, 3)
    K = to_sp(oracle_csr("laplace", dim, fe, conn, co))
    scale = abs(K).max()
    assert abs(K - K.T).max() <= 1e-14 * scale
    assert abs(K @ np.ones(K.shape[0])).max() <= 1e-13 * scale
    lin = 0.3 + co @ np.arange(1, dim + 1)
    interior = np.all((co > 1e-12) & (co < 1 - 1e-12), axis=1)
    assert abs((K @ lin)[interior]).max() <= 1e-12 * scale
    if fe == "P2":  # x^T K x = int |grad u|^2 exactly for quadratic u
        uq = co[:, 0] ** 2 + co[:, 0] * co[:, 1]
        exact = 4.0 / 3.0 + 1.0 + 1.0 / 3.0 + 1.0 / 3.0  # int (2x+y)^2 + x^2 over the unit square/cube
        assert abs(uq @ (K @ uq) - exact) <= 1e-12


@pytest.mark.parametrize("dim,fe", [(2, "P1"), (2, "P2"), (3, "P1"), (3, "P2")])
def test_elasticity_rigid_body_modes(dim, fe):
    conn, co, _ = OM.structured(dim, fe, 1, 2)
    n = co.shape[0]
    K = to_sp(oracle_csr("linelas", dim, fe, conn, co, lam=8e6, mu=2e6))
    scale = abs(K).max()
    assert abs(K - K.T).max() <= 1e-14 * scale
    modes = []
    for d in range(dim):
        m = np.zeros((n, dim)); m[:, d] = 1.0; modes.append(m.ravel())
    for a in range(dim):
        for b in range(a + 1, dim):
            m = np.zeros((n, dim)); m[:, a] = -co[:, b]; m[:, b] = co[:, a]; modes.append(m.ravel())
    for m in modes:
        assert abs(K @ m).max() <= 1e-13 * scale * max(1.0, abs(m).max())


def test_elasticity_closed_form_matches_epsilon_form():
    """K^{ab}_ij = |det| sum_q w_q [mu(delta_ab g_i.g_j + g_i[b] g_j[a]) + lambda g_i[a] g_j[b]]."""
    rng = np.random.default_rng(7)
    co1 = rng.uniform(0, 1, (4, 3))
    conn, co, _ = OM.p2_of_p1(np.array([[0, 1, 2, 3]], dtype=np.int32), co1)
    lam, mu = 3.7, 1.3
    K = to_sp(oracle_csr("linelas", 3, "P2", conn, co, lam=lam, mu=mu)).toarray()
    dphi, w = O.get_dphi(3, "P2", 2)
    B = (co1[1:] - co1[0]).T
    Binv = np.linalg.inv(B)
    g = dphi @ Binv  # [q, i, d]
    adet = abs(np.linalg.det(B))
    loc = conn[0]
    Kc = np.zeros_like(K)
    for i in range(10):
        for j in range(10):
            for a in range(3):
                for b in range(3):
                    s = sum(w[q] * (mu * ((a == b) * g[q, i] @ g[q, j] + g[q, i, b] * g[q, j, a]) +
                                    lam * g[q, i, a] * g[q, j, b]) for q in range(len(w)))
                    Kc[3 * loc[i] + a, 3 * loc[j] + b] = adet * s
    assert np.linalg.norm(K - Kc) <= 1e-14 * np.linalg.norm(Kc)


@pytest.mark.parametrize("dim,fe", [(2, "P1"), (2, "P2"), (3, "P1"), (3, "P2")])
def test_advection_properties(dim, fe):
    conn, co, _ = OM.structured(dim, fe, 1, 2)
    n = co.shape[0]
    u = random_u(dim, n)
    N = to_sp(oracle_csr("advection", dim, fe, conn, co, u=u))
    assert abs(N @ np.ones(dim * n)).max() <= 1e-13 * max(abs(N).max(), 1e-300)      # sum_j phi_j = 1
    Wc = to_sp(oracle_csr("advection_in_u", dim, fe, conn, co, u=np.tile(np.arange(1.0, dim + 1), n)))
    assert abs(Wc).max() <= 1e-14                                                      # constant u -> W = 0


def test_advection_in_u_linear_field_is_J_times_mass():
    """u = (y, -x, 0.1 z): W^{d1 d2} = J[d1][d2] * Mass with J = du/dx (P2 reproduces linear fields)."""
    conn, co, _ = OM.structured(3, "P2", 1, 2)
    n = co.shape[0]
    u = np.stack([co[:, 1], -co[:, 0], 0.1 * co[:, 2]], axis=1).ravel()
    W = to_sp(oracle_csr("advection_in_u", 3, "P2", conn, co, u=u)).tocsr()
    J = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 0.1]])
    mass = W[2::3, 2::3] / 0.1
    assert abs(mass.sum() - 1.0) <= 1e-12                                              # volume of the cube
    for d1 in range(3):
        for d2 in range(3):
            assert abs(W[d1::3, d2::3] - J[d1, d2] * mass).max() <= 1e-13


@pytest.mark.parametrize("dim,fe1,fe2", [(2, "P2", "P1"), (3, "P2", "P1"), (3, "P1", "P1"), (2, "P1", "P1")])
def test_div_properties(dim, fe1, fe2):
    conn, co, _ = OM.structured(dim, fe1, 1, 2)
    c1 = OM.p1_of_p2(conn, dim) if fe1 == "P2" else conn
    n = co.shape[0]
    np1 = int(c1.max()) + 1
    B = to_sp(oracle_csr("div", dim, fe1, conn, co, fe2=fe2, conn2=c1), dim * n)
    BT = to_sp(oracle_csr("divT", dim, fe1, conn, co, fe2=fe2, conn2=c1), np1)
    assert abs(B - BT.T).max() == 0.0                                                  # same expression, bitwise
    # sum_i psi_i = 1  ->  column sums = int d_d phi_j ; applied to a linear field u = x e_0: div u = 1
    u = np.zeros((n, dim)); u[:, 0] = co[:, 0]
    assert abs((B @ u.ravel()).sum() - 1.0) <= 1e-12
    # fast variant inserts the same values
    Bf = O.Matrix(np1, 64); BTf = O.Matrix(dim * n)
    O.assembly_div_divT(dim, fe1, fe2, conn, co, np.arange(n), c1, np.arange(np1), Bf, BTf, fast=True)
    assert abs(to_sp(Bf.csr(), dim * n) - B).max() <= 1e-18


def test_partition_invariance_of_global_matrix():
    """Globally summed matrix is identical (to rounding) for 1 and 8 element partitions."""
    M = 2
    conn, co, gid = OM.structured(3, "P2", 1, 2 * M)
    n = co.shape[0]
    A1 = O.Matrix(3 * n, 64); O.assembly_linelas(3, "P2", conn, co, gid, 8e6, 2e6, A1)
    A8 = O.Matrix(3 * n, 64)
    for c, x, g in OM.structured_global(3, "P2", 2, M):
        O.assembly_linelas(3, "P2", c, x, g, 8e6, 2e6, A8)
    r1, c1, v1 = A1.csr(); r8, c8, v8 = A8.csr()
    assert np.array_equal(r1, r8) and np.array_equal(c1, c8)
    assert np.linalg.norm(v1 - v8) <= 1e-13 * np.linalg.norm(v1)


def test_unstructured_dfg_mesh_properties():
    conn, co = mesh_dfg("P2")
    assert conn.shape == (5476, 10)
    K = to_sp(oracle_csr("laplace", 3, "P2", conn, co))
    assert abs(K @ np.ones(K.shape[0])).max() <= 1e-12 * abs(K).max()
    vol = 2.5 * 0.41 * 0.41 - np.pi * 0.05**2 * 0.41
    uq = co[:, 0]
    assert abs(uq @ (K @ uq) - vol) <= 2e-3 * vol    # polyhedral approximation of the cylinder hole


def test_quadrature_exactness():
    for dim, deg, nq in ((2, 1, 1), (2, 2, 3), (2, 5, 7), (3, 1, 1), (3, 3, 5), (3, 5, 15)):
        pts, w = O.quadrature(dim, deg)
        assert len(w) == nq
        assert abs(w.sum() - (0.5 if dim == 2 else 1.0 / 6.0)) <= 1e-14  # the 7-pt constants are 15-digit literals
        # int x^deg over the reference simplex = deg! / (deg+dim)!
        from math import factorial
        exact = factorial(deg) / factorial(deg + dim)
        assert abs((w * pts[:, 0] ** deg).sum() - exact) <= 1e-13
    assert O.determine_degree(3, "P2", "P2", O.GRAD, O.GRAD) == 2
    assert O.determine_degree(3, "P1", "P1", O.GRAD, O.GRAD) == 1
    assert O.determine_degree(3, "P2", "P2", O.GRAD, O.STD, O.determine_degree1(3, "P2", O.STD)) == 5
    assert O.determine_degree(3, "P2", "P2", O.STD, O.STD, O.determine_degree1(3, "P2", O.GRAD)) == 5
    assert O.determine_degree(3, "P1", "P1", O.STD, O.STD, O.determine_degree1(3, "P1", O.GRAD)) == 3
    assert O.determine_degree(3, "P2", "P1", O.GRAD, O.STD) == 2


def test_insert_fill_complete_semantics():
    """Appendix C: duplicates are kept until fillComplete, which sorts by column and sums."""
    A = O.Matrix(3, 2)
    A.insertGlobalValues(1, [2, 0, 2], [1.0, 2.0, 3.0])
    A.insertGlobalValues(1, [0], [0.5])
    A.insertGlobalValues(0, [1], [0.0])          # explicit zeros stay in the pattern
    rp, c, v = A.csr()
    assert rp.tolist() == [0, 1, 3, 3] and c.tolist() == [1, 0, 2] and v.tolist() == [0.0, 2.5, 4.0]
    with pytest.raises(IndexError):
        A.insertGlobalValues(5, [0], [1.0])


def test_unsupported_fe_type_is_an_error():
    conn, co, gid = OM.structured(3, "P1", 1, 1)
    with pytest.raises(ValueError):
        O.assembly_laplace(3, "P0", conn, co, gid, O.Matrix(co.shape[0]))


@pytest.mark.parametrize("dim,fe,M", [(2, "P1", 5), (2, "P2", 4), (3, "P1", 3), (3, "P2", 2)])
def test_mass_matrix_known_answers(dim, fe, M):
    """assemblyMass: 1^T M 1 = |Omega| (partition of unity), M symmetric, and x^T M x = int x_0^2 = 1/3 exactly for the
    coordinate function (P2 and P1 rules are exact for it only with P2; P1 checks the volume and symmetry)."""
    from oracle import mesh as OM
    conn, coords, gid = OM.structured(dim, fe, 1, M)
    n = coords.shape[0]
    A = O.Matrix(n, 64)
    O.assembly_mass(dim, fe, conn, coords, gid, A, False)
    S = A.scipy().tocsr()
    one = np.ones(n)
    assert abs(one @ (S @ one) - 1.0) < 1e-13
    assert abs(S - S.T).max() < 1e-15
    assert S.diagonal().min() > 0
    if fe == "P2":
        x = coords[:, 0]
        assert abs(x @ (S @ x) - 1.0 / 3.0) < 1e-13
    V = O.Matrix(dim * n, 64)
    O.assembly_mass(dim, fe, conn, coords, gid, V, True)
    Sv = V.scipy().tocsr()
    for d in range(dim):   # "Vector": the scalar matrix on every diagonal block, nothing off the diagonal blocks
        blk = Sv[d::dim, :][:, d::dim]
        assert abs(blk - S).max() == 0.0
    assert Sv.nnz == dim * S.nnz


@pytest.mark.parametrize("dim,M", [(2, 5), (3, 3)])
def test_bd_stabilization_known_answers(dim, M):
    """assemblyBDStabilization (FE_def.hpp:2151-2220): C = M_P1 - |T| * scale * 1 1^T per element.  With scale = 1/9
    (2D: 3 nodes) and 1/16 (3D: 4 nodes) the element row sums of the mean-value term equal those of the mass matrix, so
    C 1 = 0; C is symmetric and C = M - (rank-one element terms) is positive semi-definite on mean-free vectors."""
    from oracle import mesh as OM
    conn, coords, gid = OM.structured(dim, "P1", 1, M)
    n = coords.shape[0]
    A = O.Matrix(n, 64)
    O.assembly_bdstab(dim, "P1", conn, coords, gid, A)
    S = A.scipy().tocsr()
    if dim == 2:   # 3 * (1/2 * 1/9) = 1/6 = row sum of the reference P1 mass matrix
        assert abs(S @ np.ones(n)).max() < 1e-15
    else:          # 3D: 4 * (1/6 * 1/16) = 1/24 = row sum of the reference P1 mass matrix
        assert abs(S @ np.ones(n)).max() < 1e-15
    assert abs(S - S.T).max() < 1e-16
    Mm = O.Matrix(n, 64)
    O.assembly_mass(dim, "P1", conn, coords, gid, Mm, False)
    D = (Mm.scipy().tocsr() - S).toarray()
    assert D.min() >= 0 and D.max() > 0 and np.linalg.matrix_rank(D) <= conn.shape[0]
    with pytest.raises(ValueError):
        O.assembly_bdstab(dim, "P2", conn, coords, gid, O.Matrix(n, 64))


@pytest.mark.parametrize("dim,fe,M", [(2, "P1", 4), (2, "P2", 3), (3, "P1", 3), (3, "P2", 2)])
def test_stress_known_answers(dim, fe, M):
    """assemblyStress (FE_def.hpp:2407-2735): K^{ab}_ij = int f (delta_ab grad phi_i . grad phi_j + d_b phi_i d_a phi_j).
    With f = 1 it is the mu-part of assemblyLinElasXDim (lambda = 0, mu = 1); it is symmetric for any f, annihilates the
    rigid translations and rotations (2 eps(u) = 0), and scales linearly with a constant f."""
    from oracle import mesh as OM
    conn, coords, gid = OM.structured(dim, fe, 1, M)
    n = coords.shape[0]

    def stress(func):
        A = O.Matrix(dim * n, 64)
        O.assembly_stress(dim, fe, conn, coords, gid, func, A)
        return A.scipy().tocsr()

    S1 = stress(lambda x: 1.0)
    E = O.Matrix(dim * n, 64)
    O.assembly_linelas(dim, fe, conn, coords, gid, 0.0, 1.0, E)
    Es = E.scipy().tocsr()
    assert abs(S1 - Es).max() <= 1e-14 * abs(Es).max()
    Sf = stress(lambda x: 1.0 + 0.5 * x[0] + 0.25 * x[1] * x[1])
    scale = abs(Sf).max()
    assert abs(Sf - Sf.T).max() <= 1e-14 * scale
    for d in range(dim):                                   # translations
        t = np.zeros(dim * n); t[d::dim] = 1.0
        assert abs(Sf @ t).max() <= 1e-13 * scale
    rot = np.zeros(dim * n)                                # rotation about the last axis: u = (-y, x[, 0])
    rot[0::dim], rot[1::dim] = -coords[:, 1], coords[:, 0]
    assert abs(Sf @ rot).max() <= 1e-12 * scale
    assert abs(stress(lambda x: 2.5) - 2.5 * S1).max() <= 1e-14 * abs(S1).max()
