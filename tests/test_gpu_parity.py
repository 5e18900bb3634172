"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Pattern (rowptr/colind) must be bit-exact; values within a relative Frobenius error of 1e-12
(BASELINE.json north_star).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from util import TOL, mesh_dfg, mesh_structured, oracle_csr, random_u, rel_frobenius

pytestmark = pytest.mark.gpu

MODES = ["gather", "coloured", "atomic"]
LAM, MU = 8.0e6, 2.0e6   # steadyLinElas_Perf/parametersProblem.xml: nu = 0.4, mu = 2e6

# (id, dim, fe, mesh factory, pressure-mesh factory or None)
# the pressure (P1) mesh shares the element index with the velocity mesh (FE_def.hpp:1982-2016)
def _p1_of(dim, M):
    return lambda: mesh_structured(dim, "P1", M)


def _dfg_p1():
    return mesh_dfg("P1")


MESHES = [
    ("sq-P1", 2, "P1", lambda: mesh_structured(2, "P1", 7), _p1_of(2, 7)),
    ("sq-P2", 2, "P2", lambda: mesh_structured(2, "P2", 5), _p1_of(2, 5)),
    ("cube-P1", 3, "P1", lambda: mesh_structured(3, "P1", 4), _p1_of(3, 4)),
    ("cube-P2", 3, "P2", lambda: mesh_structured(3, "P2", 3), _p1_of(3, 3)),
    ("sq-P2-warp-shuffle", 2, "P2", lambda: mesh_structured(2, "P2", 6, warp=True, shuffle=True, seed=3), None),
    ("cube-P2-warp-shuffle", 3, "P2", lambda: mesh_structured(3, "P2", 3, warp=True, shuffle=True, seed=5), None),
    ("cube-P1-warp-shuffle", 3, "P1", lambda: mesh_structured(3, "P1", 4, warp=True, shuffle=True, seed=9), None),
    ("dfg-P2", 3, "P2", lambda: mesh_dfg("P2"), _dfg_p1),
    ("dfg-P1", 3, "P1", lambda: mesh_dfg("P1"), _dfg_p1),
]


@pytest.fixture(scope="module", params=MESHES, ids=[m[0] for m in MESHES])
def case(request, engine_ctx):
    from feddlib_b200 import Mesh, Pattern
    name, dim, fe, make, make_p = request.param
    conn, coords = make()
    mesh = Mesh(engine_ctx, dim, conn, coords)
    pat = Pattern(engine_ctx, mesh)
    return dict(name=name, dim=dim, fe=fe, conn=conn, coords=coords, mesh=mesh, pat=pat, ctx=engine_ctx,
                pressure=make_p)


def check(case, op, got_vals, rd, cd, mode, **kw):
    from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR
    rp_o, ci_o, v_o = oracle_csr(op, case["dim"], case["fe"], case["conn"], case["coords"], **kw)
    rp, ci = case["pat"].expand(rd, cd, mode)
    assert np.array_equal(rp, rp_o), "rowptr differs from the oracle"
    assert np.array_equal(ci, ci_o), "colind differs from the oracle"
    err = rel_frobenius(got_vals, v_o)
    assert err <= TOL, f"relative Frobenius error {err:.3e}"
    return err


@pytest.mark.parametrize("mode", MODES)
def test_laplace(case, mode):
    from feddlib_b200 import BLOCK_DIAG, BLOCK_SCALAR
    case["ctx"].set_scatter_mode(mode)
    check(case, "laplace", case["pat"].assemble_laplace(False), 1, 1, BLOCK_SCALAR)
    d = case["dim"]
    check(case, "laplace_vec", case["pat"].assemble_laplace(True), d, d, BLOCK_DIAG)


def test_rhs(case):
    """FE::assemblyRHS, constant source (FE_def.hpp:4694-4766): the load vector on the repeated
    map equals the oracle's; rows are added in the reference's element order, so the tolerance is that of the tables."""
    from oracle import oracle as O
    d = case["dim"]
    f = np.array([1.5, -2.0, 0.25])[:d]
    for vec in (False, True):
        for deg_func in (0, 1):
            want = O.assembly_rhs(d, case["fe"], case["conn"], case["coords"], f, deg_func, vec)
            got = case["pat"].assemble_rhs(f, deg_func, vec)
            assert got.shape == want.shape
            assert np.linalg.norm(got - want) <= TOL * np.linalg.norm(want)


@pytest.mark.parametrize("mode", MODES)
def test_mass(case, mode):
    """FE::assemblyMass, fieldType "Scalar" and "Vector" (FE_def.hpp:454-521)."""
    from feddlib_b200 import BLOCK_DIAG, BLOCK_SCALAR
    case["ctx"].set_scatter_mode(mode)
    check(case, "mass", case["pat"].assemble_mass(False), 1, 1, BLOCK_SCALAR)
    d = case["dim"]
    check(case, "mass_vec", case["pat"].assemble_mass(True), d, d, BLOCK_DIAG)


@pytest.mark.parametrize("mode", MODES)
def test_stress(case, mode):
    """FE::assemblyStress (FE_def.hpp:2407-2735): constant coefficient (all scatter modes, the
    row-gather kernels in gather mode) and a coefficient that varies inside the elements (quadrature loop)."""
    from feddlib_b200 import BLOCK_FULL
    from util import stress_coefficient
    case["ctx"].set_scatter_mode(mode)
    d = case["dim"]
    check(case, "stress", case["pat"].assemble_stress(1.0), d, d, BLOCK_FULL, func=lambda x: 1.0)
    check(case, "stress", case["pat"].assemble_stress(2.5), d, d, BLOCK_FULL, func=lambda x: 2.5)
    xyz = case["pat"].stress_points(case["conn"], case["coords"])
    coef = np.array([[stress_coefficient(x) for x in el] for el in xyz])
    check(case, "stress", case["pat"].assemble_stress(coef), d, d, BLOCK_FULL, func=stress_coefficient)


@pytest.mark.parametrize("mode", MODES)
def test_bd_stabilization(case, mode):
    """FE::assemblyBDStabilization (FE_def.hpp:2151-2220): P1 only, logic_error otherwise."""
    from feddlib_b200 import BLOCK_SCALAR, LogicError
    case["ctx"].set_scatter_mode(mode)
    if case["fe"] == "P1":
        check(case, "bdstab", case["pat"].assemble_bdstab(), 1, 1, BLOCK_SCALAR)
    else:
        with pytest.raises(LogicError):
            case["pat"].assemble_bdstab()


@pytest.mark.parametrize("mode", MODES)
def test_linear_elasticity(case, mode):
    from feddlib_b200 import BLOCK_FULL
    case["ctx"].set_scatter_mode(mode)
    d = case["dim"]
    check(case, "linelas", case["pat"].assemble_linelas(LAM, MU), d, d, BLOCK_FULL, lam=LAM, mu=MU)


@pytest.mark.parametrize("mode", MODES)
def test_advection_N_and_W(case, mode):
    from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL
    case["ctx"].set_scatter_mode(mode)
    d = case["dim"]
    u = random_u(d, case["coords"].shape[0])
    check(case, "advection", case["pat"].assemble_advection(u), d, d, BLOCK_DIAG, u=u)
    check(case, "advection_in_u", case["pat"].assemble_advection_in_u(u), d, d, BLOCK_FULL, u=u)


@pytest.mark.parametrize("mode", MODES)
def test_div_and_divT(case, mode):
    from feddlib_b200 import BLOCK_FULL, Mesh, Pattern, assemble_div_divT
    ctx = case["ctx"]
    ctx.set_scatter_mode(mode)
    d = case["dim"]
    conn = case["conn"]
    if case["pressure"] is None:
        pytest.skip("no element-aligned pressure mesh for the shuffled numbering")
    conn_p, coords_p = case["pressure"]()
    assert conn_p.shape[0] == conn.shape[0]
    npres = coords_p.shape[0]
    mesh_p = Mesh(ctx, d, conn_p, coords_p)
    patB, patBT = Pattern(ctx, mesh_p, case["mesh"]), Pattern(ctx, case["mesh"], mesh_p)
    vB, vBT = assemble_div_divT(ctx, patB, patBT)
    for op, pat, vals, rd, cd in (("div", patB, vB, 1, d), ("divT", patBT, vBT, d, 1)):
        rp_o, ci_o, v_o = oracle_csr(op, d, case["fe"], conn, case["coords"], fe2="P1", conn2=conn_p)
        rp, ci = pat.expand(rd, cd, BLOCK_FULL)
        assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o)
        assert rel_frobenius(vals, v_o) <= TOL
    # B^T values equal transpose(B) bitwise: same expression on both sides (Appendix D)
    import scipy.sparse as sp
    rpB, ciB = patB.expand(1, d, BLOCK_FULL)
    rpT, ciT = patBT.expand(d, 1, BLOCK_FULL)
    B = sp.csr_matrix((vB, ciB, rpB), shape=(npres, d * case["coords"].shape[0]))
    BT = sp.csr_matrix((vBT, ciT, rpT), shape=(d * case["coords"].shape[0], npres))
    if mode == "coloured":     # element-row kernels: the same expression on both sides
        assert abs(B - BT.T).max() <= 1e-15 * abs(B).max()
    elif mode == "gather":     # row-gather kernels: B and B^T sum their incidences in different orders
        assert abs(B - BT.T).max() <= 1e-13 * abs(B).max()


@pytest.mark.parametrize("mode", MODES)
def test_div_and_divT_p0_pressure(case, mode):
    """P0 pressure (FE_def.hpp:1954-1957, 2012-2013, 2039-2040): rows of B / columns of B^T on the element map, described to the
    engine as one pseudo-node per element.  2D only (FE::phi has no P0 case for dim 3); gather mode falls back to the
    coloured element-row kernels for this pair of spaces."""
    from feddlib_b200 import BLOCK_FULL, LogicError, Mesh, Pattern, assemble_div_divT
    ctx = case["ctx"]
    ctx.set_scatter_mode(mode)
    d, conn, coords = case["dim"], case["conn"], case["coords"]
    ne = conn.shape[0]
    conn0 = np.arange(ne, dtype=np.int32)[:, None]
    if d == 3:
        with pytest.raises(LogicError):
            Mesh(ctx, d, conn0, np.zeros((ne, d)))
        return
    mesh0 = Mesh(ctx, d, conn0, np.zeros((ne, d)))
    patB, patBT = Pattern(ctx, mesh0, case["mesh"]), Pattern(ctx, case["mesh"], mesh0)
    vB, vBT = assemble_div_divT(ctx, patB, patBT)
    for op, pat, vals, rd, cd in (("div", patB, vB, 1, d), ("divT", patBT, vBT, d, 1)):
        rp_o, ci_o, v_o = oracle_csr(op, d, case["fe"], conn, coords, fe2="P0", conn2=conn0)
        rp, ci = pat.expand(rd, cd, BLOCK_FULL)
        assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o)
        assert rel_frobenius(vals, v_o) <= TOL
    import scipy.sparse as sp
    rpB, ciB = patB.expand(1, d, BLOCK_FULL)
    rpT, ciT = patBT.expand(d, 1, BLOCK_FULL)
    B = sp.csr_matrix((vB, ciB, rpB), shape=(ne, d * coords.shape[0]))
    BT = sp.csr_matrix((vBT, ciT, rpT), shape=(d * coords.shape[0], ne))
    assert abs(B - BT.T).max() <= 1e-15 * abs(B).max()
    # divergence theorem per element: sum_j B_{e,(j,d)} x_j[d] summed over d = dim * |e| for u = x  (a KAT that needs no oracle)
    if case["fe"] == "P1":
        from feddlib_b200.mesh import element_volumes
        vol = np.abs(element_volumes(conn[:, : d + 1], coords))
        assert np.allclose(B @ coords.reshape(-1), d * vol, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("mode", MODES)
def test_ns_jacobian_fused_equals_sum_of_parts(case, mode):
    """rho*nu*A + rho*N + rho*W on the union pattern (NavierStokes_def.hpp:140-152, 297-313); gather mode is the
    fused row-gather kernel k_gatherx<X_NSJ> that bench.py times."""
    from feddlib_b200 import BLOCK_FULL
    import scipy.sparse as sp
    case["ctx"].set_scatter_mode(mode)
    d = case["dim"]
    n = case["coords"].shape[0]
    u = random_u(d, n, seed=77)
    rho, nu = 1.3, 1.0e-3
    parts = {}
    for op in ("laplace_vec", "advection", "advection_in_u"):
        rp, ci, v = oracle_csr(op, d, case["fe"], case["conn"], case["coords"], u=u)
        parts[op] = sp.csr_matrix((v, ci, rp), shape=(d * n, d * n))
    for newton in (False, True):
        ref = rho * nu * parts["laplace_vec"] + rho * parts["advection"]
        if newton:
            ref = ref + rho * parts["advection_in_u"]
        vals = case["pat"].assemble_ns_jacobian(u, rho, nu, newton)
        rp, ci = case["pat"].expand(d, d, BLOCK_FULL)
        got = sp.csr_matrix((vals, ci, rp), shape=(d * n, d * n))
        diff = (got - ref)
        err = np.sqrt(diff.multiply(diff).sum()) / np.sqrt(ref.multiply(ref).sum())
        assert err <= TOL, f"newton={newton}: {err:.3e}"


def test_gather_and_coloured_are_bitwise_reproducible(case):
    from feddlib_b200 import BLOCK_FULL
    for mode in ("gather", "coloured"):
        case["ctx"].set_scatter_mode(mode)
        a = case["pat"].assemble_linelas(LAM, MU)
        b = case["pat"].assemble_linelas(LAM, MU)
        assert np.array_equal(a, b), f"{mode} scatter is not run-to-run reproducible"


def test_pattern_node_level_properties(case):
    rp, ci = case["pat"].nodes()
    assert rp[0] == 0 and rp[-1] == ci.size == case["pat"].nnz_nodes
    for r in range(0, rp.size - 1, max(1, (rp.size - 1) // 50)):
        seg = ci[rp[r]: rp[r + 1]]
        assert np.all(np.diff(seg) > 0) and r in seg              # ascending, diagonal present
    import scipy.sparse as sp
    P = sp.csr_matrix((np.ones(ci.size), ci, rp), shape=(rp.size - 1,) * 2)
    assert (P - P.T).nnz == 0                                       # symmetric pattern
    assert case["pat"].max_row_len == np.diff(rp).max()
    assert 1 <= case["pat"].n_colours <= 256
