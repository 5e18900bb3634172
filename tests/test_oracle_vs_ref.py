"""Pins the oracle restatement (oracle/fedd_oracle.c) against the REFERENCE's own code: oracle/_ref is
FEDDLib's FE_def.hpp hot-path routines compiled from /root/reference against mock Trilinos containers
(oracle/ref_shim).  Both use the same accumulate emulation, so the comparison isolates the element loops:
they must agree BITWISE.  Skipped where oracle/_ref has not been built (no reference tree)."""
import numpy as np
import pytest

from oracle import mesh as OM
from oracle import oracle as O
from oracle import ref as R

from util import mesh_dfg, mesh_structured, oracle_csr, random_u, stress_coefficient

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [(2, "P1", lambda: mesh_structured(2, "P1", 5)), (2, "P2", lambda: mesh_structured(2, "P2", 4, warp=True)),
         (3, "P1", lambda: mesh_structured(3, "P1", 3, warp=True)), (3, "P2", lambda: mesh_structured(3, "P2", 2)),
         (3, "P2", lambda: mesh_structured(3, "P2", 2, warp=True, shuffle=True, seed=2)),
         (3, "P2", lambda: tuple(a[:600] if a.ndim == 2 and a.shape[1] == 10 else a for a in mesh_dfg("P2")))]


@pytest.mark.parametrize("dim,fe,make", CASES, ids=[f"{c[0]}d-{c[1]}-{i}" for i, c in enumerate(CASES)])
def test_restatement_is_bitwise_equal_to_reference_code(dim, fe, make):
    conn, coords = make()
    u = random_u(dim, coords.shape[0])
    for op, kw in (("mass", {}), ("mass_vec", {}), ("laplace", {}), ("laplace_vec", {}), ("linelas", dict(lam=8e6, mu=2e6)),
                   ("advection", dict(u=u)), ("advection_in_u", dict(u=u))):
        a = oracle_csr(op, dim, fe, conn, coords, **kw)
        b = R.assemble(op, dim, fe, conn, coords, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), op
        assert np.array_equal(a[2], b[2]), f"{op}: max abs diff {np.abs(a[2] - b[2]).max():.3e}"
    for f in (lambda x: 1.0, stress_coefficient):   # assemblyStress (FE_def.hpp:2407-2735), constant and varying coefficient
        a = oracle_csr("stress", dim, fe, conn, coords, func=f)
        b = R.assemble_stress(dim, fe, conn, coords, f)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), "stress"
    if fe == "P1":  # assemblyBDStabilization (FE_def.hpp:2151-2220) is P1 only
        a = oracle_csr("bdstab", dim, fe, conn, coords)
        b = R.assemble("bdstab", dim, fe, conn, coords)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), "bdstab"


@pytest.mark.parametrize("dim,fe1", [(2, "P2"), (3, "P2"), (3, "P1"), (2, "P1")])
def test_div_restatement_matches_reference(dim, fe1):
    conn, coords = mesh_structured(dim, fe1, 3)
    conn_p, _ = mesh_structured(dim, "P1", 3)
    for op, fast in (("div", False), ("div_fast", True)):
        (B, BT) = R.assemble(op, dim, fe1, conn, coords, fe2="P1", conn2=conn_p)
        n = coords.shape[0]; np_ = int(conn_p.max()) + 1
        Bo, BTo = O.Matrix(np_, 64), O.Matrix(dim * n)
        O.assembly_div_divT(dim, fe1, "P1", conn, coords, np.arange(n), conn_p, np.arange(np_), Bo, BTo, fast=fast)
        for got, ref in ((Bo.csr(), B), (BTo.csr(), BT)):
            assert all(np.array_equal(x, y) for x, y in zip(got, ref)), op


@pytest.mark.parametrize("dim,fe1", [(2, "P2"), (2, "P1")])
def test_div_p0_pressure_restatement_matches_reference(dim, fe1):
    """P0 pressure (FE_def.hpp:1954-1957, 2012-2013, 2039-2040): the rows of B / columns of B^T live on the ELEMENT map;
    described to both sides as one pseudo-node per element whose global ids are the element map.  2D only: FE::phi has
    no P0 case for dim 3 (the reference returns uninitialised values there)."""
    conn, coords = mesh_structured(dim, fe1, 3, warp=True)
    ne, n = conn.shape[0], coords.shape[0]
    conn0 = np.arange(ne, dtype=np.int32)[:, None]
    egid = np.random.default_rng(7).permutation(ne).astype(np.int64)      # element map with permuted global ids
    for op, fast in (("div", False), ("div_fast", True)):
        (B, BT) = R.assemble(op, dim, fe1, conn, coords, fe2="P0", conn2=conn0, gid2=egid)
        Bo, BTo = O.Matrix(ne, 64), O.Matrix(dim * n)
        O.assembly_div_divT(dim, fe1, "P0", conn, coords, np.arange(n), conn0, egid, Bo, BTo, fast=fast)
        for got, ref in ((Bo.csr(), B), (BTo.csr(), BT)):
            assert all(np.array_equal(x, y) for x, y in zip(got, ref)), op
        assert B[0][-1] == ne * conn.shape[1] * dim                         # every element row holds dim * nloc entries
    assert O.determine_degree(dim, fe1, "P0", 1, 0, 0) == R.determine_degree(dim, fe1, "P0", 1, 0, 0)


def test_tables_and_degrees_match_reference():
    for dim in (2, 3):
        for fe in ("P1", "P2"):
            for deg in ((1, 2, 5) if dim == 2 else (1, 3, 5)):
                d0, w0 = O.get_dphi(dim, fe, deg); d1, w1 = R.get_dphi(dim, fe, deg)
                assert np.array_equal(d0, d1) and np.array_equal(w0, w1)
                p0, _ = O.get_phi(dim, fe, deg); p1, _ = R.get_phi(dim, fe, deg)
                assert np.array_equal(p0, p1)
            for fe2 in ("P1", "P2"):
                for t1 in (0, 1):
                    for t2 in (0, 1):
                        for ex in (0, 1, 2):
                            assert O.determine_degree(dim, fe, fe2, t1, t2, ex) == R.determine_degree(dim, fe, fe2, t1, t2, ex)


def test_multi_rank_insertion_equals_reference():
    """8 ranks of the structured cube inserted into one global matrix, reference code vs restatement."""
    n = (2 * 2 * 2 + 1) ** 3
    A = O.Matrix(3 * n, 64)
    import ctypes as C
    L = R.lib()
    h = L.fo_matrix_new(3 * n, 64)
    hB = L.fo_matrix_new(1, 1)
    for c, x, g in OM.structured_global(3, "P2", 2, 2):
        O.assembly_linelas(3, "P2", c, x, g, 8e6, 2e6, A)
        c = np.ascontiguousarray(c); x = np.ascontiguousarray(x); g = np.ascontiguousarray(g)
        rc = L.ref_assemble(2, 3, b"P2", b"P2", c.shape[0], R._p(c), 10, R._p(x), x.shape[0], R._p(g), None, 0, 0, None,
                            None, 8e6, 2e6, h, hB)
        assert rc == 0
    a = A.csr(); b = R._csr(L, h, 3 * n)
    L.fo_matrix_free(h); L.fo_matrix_free(hB)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_reference_error_behaviour():
    conn, coords = mesh_structured(2, "P1", 2)
    with pytest.raises(ValueError, match="Not implemented for P0"):
        R.assemble("laplace", 2, "P0", conn, coords)


@pytest.mark.parametrize("dim,fe,make", CASES, ids=[f"{c[0]}d-{c[1]}-{i}" for i, c in enumerate(CASES)])
def test_rhs_restatement_is_bitwise_equal_to_reference_code(dim, fe, make):
    """FE::assemblyRHS (constant source, FE_def.hpp:4694-4766), "Scalar" and "Vector", function degrees 0..2."""
    conn, coords = make()
    f = np.array([1.5, -2.0, 0.25])[:dim]
    for vec in (False, True):
        for deg_func in (0, 1, 2):
            a = O.assembly_rhs(dim, fe, conn, coords, f, deg_func, vec)
            b = R.assemble_rhs(dim, fe, conn, coords, f, deg_func, vec)
            assert np.array_equal(a, b), (vec, deg_func)
            assert abs(a.reshape(-1, dim if vec else 1).sum(axis=0) - f[: dim if vec else 1] * _volume(dim, conn, coords)).max() < 1e-12


def _volume(dim, conn, coords):
    v = coords[conn[:, 1:dim + 1]] - coords[conn[:, :1]]
    return float(np.abs(np.linalg.det(v)).sum() / (2 if dim == 2 else 6))
