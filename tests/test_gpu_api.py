"""GPU tests of the reference-facing interface mirror (FE / Domain / Matrix) and of the C ABI's error
behaviour, edge cases included (empty mesh, wrong FE type, mismatched maps)."""
import numpy as np
import pytest

from util import TOL, oracle_csr, random_u, rel_frobenius

pytestmark = pytest.mark.gpu


def test_fe_interface_laplace_like_reference_driver(engine_ctx):
    """Reads like feddlib/core/FE/tests/fe.cpp:60-101 + Laplace::assemble (Laplace_def.hpp:36-60)."""
    from feddlib_b200 import FE, Domain, Matrix
    dim, FEType, M = 2, "P1", 16
    domain = Domain.buildMesh(dim, FEType, 1, M)
    fe = FE(ctx=engine_ctx)
    fe.addFE(domain)
    A = Matrix(domain.getMapUnique(), domain.getApproxEntriesPerRow())
    fe.assemblyLaplace(dim, FEType, 2, A)
    assert A.isFillComplete()
    rp, ci, v = oracle_csr("laplace", dim, FEType, domain.getElementsC(), domain.getPointsRepeated())
    assert np.array_equal(A.rowptr, rp) and np.array_equal(A.colind, ci)
    assert rel_frobenius(A.numpy_values(), v) <= TOL
    S = A.toScipy()
    assert abs(S @ np.ones(S.shape[1])).max() < 1e-12
    # post-assembly contract of the problem classes: resumeFill -> scale -> fillComplete(dom, rng)
    A.resumeFill(); A.scale(0.5); A.fillComplete(domain.getMapUnique(), domain.getMapUnique())
    assert rel_frobenius(A.numpy_values(), 0.5 * v) <= TOL


def test_fe_interface_all_entry_points(engine_ctx):
    from feddlib_b200 import FE, Domain, Map, Matrix
    from feddlib_b200 import mesh as PM
    dim, M = 3, 2
    dv = Domain.buildMesh(dim, "P2", 1, M)
    conn_p = PM.vertex_connectivity(dv.getElementsC(), dim)
    # pressure domain on the vertex nodes: keep only referenced nodes, renumbered 0..n_p-1
    used = np.unique(conn_p)
    renum = -np.ones(dv.getPointsRepeated().shape[0], dtype=np.int64); renum[used] = np.arange(used.size)
    dp = Domain(dim, "P1", renum[conn_p].astype(np.int32), dv.getPointsRepeated()[used], Map(np.arange(used.size)))
    fe = FE(ctx=engine_ctx)
    fe.addFE(dv); fe.addFE(dp)
    n = dv.getMapUnique().getNodeNumElements()
    u = random_u(dim, n)
    conn, co = dv.getElementsC(), dv.getPointsRepeated()

    K = Matrix(dv.getMapVecFieldUnique(), dim * dv.getApproxEntriesPerRow())
    fe.assemblyLinElasXDim(dim, "P2", K, 8e6, 2e6, True)
    assert rel_frobenius(K.numpy_values(), oracle_csr("linelas", dim, "P2", conn, co, lam=8e6, mu=2e6)[2]) <= TOL

    A = Matrix(dv.getMapVecFieldUnique(), dim * dv.getApproxEntriesPerRow())
    fe.assemblyLaplaceVecField(dim, "P2", 2, A, True)
    assert rel_frobenius(A.numpy_values(), oracle_csr("laplace_vec", dim, "P2", conn, co)[2]) <= TOL

    N = Matrix(dv.getMapVecFieldUnique(), dim * dv.getApproxEntriesPerRow())
    fe.assemblyAdvectionVecField(dim, "P2", N, u, True)
    assert rel_frobenius(N.numpy_values(), oracle_csr("advection", dim, "P2", conn, co, u=u)[2]) <= TOL

    W = Matrix(dv.getMapVecFieldUnique(), dim * dv.getApproxEntriesPerRow())
    fe.assemblyAdvectionInUVecField(dim, "P2", W, u, True)
    assert rel_frobenius(W.numpy_values(), oracle_csr("advection_in_u", dim, "P2", conn, co, u=u)[2]) <= TOL

    B = Matrix(dp.getMapUnique(), dim * dv.getApproxEntriesPerRow())
    BT = Matrix(dv.getMapVecFieldUnique(), dp.getApproxEntriesPerRow())
    fe.assemblyDivAndDivTFast(dim, "P2", "P1", 2, B, BT, dv.getMapVecFieldUnique(), dp.getMapUnique(), True)
    conn_p_local = dp.getElementsC()
    ro = oracle_csr("div", dim, "P2", conn, co, fe2="P1", conn2=conn_p_local)
    assert np.array_equal(B.rowptr, ro[0]) and np.array_equal(B.colind, ro[1])
    assert rel_frobenius(B.numpy_values(), ro[2]) <= TOL
    assert B.domainMap.isSameAs(dv.getMapVecFieldUnique()) and B.rangeMap.isSameAs(dp.getMapUnique())
    B.resumeFill(); B.scale(-1.0); B.fillComplete(dv.getMapVecFieldUnique(), dp.getMapUnique())   # NavierStokes_def.hpp:217-221
    assert rel_frobenius(B.numpy_values(), -ro[2]) <= TOL


def test_fe_interface_p0_pressure(engine_ctx):
    """assemblyDivAndDivT with FEType2 = "P0" through the reference-interface mirror: rows of B on the element map
    (FE_def.hpp:1954-1957); 3D raises like any combination the reference cannot evaluate (FE::phi has no P0 case there)."""
    from feddlib_b200 import FE, Domain, LogicError, Map, Matrix
    dim = 2
    dv = Domain.buildMesh(dim, "P2", 1, 5)
    ne = dv.getElementsC().shape[0]
    egid = np.random.default_rng(3).permutation(ne).astype(np.int64)
    d0 = Domain.p0_of(dv, Map(egid))
    fe = FE(ctx=engine_ctx)
    fe.addFE(dv); fe.addFE(d0)
    B = Matrix(d0.getMapUnique(), dim * dv.getApproxEntriesPerRow())
    BT = Matrix(dv.getMapVecFieldUnique(), 8)
    fe.assemblyDivAndDivT(dim, "P2", "P0", 2, B, BT, dv.getMapVecFieldUnique(), d0.getMapUnique(), True)
    conn, co = dv.getElementsC(), dv.getPointsRepeated()
    n = co.shape[0]
    from oracle import oracle as O
    Bo, BTo = O.Matrix(ne, 64), O.Matrix(dim * n)
    conn0 = np.arange(ne, dtype=np.int32)[:, None]
    O.assembly_div_divT(dim, "P2", "P0", conn, co, np.arange(n), conn0, np.arange(ne), Bo, BTo)   # local numbering: gid == lid
    for M_, ref in ((B, Bo.csr()), (BT, BTo.csr())):
        assert np.array_equal(M_.rowptr, ref[0]) and np.array_equal(M_.colind, ref[1])
        assert rel_frobenius(M_.numpy_values(), ref[2]) <= TOL
    assert B.getMap().isSameAs(d0.getMapUnique()) and d0.getElementMap().isSameAs(Map(egid))
    d3 = Domain.buildMesh(3, "P2", 1, 2)
    fe3 = FE(ctx=engine_ctx)
    fe3.addFE(d3)
    with pytest.raises(LogicError):
        fe3.assemblyDivAndDivT(3, "P2", "P0", 2, B, BT, None, None, True)


def test_error_behaviour_matches_reference(engine_ctx):
    from feddlib_b200 import FE, Domain, LogicError, Matrix
    d = Domain.buildMesh(2, "P1", 1, 3)
    fe = FE(ctx=engine_ctx)
    A = Matrix(d.getMapUnique(), 20)
    with pytest.raises(LogicError, match="Use addFE"):          # FE_def.hpp:6950
        fe.assemblyLaplace(2, "P1", 2, A)
    fe.addFE(d)
    with pytest.raises(LogicError, match="Not implemented for P0"):   # FE_def.hpp:610
        fe.assemblyLaplace(2, "P0", 2, A)
    with pytest.raises(LogicError, match="Use addFE"):
        fe.assemblyLaplace(3, "P1", 2, A)
    with pytest.raises(LogicError):                              # wrong row map (vector field on a scalar map)
        fe.assemblyLaplaceVecField(2, "P1", 2, A)
    with pytest.raises(LogicError, match="numberMV"):            # FE_def.hpp:1691
        fe.assemblyAdvectionVecField(2, "P1", Matrix(d.getMapVecFieldUnique(), 40), np.zeros((32, 2)))


def test_c_abi_argument_errors(engine_ctx):
    from feddlib_b200 import LogicError, Mesh, Pattern
    co = np.zeros((4, 3))
    with pytest.raises(LogicError, match="only P1/P2"):
        Mesh(engine_ctx, 3, np.zeros((1, 8), dtype=np.int32), co)            # Q1 hexahedron: not implemented
    with pytest.raises(LogicError, match="out of range"):
        Mesh(engine_ctx, 3, np.array([[0, 1, 2, 9]], dtype=np.int32), co)
    m3 = Mesh(engine_ctx, 3, np.array([[0, 1, 2, 3]], dtype=np.int32), np.eye(4, 3))
    m2 = Mesh(engine_ctx, 2, np.array([[0, 1, 2]], dtype=np.int32), np.eye(3, 2))
    with pytest.raises(LogicError, match="share the element list"):
        Pattern(engine_ctx, m3, m2)
    p = Pattern(engine_ctx, m3)
    with pytest.raises(LogicError, match="unsupported dof layout"):
        p.nnz(4, 4, 2)


def test_empty_and_single_element_meshes(engine_ctx):
    from feddlib_b200 import Mesh, Pattern
    empty = Mesh(engine_ctx, 3, np.zeros((0, 4), dtype=np.int32), np.zeros((5, 3)))
    p = Pattern(engine_ctx, empty)
    assert p.nnz_nodes == 0 and p.n_rows == 5
    for mode in ("gather", "coloured", "atomic"):
        engine_ctx.set_scatter_mode(mode)
        assert p.assemble_laplace().size == 0
    # a single P2 tetrahedron with random vertices
    from oracle import mesh as OM
    rng = np.random.default_rng(11)
    conn, co, _ = OM.p2_of_p1(np.array([[0, 1, 2, 3]], dtype=np.int32), rng.uniform(0, 1, (4, 3)))
    pat = Pattern(engine_ctx, Mesh(engine_ctx, 3, conn, co))
    for mode in ("gather", "coloured", "atomic"):
        engine_ctx.set_scatter_mode(mode)
        got = pat.assemble_linelas(3.0, 1.5)
        assert rel_frobenius(got, oracle_csr("linelas", 3, "P2", conn, co, lam=3.0, mu=1.5)[2]) <= TOL


def test_ragged_mesh_with_unused_nodes(engine_ctx):
    """Nodes that no element references still get (empty) rows, like an owned node without local elements."""
    from feddlib_b200 import BLOCK_SCALAR, Mesh, Pattern
    conn = np.array([[0, 2, 5], [2, 5, 6]], dtype=np.int32)
    co = np.array([[0, 0], [9, 9], [1, 0], [9, 9], [9, 9], [0, 1], [1, 1.0]])
    pat = Pattern(engine_ctx, Mesh(engine_ctx, 2, conn, co))
    rp, ci = pat.expand(1, 1, BLOCK_SCALAR)
    rpo, cio, vo = oracle_csr("laplace", 2, "P1", conn, co)
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    for mode in ("gather", "coloured", "atomic"):
        engine_ctx.set_scatter_mode(mode)
        assert rel_frobenius(pat.assemble_laplace(), vo) <= TOL


@pytest.mark.parametrize("n", [12, 40])
def test_large_vertex_and_edge_stars(engine_ctx, n):
    """A double fan: 2n tetrahedra around the centre vertex, n around each of the edges (centre, top) and (centre, bottom).
    n = 40 exceeds the 32 incidences the chain / ring ordering of the row-gather kernels handles (vertex star of 80,
    edge rings of 40): those rows take the unordered records and the generic kernels; n = 12 keeps everything ordered
    (closed rings of 12, a vertex star that decomposes into two closed fans)."""
    from feddlib_b200 import BLOCK_FULL, BLOCK_SCALAR, Mesh, Pattern
    from oracle import mesh as OM
    ang = 2 * np.pi * np.arange(n) / n
    ring = np.stack([np.cos(ang), np.sin(ang), 0.1 * np.sin(3 * ang)], axis=1)
    pts = np.concatenate([[[0.0, 0.0, 0.0], [0.1, 0.0, 1.0], [0.0, -0.1, -1.2]], ring])   # centre, top, bottom, ring
    c, t, b = 0, 1, 2
    conn1 = np.array([[c, apex, 3 + i, 3 + (i + 1) % n] for apex in (t, b) for i in range(n)], dtype=np.int32)
    conn, coords, _ = OM.p2_of_p1(conn1, pts)
    conn = np.ascontiguousarray(conn.astype(np.int32)); coords = np.ascontiguousarray(coords)
    pat = Pattern(engine_ctx, Mesh(engine_ctx, 3, conn, coords))
    rp, ci = pat.expand(3, 3, BLOCK_FULL)
    rpo, cio, vo = oracle_csr("linelas", 3, "P2", conn, coords, lam=3.0, mu=1.5)
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    lo = oracle_csr("laplace", 3, "P2", conn, coords)[2]
    for mode in ("gather", "coloured", "atomic"):
        engine_ctx.set_scatter_mode(mode)
        assert rel_frobenius(pat.assemble_linelas(3.0, 1.5), vo) <= TOL
        assert rel_frobenius(pat.assemble_laplace(), lo) <= TOL
    engine_ctx.set_scatter_mode("gather")
    a, b2 = pat.assemble_linelas(3.0, 1.5), pat.assemble_linelas(3.0, 1.5)
    assert np.array_equal(a, b2)           # write-once / fixed-order sums: bitwise reproducible


def test_device_resident_path_and_launch_counter(engine_ctx):
    import torch
    from feddlib_b200 import BLOCK_FULL, Mesh, Pattern
    from util import mesh_structured
    conn, co = mesh_structured(3, "P2", 3)
    pat = Pattern(engine_ctx, Mesh(engine_ctx, 3, conn, co))
    engine_ctx.set_scatter_mode("gather")
    vals = engine_ctx.empty_values(pat.nnz(3, 3, BLOCK_FULL))
    before = engine_ctx.launches
    pat.assemble_linelas_d(vals, 8e6, 2e6)
    engine_ctx.synchronize()
    assert engine_ctx.launches > before
    assert rel_frobenius(vals.cpu().numpy(), oracle_csr("linelas", 3, "P2", conn, co, lam=8e6, mu=2e6)[2]) <= TOL
    assert torch.isfinite(vals).all()


def test_multi_gpu_exchange_if_available():
    """Runs tests/dist_gpu_check.py under torchrun when the box has >= 2 GPUs (NCCL ghost-row exchange)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box; the multi-rank logic is covered on CPU by tests/test_dist_cpu.py")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 8)}",
                          "--master-addr", "127.0.0.1", "--master-port", "29561",
                          os.path.join(root, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


def test_alternative_star_kernels_in_a_subprocess():
    """k_fan (FEDDB200_FAN=1) and k_star (FEDDB200_STAR=1) are alternative row kernels of 3D P2 that are off by default (their
    knobs are read once per process): the same parity check as the default path, in a child process per setting.  The same
    goes for the measured-but-off protocols (fragments, Laplace points path)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from util import TOL, mesh_structured, mesh_dfg, oracle_csr, rel_frobenius
from feddlib_b200 import BLOCK_FULL, BLOCK_SCALAR, Context, Mesh, Pattern
ctx = Context(0)
for name, make in (("cube", lambda: mesh_structured(3, "P2", 4)), ("warp", lambda: mesh_structured(3, "P2", 3, warp=True, shuffle=True, seed=5)),
                   ("dfg", lambda: mesh_dfg("P2"))):
    conn, coords = make()
    pat = Pattern(ctx, Mesh(ctx, 3, conn, coords))
    l0 = ctx.launches
    for op, vals, rd, mode, kw in (("linelas", pat.assemble_linelas(8e6, 2e6), 3, BLOCK_FULL, dict(lam=8e6, mu=2e6)),
                                   ("laplace", pat.assemble_laplace(False), 1, BLOCK_SCALAR, {})):
        rp_o, ci_o, v_o = oracle_csr(op, 3, "P2", conn, coords, **kw)
        rp, ci = pat.expand(rd, rd, mode)
        assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o), (name, op)
        err = rel_frobenius(vals, v_o)
        assert err <= TOL, (name, op, err)
print("ok")
''' % (root, os.path.join(root, "tests"))
    # ... and the optional write / input protocols of the default kernels: boundary-sector fragments + k_stitch
    # (FEDDB200_FRAG=1), Laplace geometry from the vertex coordinates (FEDDB200_LAPLACE_POINTS=1)
    for env in ({"FEDDB200_FAN": "1"}, {"FEDDB200_STAR": "1"}, {"FEDDB200_TASK": "0"}, {"FEDDB200_FRAG": "1"},
                {"FEDDB200_LAPLACE_POINTS": "1"}, {"FEDDB200_FRAG": "1", "FEDDB200_LAPLACE_POINTS": "1"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "ok" in r.stdout, (env, r.stdout[-500:], r.stderr[-1500:])
