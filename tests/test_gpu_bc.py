"""Dirichlet row replacement on resident values: feddb200_set_dirichlet_rows_d against a numpy
restatement of BCBuilder::setLocalRowOne / setLocalRowZero (core/General/BCBuilder_def.hpp:653-709)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from util import mesh_structured

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,fe,M", [(2, "P2", 5), (3, "P2", 3), (3, "P1", 4)])
@pytest.mark.parametrize("layout", ["scalar", "diag", "full"])
@pytest.mark.parametrize("diagonal_block", [True, False])
def test_dirichlet_rows(engine_ctx, dim, fe, M, layout, diagonal_block):
    from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, Mesh, Pattern
    conn, coords = mesh_structured(dim, fe, M, warp=True)
    mesh = Mesh(engine_ctx, dim, conn, coords)
    pat = Pattern(engine_ctx, mesh)
    rd = cd = 1 if layout == "scalar" else dim
    mode = {"scalar": BLOCK_SCALAR, "diag": BLOCK_DIAG, "full": BLOCK_FULL}[layout]
    rp, ci = pat.expand(rd, cd, mode)
    nn = coords.shape[0]
    rng = np.random.default_rng(5)
    vals = rng.uniform(-1, 1, rp[-1])
    # boundary nodes: x == 0 -> "Dirichlet" (all dofs); y == 1 -> "Dirichlet_X_Z"-like partial masks; rest free
    mask = np.zeros(nn, dtype=np.uint8)
    mask[np.abs(coords[:, 0]) < 1e-12] = (1 << rd) - 1
    part = np.abs(coords[:, 1] - 1.0) < 1e-12
    mask[part] |= 0b101 & ((1 << rd) - 1)
    dofmask = np.zeros(nn * rd, dtype=bool)
    for a in range(rd):
        dofmask[a::rd] = (mask >> a) & 1
    want = O.set_dirichlet_rows(rp, ci, vals, np.arange(nn * rd), dofmask, diagonal_block)
    v_d = torch.from_numpy(vals).cuda()
    pat.set_dirichlet_rows_d(v_d, torch.from_numpy(mask).cuda(), rd, cd, mode, diagonal_block)
    got = v_d.cpu().numpy()
    assert np.array_equal(got, want)
    assert (got != vals).any()


@pytest.mark.parametrize("dofs", [1, 3])
def test_dirichlet_rhs(engine_ctx, dofs):
    """feddb200_set_dirichlet_rhs_d against the restatement of BCBuilder::setRHS (pinned by the reference's own routine in
    tests/test_bc_vs_ref.py): flags by position, a table of Dirichlet types, a boundary function of the coordinates."""
    from feddlib_b200 import Mesh, Pattern
    conn, coords = mesh_structured(3, "P2", 3, warp=True)
    nn = coords.shape[0]
    pat = Pattern(engine_ctx, Mesh(engine_ctx, 3, conn, coords))
    flags = np.zeros(nn, dtype=np.int32)
    flags[np.abs(coords[:, 0]) < 1e-12] = 1
    flags[np.abs(coords[:, 1] - 1) < 1e-12] = 2
    flags[np.abs(coords[:, 2]) < 1e-12] = 3
    bcs = [(1, 0, "Dirichlet", dofs), (2, 0, "Dirichlet_X_Z" if dofs == 3 else "Dirichlet", dofs), (3, 0, "Dirichlet_Y" if dofs == 3 else "Neumann", dofs)]
    func = (lambda x, t, par: [par[0] * x[0] + t, par[1] - x[1] * x[2], 2.0 + x[2]][:dofs])
    par, t = np.array([1.5, -0.25]), 0.75
    rhs = np.random.default_rng(3).uniform(-1, 1, nn * dofs)
    want = O.set_dirichlet_rhs(rhs, flags, coords, bcs, 0, dofs, func, par, t)
    mask, _ = O.dirichlet_dof_masks(flags, bcs, 0, dofs)
    bc_values = np.array([func(x, t, par) for x in coords]).reshape(-1)      # the host glue evaluates the user's function per node
    r_d = torch.from_numpy(rhs).cuda()
    pat.set_dirichlet_rhs_d(r_d, torch.from_numpy(mask).cuda(), torch.from_numpy(bc_values).cuda(), dofs)
    got = r_d.cpu().numpy()
    assert np.array_equal(got, want) and (got != rhs).any()
