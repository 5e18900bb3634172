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
