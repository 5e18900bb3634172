"""Synthetic structured meshes for benchmarks and examples (host side, numpy).

Input spec = FEDDLib's built-in generators `MeshStructured::buildMesh2D/3D`
(reference: feddlib/core/Mesh/MeshStructured_def.hpp:283-619, 622-1009):
unit square / cube, N^dim sub-domains (one per rank), H/h = M cells per direction and rank,
every cell split into 2 triangles / 6 Kuhn tetrahedra, P2 nodes on the half grid, node order
vertices first then edge mid-points (0,1),(1,2),(0,2)[,(0,3),(1,3),(2,3)].
"""
from __future__ import annotations

import numpy as np

# vertices of the sub-simplices of a unit cell as corner codes: bit0 = +x, bit1 = +y, bit2 = +z
_TRI = ((1, 0, 3), (2, 0, 3))
_TET = ((1, 0, 5, 7), (4, 0, 5, 7), (1, 0, 3, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 6, 7))
_MID = {2: ((0, 1), (1, 2), (0, 2)), 3: ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))}


def nloc(dim: int, fe: str) -> int:
    try:
        return {("P1", 2): 3, ("P2", 2): 6, ("P1", 3): 4, ("P2", 3): 10}[(fe, dim)]
    except KeyError:
        raise ValueError("Wrong FE-Type, either P1 or P2.") from None


def build_structured(dim: int, fe: str, N: int, M: int, rank: int = 0, length: float = 1.0):
    """Return (conn int32[ne,nloc], coords float64[nn,dim], gid int64[nn]) of `rank`'s sub-domain."""
    if M < 1:
        raise ValueError("H/h is to small.")
    nl = nloc(dim, fe)
    k = 2 if fe == "P2" else 1                 # grid refinement of the node lattice
    n = k * M + 1                              # lattice points per direction (this rank)
    ng = N * n - (N - 1)                       # lattice points per direction (global)
    h, H = length / (M * N), length / N
    offs = [rank % N, (rank % (N * N)) // N, (rank % (N ** 3)) // (N * N)][:dim]
    eps = np.finfo(np.float64).eps * (100.0 if dim == 2 else 1.0)

    lat = np.indices((n,) * dim, dtype=np.int64)[::-1].reshape(dim, -1)   # row d = index along axis d, x fastest
    coords = np.empty((lat.shape[1], dim))
    gid = np.zeros(lat.shape[1], dtype=np.int64)
    for d in range(dim):
        x = lat[d] * (h / 2.0 if k == 2 else h) + offs[d] * H
        x[np.abs(x) < eps] = 0.0
        coords[:, d] = x
        gid += (lat[d] + offs[d] * (n - 1)) * ng ** d

    cells = np.indices((M,) * dim, dtype=np.int64)[::-1].reshape(dim, -1).T  # [ncell, dim], x fastest
    subs = np.array(_TRI if dim == 2 else _TET)
    corner = np.stack([(subs >> d) & 1 for d in range(dim)], axis=-1)         # [nsub, nv, dim]
    vpos = k * (cells[:, None, None, :] + corner[None])                       # lattice coordinates of the vertices
    w = n ** np.arange(dim)
    conn = (vpos * w).sum(-1)
    if k == 2:
        e = np.array(_MID[dim])
        mid = (vpos[:, :, e[:, 0]] + vpos[:, :, e[:, 1]]) // 2
        conn = np.concatenate([conn, (mid * w).sum(-1)], axis=-1)
    assert conn.shape[-1] == nl
    return conn.reshape(-1, nl).astype(np.int32), coords, gid


def vertex_connectivity(conn_p2: np.ndarray, dim: int) -> np.ndarray:
    """P1 (pressure) connectivity on the node numbering of a P2 (velocity) mesh."""
    return np.ascontiguousarray(conn_p2[:, : dim + 1])


def warp_coords(coords: np.ndarray, amp: float = 0.08) -> np.ndarray:
    """Smooth non-affine deformation of the unit box (keeps elements valid for amp < ~0.1); used to
    turn the structured connectivity into a geometrically non-uniform test mesh."""
    x = coords.copy()
    dim = x.shape[1]
    for d in range(dim):
        o = coords[:, (d + 1) % dim]
        x[:, d] = coords[:, d] + amp * np.sin(np.pi * coords[:, d]) * np.cos(np.pi * o) * 0.5
    return x


def build_structured_box(dim: int, fe: str, dims, M: int, rank: int):
    """Weak-scaling workload: a box of dims[0] x dims[1] (x dims[2]) unit sub-cubes, one per rank, each
    meshed like build_structured(N=1, M) (same cell split, same local numbering).  Returns
    (conn, coords, gid, owner): gid on the global node lattice, owner = lowest rank holding the node
    (the standalone ownership rule; the reference leaves it to Tpetra's directory, Map_def.hpp:194-199)."""
    dims = tuple(int(d) for d in dims)[:dim] + (1,) * (3 - dim)
    nranks = dims[0] * dims[1] * dims[2]
    if not 0 <= rank < nranks:
        raise ValueError("rank outside the box of sub-domains")
    conn, coords, _ = build_structured(dim, fe, 1, M, 0)
    k = 2 if fe == "P2" else 1
    n = k * M + 1
    off = (rank % dims[0], (rank // dims[0]) % dims[1], rank // (dims[0] * dims[1]))
    ng = [dims[d] * (n - 1) + 1 for d in range(3)]
    lat = np.indices((n,) * dim, dtype=np.int64)[::-1].reshape(dim, -1)
    gid = np.zeros(lat.shape[1], dtype=np.int64)
    owner = np.zeros(lat.shape[1], dtype=np.int64)
    stride_g, stride_r = 1, 1
    coords = coords.copy()
    for d in range(dim):
        X = lat[d] + off[d] * (n - 1)                       # global lattice coordinate
        gid += X * stride_g
        stride_g *= ng[d]
        ob = np.where(X == 0, 0, np.minimum((X - 1) // (n - 1), dims[d] - 1))   # lowest sub-box holding X
        owner += ob * stride_r
        stride_r *= dims[d]
        coords[:, d] += off[d]
    return conn, coords, gid, owner.astype(np.int32)
