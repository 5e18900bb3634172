"""ORACLE mesh inputs (test infrastructure, not product code).

numpy restatement of the reference's structured generators and P2 construction -- the
*input spec* of the synthetic benchmarks:

  feddlib/core/Mesh/MeshStructured_def.hpp:283-619   buildMesh2D  (P1 :356-463, P2 :476-609)
  feddlib/core/Mesh/MeshStructured_def.hpp:622-1009  buildMesh3D  (P1 :703-806, P2 :808-994)
  feddlib/core/Mesh/MeshUnstructured_def.hpp:129-410 buildP2ofP1MeshEdge (edge-midpoint P2 nodes)
  feddlib/core/Mesh/MeshPartitioner_def.hpp:660-703  buildEdgeListParallel (edges sorted by (min,max))
  feddlib/core/FE/Elements.cpp:242-280               local edge order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) for
                                                     the edge *list*; P2 slots follow FE_def.hpp phi order
                                                     4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3)

Every function returns plain numpy arrays: conn int32 [ne, nloc] (local repeated node ids),
coords float64 [nn, dim], gid int64 [nn] (repeated map: local id -> global id).
"""
from __future__ import annotations

import numpy as np

# corner offsets (dr, ds[, dt]) of each sub-simplex of a cell, local vertex order as in the reference
_TRI_CORNERS = np.array([[[1, 0], [0, 0], [1, 1]],
                         [[0, 1], [0, 0], [1, 1]]], dtype=np.int64)
_TET_CORNERS = np.array([[[1, 0, 0], [0, 0, 0], [1, 0, 1], [1, 1, 1]],
                         [[0, 0, 1], [0, 0, 0], [1, 0, 1], [1, 1, 1]],
                         [[1, 0, 0], [0, 0, 0], [1, 1, 0], [1, 1, 1]],
                         [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]],
                         [[0, 0, 0], [0, 1, 0], [0, 1, 1], [1, 1, 1]],
                         [[0, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1]]], dtype=np.int64)
# P2 mid-node slots: local vertex pairs, in local-node order (slot dim+1, dim+2, ...)
P2_EDGES_2D = np.array([[0, 1], [1, 2], [0, 2]], dtype=np.int64)
P2_EDGES_3D = np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]], dtype=np.int64)


def nloc_of(dim: int, fe: str) -> int:
    if fe == "P1":
        return dim + 1
    if fe == "P2":
        return 6 if dim == 2 else 10
    raise ValueError(f"unsupported FE type {fe}")


def _rank_offsets(rank: int, N: int, dim: int):
    ox = rank % N
    oy = (rank % (N * N)) // N
    oz = (rank % (N * N * N)) // (N * N) if dim == 3 else 0
    return ox, oy, oz


def structured(dim: int, fe: str, N: int, M: int, rank: int = 0, length: float = 1.0):
    """One rank's sub-square / sub-cube of the built-in generator (N^dim ranks, H/h = M)."""
    if M < 1:
        raise ValueError("H/h is too small")
    if fe not in ("P1", "P2"):
        raise ValueError("Wrong FE-Type, either P1 or P2")
    h = length / (M * N)
    H = length / N
    ox, oy, oz = _rank_offsets(rank, N, dim)
    off = np.array([ox, oy, oz][:dim], dtype=np.float64)
    half = 2 if fe == "P2" else 1
    n1 = half * M + 1                        # points per direction on this rank
    nglob = N * n1 - (N - 1)                 # points per direction globally
    step = h / 2.0 if fe == "P2" else h
    eps = np.finfo(np.float64).eps * (100.0 if dim == 2 else 1.0)

    ax = np.arange(n1, dtype=np.int64)
    if dim == 2:
        s, r = np.meshgrid(ax, ax, indexing="ij")
        grid = [r.ravel(), s.ravel()]
    else:
        t, s, r = np.meshgrid(ax, ax, ax, indexing="ij")
        grid = [r.ravel(), s.ravel(), t.ravel()]
    coords = np.empty((grid[0].size, dim), dtype=np.float64)
    for d in range(dim):
        x = grid[d].astype(np.float64) * step + off[d] * H
        x[(x < eps) & (x > -eps)] = 0.0
        coords[:, d] = x
    gid = np.zeros(grid[0].size, dtype=np.int64)
    stride = 1
    for d in range(dim):
        gid += (grid[d] + int(off[d]) * (n1 - 1)) * stride
        stride *= nglob

    # elements: cell-major (t, s, r), sub-simplices in source order
    cax = np.arange(M, dtype=np.int64)
    corners = _TRI_CORNERS if dim == 2 else _TET_CORNERS
    edges = P2_EDGES_2D if dim == 2 else P2_EDGES_3D
    if dim == 2:
        cs, cr = np.meshgrid(cax, cax, indexing="ij")
        cell = np.stack([cr.ravel(), cs.ravel()], axis=1)
    else:
        ct, cs, cr = np.meshgrid(cax, cax, cax, indexing="ij")
        cell = np.stack([cr.ravel(), cs.ravel(), ct.ravel()], axis=1)
    # vertex positions on the (half-)grid: [ncell, nsub, nvert, dim]
    vpos = half * (cell[:, None, None, :] + corners[None, :, :, :])
    strides = np.array([n1 ** d for d in range(dim)], dtype=np.int64)
    vert = (vpos * strides).sum(-1)
    if fe == "P1":
        conn = vert
    else:
        mpos = (vpos[:, :, edges[:, 0], :] + vpos[:, :, edges[:, 1], :]) // 2
        conn = np.concatenate([vert, (mpos * strides).sum(-1)], axis=2)
    conn = conn.reshape(-1, conn.shape[-1]).astype(np.int32)
    return conn, coords, gid


def structured_global(dim: int, fe: str, N: int, M: int):
    """All N^dim ranks of the structured mesh: list of (conn, coords, gid) per rank."""
    return [structured(dim, fe, N, M, rank) for rank in range(N ** dim)]


def p1_of_p2(conn_p2: np.ndarray, dim: int):
    """Vertex (P1) connectivity sharing the element index with a P2 mesh; P1 nodes keep the
    P2 mesh's local ids (they are a subset), which is how `Domain::buildP2ofP1Domain` relates
    the pressure and velocity domains (P1 nodes come first in the P2 repeated map)."""
    return np.ascontiguousarray(conn_p2[:, : dim + 1])


def read_inria_mesh(path: str):
    """INRIA .mesh reader (vertices + top-dimensional simplices), 0-based ids."""
    with open(path) as f:
        tok = f.read().split()
    dim = None
    verts = None
    cells = {}
    i = 0
    while i < len(tok):
        t = tok[i]
        if t == "Dimension":
            dim = int(tok[i + 1]); i += 2
        elif t == "Vertices":
            n = int(tok[i + 1]); i += 2
            a = np.array(tok[i:i + n * (dim + 1)], dtype=np.float64).reshape(n, dim + 1)
            verts = np.ascontiguousarray(a[:, :dim]); i += n * (dim + 1)
        elif t in ("Edges", "Triangles", "Tetrahedra"):
            k = {"Edges": 2, "Triangles": 3, "Tetrahedra": 4}[t]
            n = int(tok[i + 1]); i += 2
            a = np.array(tok[i:i + n * (k + 1)], dtype=np.int64).reshape(n, k + 1)
            cells[t] = a[:, :k] - 1; i += n * (k + 1)
        else:
            i += 1
    top = cells["Tetrahedra"] if dim == 3 else cells["Triangles"]
    return dim, verts, top.astype(np.int32)


def p2_of_p1(conn_p1: np.ndarray, coords_p1: np.ndarray):
    """Single-rank P2-of-P1: one node per mesh edge, edges numbered in lexicographic
    (min id, max id) order, node id = n_P1 + edge index, coordinates = edge midpoint."""
    dim = coords_p1.shape[1]
    edges = P2_EDGES_2D if dim == 2 else P2_EDGES_3D
    a = conn_p1[:, edges[:, 0]].astype(np.int64)
    b = conn_p1[:, edges[:, 1]].astype(np.int64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nv = coords_p1.shape[0]
    key = lo * nv + hi
    ukey, inv = np.unique(key.ravel(), return_inverse=True)
    mids = nv + inv.reshape(key.shape)
    conn = np.concatenate([conn_p1.astype(np.int64), mids], axis=1).astype(np.int32)
    elo, ehi = ukey // nv, ukey % nv
    coords = np.concatenate([coords_p1, (coords_p1[elo] + coords_p1[ehi]) / 2.0], axis=0)
    gid = np.arange(coords.shape[0], dtype=np.int64)
    return conn, coords, gid


def refine_regular(conn_p1: np.ndarray, coords: np.ndarray):
    """Uniform (red) refinement: each triangle -> 4, each tetrahedron -> 8 (Bey's rule)."""
    dim = coords.shape[1]
    conn2, coords2, _ = p2_of_p1(conn_p1, coords)
    c = conn2.astype(np.int64)
    if dim == 2:
        sub = [(0, 3, 5), (3, 1, 4), (5, 4, 2), (3, 4, 5)]
    else:
        # P2 slots: 4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3)
        sub = [(0, 4, 6, 7), (4, 1, 5, 8), (6, 5, 2, 9), (7, 8, 9, 3),
               (4, 6, 7, 8), (4, 6, 5, 8), (6, 7, 8, 9), (6, 5, 9, 8)]
    new = np.stack([c[:, list(s)] for s in sub], axis=1).reshape(-1, dim + 1)
    return new.astype(np.int32), coords2


def lowest_rank_owner(parts):
    """Documented standalone ownership rule (the reference delegates this to Tpetra's
    directory, Map_def.hpp:194-199): the lowest rank holding a node owns it.
    parts: list of gid arrays.  Returns list of owner-rank arrays (int32) per rank."""
    nglob = int(max(g.max() for g in parts)) + 1
    owner = np.full(nglob, np.iinfo(np.int32).max, dtype=np.int64)
    for r, g in enumerate(parts):
        np.minimum.at(owner, g, r)
    return [owner[g].astype(np.int32) for g in parts]
