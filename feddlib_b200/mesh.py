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


# ------------------------------------------------------------------------------------------------------
# Unstructured meshes: INRIA .mesh files, P2-of-P1, regular refinement (BASELINE.json config 4: the refined
# DFG3DCylinder mesh of feddlib/problems/tests/steadyNavierStokes)
# ------------------------------------------------------------------------------------------------------
_SECTION_WIDTH = {"Vertices": None, "Edges": 2, "Triangles": 3, "Quadrilaterals": 4, "Tetrahedra": 4, "Hexahedra": 8}


def read_mesh(path: str):
    """INRIA/medit ASCII `.mesh` file -> (dim, vertices float64[nv,dim], cells int32[ne,dim+1], cell_flags,
    vertex_flags).  The reference reads the same sections (`MeshFileReader`, driven by
    MeshUnstructured::readMeshSize / readMeshEntity, feddlib/core/Mesh/MeshUnstructured_def.hpp:1010-1100):
    every entity line ends with its flag, ids are 1-based in the file and 0-based here; the top-dimensional
    simplices (Tetrahedra in 3D, Triangles in 2D) are the elements."""
    with open(path) as f:
        words = f.read().split()
    dim, at, sections = None, 0, {}
    while at < len(words):
        w = words[at]
        if w == "Dimension":
            dim = int(words[at + 1])
            at += 2
        elif w in _SECTION_WIDTH:
            if dim is None:
                raise ValueError("Dimension must precede the entity sections of a .mesh file")
            count = int(words[at + 1])
            width = (dim if w == "Vertices" else _SECTION_WIDTH[w]) + 1
            body = words[at + 2: at + 2 + count * width]
            if len(body) != count * width:
                raise ValueError(f"truncated section {w}")
            sections[w] = np.asarray(body, dtype=np.float64 if w == "Vertices" else np.int64).reshape(count, width)
            at += 2 + count * width
        elif w == "End":
            break
        else:
            at += 1
    if dim not in (2, 3) or "Vertices" not in sections:
        raise ValueError("not a 2D/3D .mesh file")
    top = "Tetrahedra" if dim == 3 else "Triangles"
    if top not in sections:
        raise ValueError(f"no {top} section")
    v, c = sections["Vertices"], sections[top]
    return (dim, np.ascontiguousarray(v[:, :dim]), np.ascontiguousarray(c[:, :dim + 1] - 1).astype(np.int32),
            c[:, dim + 1].astype(np.int32), v[:, dim].astype(np.int32))


def mesh_edges(conn_p1: np.ndarray):
    """Edges of a simplicial mesh in the reference's order: node pairs sorted (low, high), the list sorted
    lexicographically and made unique (MeshPartitioner_def.hpp:545-560 + EdgeElements::sortUniqueAndSetGlobalIDs,
    feddlib/core/FE/EdgeElements.cpp:105-160).  Returns (edges int64[nedge,2], edge_of int64[ne, nedges_loc]) with the
    local edge order (0,1),(1,2),(0,2)[,(0,3),(1,3),(2,3)] of the P2 element (MeshUnstructured_def.hpp:745-773)."""
    nv = conn_p1.shape[1]
    loc = np.array(_MID[nv - 1])
    a, b = conn_p1[:, loc[:, 0]].astype(np.int64), conn_p1[:, loc[:, 1]].astype(np.int64)
    lo, hi = np.minimum(a, b).ravel(), np.maximum(a, b).ravel()
    order = np.lexsort((hi, lo))
    slo, shi = lo[order], hi[order]
    first = np.ones(order.size, dtype=bool)
    first[1:] = (slo[1:] != slo[:-1]) | (shi[1:] != shi[:-1])
    eid_sorted = np.cumsum(first) - 1
    edge_of = np.empty(order.size, dtype=np.int64)
    edge_of[order] = eid_sorted
    return np.stack([slo[first], shi[first]], axis=1), edge_of.reshape(a.shape)


def build_p2_of_p1(conn_p1: np.ndarray, coords_p1: np.ndarray):
    """Domain::buildP2ofP1Domain on one rank (MeshUnstructured::buildP2ofP1MeshEdge, MeshUnstructured_def.hpp:129-410):
    one new node per mesh edge at its midpoint, numbered n_P1 + edge index (P1 nodes keep their ids), placed at the
    P2 slot of its local edge.  Returns (conn int32[ne, 6|10], coords)."""
    edges, edge_of = mesh_edges(conn_p1)
    n1 = coords_p1.shape[0]
    conn = np.concatenate([conn_p1.astype(np.int64), n1 + edge_of], axis=1).astype(np.int32)
    mids = 0.5 * (coords_p1[edges[:, 0]] + coords_p1[edges[:, 1]])
    return conn, np.concatenate([coords_p1, mids], axis=0)


def refine_regular(conn_p1: np.ndarray, coords_p1: np.ndarray, levels: int = 1):
    """Uniform (red) refinement of a simplicial mesh, `levels` times: every triangle -> 4, every tetrahedron -> 8
    (corner children + the inner octahedron cut along the diagonal mid(0,2)-mid(1,3), Bey's rule), new vertices at the
    edge midpoints, numbered after the old ones in edge order.  Children are stored parent-major."""
    dim = coords_p1.shape[1]
    kids = (((0, 3, 5), (3, 1, 4), (5, 4, 2), (3, 4, 5)) if dim == 2 else
            ((0, 4, 6, 7), (4, 1, 5, 8), (6, 5, 2, 9), (7, 8, 9, 3), (4, 6, 7, 8), (4, 5, 6, 8), (6, 7, 8, 9), (6, 5, 9, 8)))
    conn, coords = conn_p1, coords_p1
    for _ in range(levels):
        c2, coords = build_p2_of_p1(conn, coords)
        conn = np.ascontiguousarray(c2[:, np.array(kids)].reshape(-1, dim + 1))
    return conn, coords


def element_volumes(conn: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """Signed volumes (areas) of the simplices (vertex nodes only)."""
    dim = coords.shape[1]
    x0 = coords[conn[:, 0]]
    B = np.stack([coords[conn[:, j + 1]] - x0 for j in range(dim)], axis=-1)
    return np.linalg.det(B) / (2.0 if dim == 2 else 6.0)
