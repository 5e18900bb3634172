"""Shared helpers of the parity tests: meshes, oracle CSR, comparisons."""
from __future__ import annotations

import os

import numpy as np

from oracle import mesh as OM
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12  # relative Frobenius error allowed by BASELINE.json's north_star


def mesh_structured(dim, fe, M, warp=False, shuffle=False, seed=0):
    conn, coords, gid = OM.structured(dim, fe, 1, M)
    if warp:
        x = coords.copy()
        for d in range(dim):
            o = coords[:, (d + 1) % dim]
            x[:, d] = coords[:, d] + 0.04 * np.sin(np.pi * coords[:, d]) * np.cos(np.pi * o)
        coords = x
    if shuffle:  # random renumbering of nodes and elements: an "unstructured" ordering
        rng = np.random.default_rng(seed)
        perm = rng.permutation(coords.shape[0])          # new id of old node k is perm[k]
        inv = np.empty_like(perm); inv[perm] = np.arange(perm.size)
        coords = coords[inv]
        conn = perm[conn].astype(np.int32)
        conn = conn[rng.permutation(conn.shape[0])]
    return np.ascontiguousarray(conn), np.ascontiguousarray(coords)


def mesh_dfg(fe="P2"):
    z = np.load(os.path.join(GOLDEN, "dfg3d_1k.npz"))
    conn, coords = z["conn"], z["coords"]
    if fe == "P2":
        conn, coords, _ = OM.p2_of_p1(conn, coords)
    return np.ascontiguousarray(conn.astype(np.int32)), np.ascontiguousarray(coords)


def oracle_csr(op, dim, fe, conn, coords, u=None, lam=None, mu=None, fe2=None, conn2=None, func=None):
    """CSR (rowptr, colind, values) of the oracle for a single rank with gid == lid."""
    nn = coords.shape[0]
    gid = np.arange(nn, dtype=np.int64)
    if op == "laplace":
        A = O.Matrix(nn); O.assembly_laplace(dim, fe, conn, coords, gid, A)
    elif op == "mass":
        A = O.Matrix(nn); O.assembly_mass(dim, fe, conn, coords, gid, A, False)
    elif op == "stress":
        A = O.Matrix(dim * nn, 64); O.assembly_stress(dim, fe, conn, coords, gid, func, A)
    elif op == "bdstab":
        A = O.Matrix(nn); O.assembly_bdstab(dim, fe, conn, coords, gid, A)
    elif op == "mass_vec":
        A = O.Matrix(dim * nn); O.assembly_mass(dim, fe, conn, coords, gid, A, True)
    elif op == "laplace_vec":
        A = O.Matrix(dim * nn); O.assembly_laplace_vecfield(dim, fe, conn, coords, gid, A)
    elif op == "linelas":
        A = O.Matrix(dim * nn, 64); O.assembly_linelas(dim, fe, conn, coords, gid, lam, mu, A)
    elif op == "advection":
        A = O.Matrix(dim * nn); O.assembly_advection(dim, fe, conn, coords, gid, u, A)
    elif op == "advection_in_u":
        A = O.Matrix(dim * nn, 64); O.assembly_advection_in_u(dim, fe, conn, coords, gid, u, A)
    elif op in ("div", "divT"):
        n2 = int(conn2.max()) + 1
        gid2 = np.arange(n2, dtype=np.int64)
        B, BT = O.Matrix(n2, 64), O.Matrix(dim * nn)
        O.assembly_div_divT(dim, fe, fe2, conn, coords, gid, conn2, gid2, B, BT)
        A = B if op == "div" else BT
    else:
        raise ValueError(op)
    return A.csr()


def stress_coefficient(x, parameters=None):
    """A smooth, non-constant CoeffFunc_Type for the assemblyStress tests."""
    return 1.0 + 0.5 * x[0] + 0.25 * x[1] * x[1] + (0.125 * x[2] if len(x) > 2 else 0.0)


def rel_frobenius(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def random_u(dim, nn, seed=1234):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, dim * nn)
