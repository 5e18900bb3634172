"""Generates tests/golden/ref_vectors.npz: outputs of the REFERENCE's own routines (oracle/_ref = FEDDLib's FE_def.hpp
compiled where it lies against mock Trilinos containers) on small deterministic meshes, so that the oracle restatement
stays pinned on machines that have neither /root/reference nor a built oracle/_ref.  Run in the build container:

    python tests/golden/make_ref_vectors.py

Stored per case: CSR (rowptr, column gids, values) of every matrix routine and the load vectors of assemblyRHS."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import ref as R  # noqa: E402
from util import mesh_structured, random_u, stress_coefficient  # noqa: E402

CASES = {"2d_P1": (2, "P1", 3, False), "2d_P2": (2, "P2", 2, True), "3d_P1": (3, "P1", 2, True), "3d_P2": (3, "P2", 2, True)}


def inputs(dim, fe, M, warp):
    conn, coords = mesh_structured(dim, fe, M, warp=warp)
    conn_p, _ = mesh_structured(dim, "P1", M)
    return conn, coords, conn_p, random_u(dim, coords.shape[0])


def p0_inputs(ne):
    """One pseudo-node per element and a permuted element map."""
    return np.arange(ne, dtype=np.int32)[:, None], np.random.default_rng(11).permutation(ne).astype(np.int64)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, (dim, fe, M, warp) in CASES.items():
        conn, coords, conn_p, u = inputs(dim, fe, M, warp)
        for op, kw in (("mass", {}), ("mass_vec", {}), ("laplace", {}), ("laplace_vec", {}), ("linelas", dict(lam=8e6, mu=2e6)),
                       ("advection", dict(u=u)), ("advection_in_u", dict(u=u))):
            rp, ci, v = R.assemble(op, dim, fe, conn, coords, **kw)
            out[f"{name}/{op}/rowptr"], out[f"{name}/{op}/col"], out[f"{name}/{op}/val"] = rp, ci, v
        rp, ci, v = R.assemble_stress(dim, fe, conn, coords, stress_coefficient)
        out[f"{name}/stress/rowptr"], out[f"{name}/stress/col"], out[f"{name}/stress/val"] = rp, ci, v
        if fe == "P1":  # assemblyBDStabilization is P1 only (FE_def.hpp:2156)
            rp, ci, v = R.assemble("bdstab", dim, fe, conn, coords)
            out[f"{name}/bdstab/rowptr"], out[f"{name}/bdstab/col"], out[f"{name}/bdstab/val"] = rp, ci, v
        (B, BT) = R.assemble("div", dim, fe, conn, coords, fe2="P1", conn2=conn_p)
        for tag, (rp, ci, v) in (("B", B), ("BT", BT)):
            out[f"{name}/div{tag}/rowptr"], out[f"{name}/div{tag}/col"], out[f"{name}/div{tag}/val"] = rp, ci, v
        if dim == 2:  # P0 pressure: rows of B / columns of B^T on the element map (FE_def.hpp:1954-1957; 2D only, see fo_phi)
            conn0, egid = p0_inputs(conn.shape[0])
            (B, BT) = R.assemble("div", dim, fe, conn, coords, fe2="P0", conn2=conn0, gid2=egid)
            for tag, (rp, ci, v) in (("B", B), ("BT", BT)):
                out[f"{name}/divP0{tag}/rowptr"], out[f"{name}/divP0{tag}/col"], out[f"{name}/divP0{tag}/val"] = rp, ci, v
        f = np.array([1.5, -2.0, 0.25])[:dim]
        out[f"{name}/rhs_scalar"] = R.assemble_rhs(dim, fe, conn, coords, f, 1, False)
        out[f"{name}/rhs_vector"] = R.assemble_rhs(dim, fe, conn, coords, f, 1, True)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")
