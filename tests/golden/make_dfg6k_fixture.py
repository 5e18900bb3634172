"""Generates tests/golden/dfg3d_6k.npz from the reference's mesh file (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz).

Source: /root/reference/meshes/DFG3DCylinder_6k.mesh (INRIA .mesh; 6 721 vertices, 27 618 tetrahedra), the mesh
BASELINE.json's Navier-Stokes configuration names (config 4).  Read with the PRODUCT's reader
(feddlib_b200.mesh.read_mesh) and cross-checked against the oracle's; stored: P1 connectivity (0-based, int32),
vertex coordinates, element and vertex flags.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT)
from feddlib_b200 import mesh as PM  # noqa: E402
from oracle import mesh as OM  # noqa: E402

src = "/root/reference/meshes/DFG3DCylinder_6k.mesh"
dim, verts, tets, eflags, vflags = PM.read_mesh(src)
d2, v2, t2 = OM.read_inria_mesh(src)
assert dim == d2 == 3 and np.array_equal(verts, v2) and np.array_equal(tets, t2)
out = os.path.join(os.path.dirname(__file__), "dfg3d_6k.npz")
np.savez_compressed(out, conn=tets.astype(np.int32), coords=verts, elem_flags=eflags.astype(np.int8), vert_flags=vflags.astype(np.int8))
print(out, tets.shape, verts.shape, os.path.getsize(out))
