"""Generates tests/golden/dfg3d_1k.npz from the reference's mesh file (run in the build container,
where /root/reference exists; the GPU box only sees the committed .npz).

Source: /root/reference/meshes/DFG3DCylinder_1k.mesh (INRIA .mesh; 1 499 vertices, 5 476 tetrahedra),
the coarse version of the mesh BASELINE.json's Navier-Stokes config names.  Stored: P1 connectivity
(0-based) and vertex coordinates, nothing else.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import mesh as OM  # noqa: E402

src = "/root/reference/meshes/DFG3DCylinder_1k.mesh"
dim, verts, tets = OM.read_inria_mesh(src)
assert dim == 3
out = os.path.join(os.path.dirname(__file__), "dfg3d_1k.npz")
np.savez_compressed(out, conn=tets.astype(np.int32), coords=verts)
print(out, tets.shape, verts.shape)
