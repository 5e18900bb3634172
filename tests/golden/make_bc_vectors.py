"""Generates tests/golden/bc_vectors.npz: outputs of the reference's own BCBuilder::setSystem / setRHS (oracle/_ref/libfedd_ref_bc.so,
compiled from /root/reference where it lies) on the block system of tests/test_bc_vs_ref.py (seed 0).  Run in the build container."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_bc as RB  # noqa: E402
import test_bc_vs_ref as T  # noqa: E402

coords, gid, flags, dofs, blocks = T.system(0)
out = {f"sys_{i}{j}": v for (i, j), v in RB.set_system(3, flags, gid, T.BCS, dofs, blocks).items()}
rng = np.random.default_rng(10)
rhs = [rng.uniform(-1, 1, flags.size * d) for d in dofs]
for b, bcs, f in T.rhs_cases():
    out[f"rhs_{b}"] = RB.set_rhs(3, flags, coords, gid, bcs, f, np.array([1.5, -0.25]), dofs, rhs, t=0.75)[b]
path = os.path.join(os.path.dirname(__file__), "bc_vectors.npz")
np.savez_compressed(path, **out)
print(path, {k: v.shape for k, v in out.items()})
