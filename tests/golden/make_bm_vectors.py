"""Generates tests/golden/bm_vectors.npz: the reference's own BlockMatrix::merge (oracle/_ref/libfedd_ref_bm.so, compiled from
/root/reference where it lies) on the block system of tests/test_csrops_vs_ref.py (seed 0).  Run in the build container."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_bm as RB  # noqa: E402
import test_csrops_vs_ref as T  # noqa: E402

n, row_gids, blocks, _ = T.system(0, True)
rp, rg, cg, va = RB.merge(row_gids, blocks)
path = os.path.join(os.path.dirname(__file__), "bm_vectors.npz")
np.savez_compressed(path, rowptr=rp, rowgid=rg, colgid=cg, values=va)
print(path, rp.shape, va.shape)
