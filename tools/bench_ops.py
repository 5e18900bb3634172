#!/usr/bin/env python
"""Times every operator of the hot path on the structured P2 cube (tuning aid; all sizes on one GPU).
    python tools/bench_ops.py [M] [modes...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, Context, Mesh, Pattern
from feddlib_b200 import mesh as PM
from feddlib_b200.engine import assemble_div_divT_d

M = int(sys.argv[1]) if len(sys.argv) > 1 else 40
modes = sys.argv[2:] or ["gather", "atomic", "coloured"]
dim = 3
ctx = Context(0)
conn, coords, gid = PM.build_structured(dim, "P2", 1, M)
verts = np.unique(conn[:, :4]); lid = -np.ones(coords.shape[0], dtype=np.int64); lid[verts] = np.arange(verts.size)
conn1 = lid[conn[:, :4]].astype(np.int32); coords1 = coords[verts]
mv, mp = Mesh(ctx, dim, conn, coords), Mesh(ctx, dim, conn1, coords1)
pat, patB, patBT = Pattern(ctx, mv), Pattern(ctx, mp, mv), Pattern(ctx, mv, mp)
ne = conn.shape[0]
u = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, dim * coords.shape[0])).cuda()
vs, vd, vf = (ctx.empty_values(pat.nnz(*a)) for a in ((1, 1, BLOCK_SCALAR), (dim, dim, BLOCK_DIAG), (dim, dim, BLOCK_FULL)))
vB, vBT = ctx.empty_values(patB.nnz(1, dim, BLOCK_FULL)), ctx.empty_values(patBT.nnz(dim, 1, BLOCK_FULL))
ops = {
    "laplace": (lambda: pat.assemble_laplace_d(vs), vs.numel()),
    "laplace_vec": (lambda: pat.assemble_laplace_d(vd, True), vd.numel()),
    "linelas": (lambda: pat.assemble_linelas_d(vf, 8e6, 2e6), vf.numel()),
    "advection N": (lambda: pat.assemble_advection_d(vd, u), vd.numel()),
    "advection_in_u W": (lambda: pat.assemble_advection_in_u_d(vf, u), vf.numel()),
    "div B+BT": (lambda: assemble_div_divT_d(ctx, patB, patBT, vB, vBT), vB.numel() + vBT.numel()),
    "ns_jacobian (A+N+W)": (lambda: pat.assemble_ns_jacobian_d(vf, u, 1.0, 1e-3, True), vf.numel()),
}
print(f"M={M}: {ne} P2 tets")
for mode in modes:
    ctx.set_scatter_mode(mode)
    for name, (fn, nnz) in ops.items():
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{mode:9s} {name:22s} {ms:9.3f} ms  {ne / ms / 1e3:9.1f} Melem/s  values-only roofline {nnz * 8 / ms / 1e6 / 6549.1:6.3f}")
