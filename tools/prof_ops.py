#!/usr/bin/env python
"""One call of one operator on the structured P2-P1 cube (profiling aid: every launch of the call is one ncu record).
    python tools/prof_ops.py <op: nsj|adv|advu|lap|lapvec|elas|div> [M=50] [calls=1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, Context, Mesh, Pattern
from feddlib_b200 import mesh as PM
from feddlib_b200.engine import assemble_div_divT_d

op = sys.argv[1] if len(sys.argv) > 1 else "nsj"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 50
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dim = 3
ctx = Context(0)
conn, coords, gid = PM.build_structured(dim, "P2", 1, M)
mv = Mesh(ctx, dim, conn, coords)
pat = Pattern(ctx, mv)
u = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, dim * coords.shape[0])).cuda()
mode = {"lap": BLOCK_SCALAR, "lapvec": BLOCK_DIAG, "adv": BLOCK_DIAG}.get(op, BLOCK_FULL)
v = ctx.empty_values(pat.nnz(1 if op == "lap" else dim, 1 if op == "lap" else dim, mode))
fn = {"nsj": lambda: pat.assemble_ns_jacobian_d(v, u, 1.0, 1e-3, True), "adv": lambda: pat.assemble_advection_d(v, u),
      "advu": lambda: pat.assemble_advection_in_u_d(v, u), "lap": lambda: pat.assemble_laplace_d(v), "lapvec": lambda: pat.assemble_laplace_d(v, True),
      "elas": lambda: pat.assemble_linelas_d(v, 8e6, 2e6)}[op]
for _ in range(calls):
    fn()
torch.cuda.synchronize()
print(op, "done", float(v[:1000].sum()))
