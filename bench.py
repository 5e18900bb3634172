#!/usr/bin/env python
"""bench.py -- FE assembly throughput (P2 tet elasticity) on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--M m] [--mode gather|coloured|atomic]

A "step" is one full assembly of the linear-elasticity matrix (FE::assemblyLinElasXDim,
reference feddlib/core/FE/FE_def.hpp:2739-3040) on the built-in structured cube with P2 tetrahedra
(config 3 of BASELINE.json: H/h = 70 -> 2 058 000 tets, 716 110 929 CSR values).  Prints ONE JSON line.

 * value       elements/s, device-resident inputs and outputs, CUDA events, max over ranks
 * e2e         the same metric through the host-buffer path of the C ABI: coordinates H2D from pinned
               memory and the CSR values D2H into pinned memory inside the timed region
 * roofline    algorithmic bytes (conn + vertex coords + values) / device time of the pass
 * cpu_baseline  FEDDLib's own assemblyLinElasXDim loop (oracle/_ref, kind "reference"; falls back to the oracle
                 restatement, kind "port", where oracle/_ref is not built) on a bounded sample, 1 core
 * --impl reference : the same CPU code on all host cores, one element partition per core
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LAM, MU = 8.0e6, 2.0e6  # nu = 0.4, mu = 2e6 (steadyLinElas_Perf/parametersProblem.xml:10-11, LinElas_def.hpp:76-77)
METRIC = "fe_assembly_p2_tet_elasticity_elements_per_s"
UNIT = "elements/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(M):
    """dram__bytes_read.sum + dram__bytes_write.sum of one assembly pass (all its launches) from the committed
    `ncu --set full` capture of this command and this build (profiles/r02_step_traffic.json, written by tools/ncu_step_traffic.py
    from the capture of tools/make_profiles.sh); None for other sizes."""
    p = os.path.join(ROOT, "profiles", "r02_step_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if int(d.get("M", -1)) == int(M):
            return d["dram_bytes_per_step"]
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_kind():
    from oracle import ref as R
    return "reference" if R.available() else "port"


def oracle_elasticity_time(M, repeat=1):
    """Seconds for one CPU assembly of the 6*M^3-tet cube: the reference's own assemblyLinElasXDim loop
    (oracle/_ref: FE_def.hpp compiled against mock Trilinos containers, kind "reference") when it has been
    built, else the oracle restatement (kind "port"); both insert one entry per call and end with the
    per-row sort/merge that stands in for Tpetra's fillComplete."""
    from oracle import mesh as OM
    from oracle import oracle as O
    from oracle import ref as R
    conn, co, gid = OM.structured(3, "P2", 1, M)
    best = 1e30
    for _ in range(repeat):
        t0 = time.perf_counter()
        if R.available():
            R.assemble("linelas", 3, "P2", conn, co, gid, lam=LAM, mu=MU)
        else:
            A = O.Matrix(3 * co.shape[0], 240)            # LinElas_def.hpp:80 capacity hint dim*80
            O.assembly_linelas(3, "P2", conn, co, gid, LAM, MU, A)
            A.fillComplete()
            del A
        best = min(best, time.perf_counter() - t0)
    return best, conn.shape[0]


def ns_block_numbers(ctx, conn, coords, n_vertices, label, peak):
    """Times the (0,0) block (rho*nu*A + rho*N(u) + rho*W(u), fused), N, W and B/B^T on a P2-P1 mesh (velocity connectivity
    conn [ne,10], P1 nodes first in the numbering or anywhere: the pressure mesh is the vertex sub-mesh)."""
    import torch
    from feddlib_b200 import BLOCK_DIAG, BLOCK_FULL, Mesh, Pattern
    from feddlib_b200.engine import assemble_div_divT_d
    dim = 3
    verts = np.unique(conn[:, :4])
    lid = -np.ones(coords.shape[0], dtype=np.int64)
    lid[verts] = np.arange(verts.size)
    t0 = time.perf_counter()
    mv, mp = Mesh(ctx, dim, conn, coords), Mesh(ctx, dim, lid[conn[:, :4]].astype(np.int32), coords[verts])
    pat, patB, patBT = Pattern(ctx, mv), Pattern(ctx, mp, mv), Pattern(ctx, mv, mp)
    ctx.synchronize()
    t_pat = time.perf_counter() - t0
    u = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, dim * coords.shape[0])).cuda()
    vd, vf = ctx.empty_values(pat.nnz(dim, dim, BLOCK_DIAG)), ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
    vB, vBT = ctx.empty_values(patB.nnz(1, dim, BLOCK_FULL)), ctx.empty_values(patBT.nnz(dim, 1, BLOCK_FULL))
    ne = conn.shape[0]
    ops = {"ns_jacobian_00_block": (lambda: pat.assemble_ns_jacobian_d(vf, u, 1.0, 1e-3, True), vf.numel()),
           "advection_N": (lambda: pat.assemble_advection_d(vd, u), vd.numel()),
           "advection_in_u_W": (lambda: pat.assemble_advection_in_u_d(vf, u), vf.numel()),
           "laplace_vec_A": (lambda: pat.assemble_laplace_d(vd, True), vd.numel()),
           "div_B_and_BT": (lambda: assemble_div_divT_d(ctx, patB, patBT, vB, vBT), vB.numel() + vBT.numel())}
    out = {"workload": f"{label}, {ne} tets, {coords.shape[0]} P2 nodes, u ~ U(-1,1) seed 1234, rho=1, nu=1e-3, scatter mode gather",
           "pattern_build_s": t_pat}
    for name, (fn, nnz) in ops.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        needs_u = name in ("ns_jacobian_00_block", "advection_N", "advection_in_u_W")
        alg = ne * 10 * 4 + n_vertices * dim * 8 + nnz * 8 + (dim * coords.shape[0] * 8 if needs_u else 0)
        out[name] = {"ms": ms, "elements_per_s": ne / (ms * 1e-3), "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes": alg}
    del vd, vf, vB, vBT, pat, patB, patBT, mv, mp
    return out


def ns_jacobian_numbers(ctx, M, peak):
    from feddlib_b200 import mesh as PM
    conn, coords, _ = PM.build_structured(3, "P2", 1, M)
    return ns_block_numbers(ctx, conn, coords, (M + 1) ** 3, f"structured P2-P1 cube H/h={M}", peak)


def laplace_numbers(ctx, dim, fe, M, label, peak, steps=10):
    """BASELINE.json configs 1 and 2: scalar Laplace (FE::assemblyLaplace) on the built-in structured mesh; config 1 (P1 square,
    ~1e5 triangles) is the launch-latency case, config 2 (P2 cube, 6M tets) the bandwidth case."""
    import torch
    from feddlib_b200 import BLOCK_SCALAR, Mesh, Pattern
    from feddlib_b200 import mesh as PM
    conn, coords, _ = PM.build_structured(dim, fe, 1, M)
    t0 = time.perf_counter()
    mesh = Mesh(ctx, dim, conn, coords)
    pat = Pattern(ctx, mesh)
    ctx.synchronize()
    t_pat = time.perf_counter() - t0
    nnz = pat.nnz(1, 1, BLOCK_SCALAR)
    v = ctx.empty_values(nnz)
    for _ in range(3):
        pat.assemble_laplace_d(v, False)
    torch.cuda.synchronize()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pat.assemble_laplace_d(v, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        pat.assemble_laplace_d(v, False)
    torch.cuda.synchronize()
    wall_us = (time.perf_counter() - t0) / steps * 1e6
    ne = conn.shape[0]
    alg = ne * conn.shape[1] * 4 + (M + 1) ** dim * dim * 8 + nnz * 8
    out = {"workload": f"{label}: Laplace {fe} on the structured {'square' if dim == 2 else 'cube'} H/h={M}, {ne} elements, nnz {nnz}",
           "ms": ms, "us_per_assembly_wall": wall_us, "elements_per_s": ne / (ms * 1e-3), "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
           "algorithmic_bytes": alg, "launches": (ctx.launches - l0) // steps, "pattern_build_s": t_pat,
           "checksum": float(v[: min(nnz, 1 << 20)].sum().item())}
    del v, pat, mesh
    return out


def config4_numbers(ctx, levels, peak):
    """BASELINE.json config 4: the Navier-Stokes blocks on meshes/DFG3DCylinder_6k.mesh (27 618 tets; committed as
    tests/golden/dfg3d_6k.npz, the reference tree is absent on the GPU box) after `levels` regular refinements
    (levels = 2: 1 767 552 tets), P2 velocity via edge midpoints (buildP2ofP1Domain), P1 pressure."""
    from feddlib_b200 import mesh as PM
    d = np.load(os.path.join(ROOT, "tests", "golden", "dfg3d_6k.npz"))
    c1, x1 = PM.refine_regular(d["conn"], d["coords"], levels)
    conn, coords = PM.build_p2_of_p1(c1, x1)
    return ns_block_numbers(ctx, conn, coords, x1.shape[0],
                            f"config 4: DFG3DCylinder_6k.mesh regular-refined k={levels}, P2-P1", peak)


def parity_check(values, pat, M, dims, rank, plan, n_sample=800):
    """Sampled rows of the FULL-SIZE matrix against the reference's own assembly loop (oracle/_ref; the oracle restatement where it is
    not built): for every sampled owned row node -- interior nodes, nodes on the faces / edges / corners of the global box and on
    the planes between ranks (rows that receive ghost contributions) -- the elements of its star are regenerated from the global
    lattice (any rank's elements), assembled by the checker, and the node's dof rows are compared: pattern exactly, values to 1e-12
    (relative Frobenius norm over all sampled rows).  Checker only: never inside a timed region."""
    from feddlib_b200 import mesh as PM
    from oracle import oracle as O
    from oracle import ref as R
    dim = 3
    dims = tuple(dims) + (1,) * (3 - len(dims))
    n = 2 * M                                                   # lattice intervals per rank and direction
    ng = [dims[d] * n + 1 for d in range(3)]
    if plan is None:
        ugid, colmap, n_owned = None, None, pat.n_owned_rows
    else:
        ugid, colmap, n_owned = plan.unique_gids, plan.colmap_gids, plan.n_owned
    rng = np.random.default_rng(1234 + rank)
    rows_all = np.arange(n_owned, dtype=np.int64)
    g_all = rows_all if ugid is None else ugid
    X = np.stack([g_all % ng[0], (g_all // ng[0]) % ng[1], g_all // (ng[0] * ng[1])], axis=1)
    on_plane = (X % n == 0).sum(axis=1)                          # 0 interior ... 3 corner of a sub-cube
    pick = []
    for k in range(4):
        cand = rows_all[on_plane == k]
        if cand.size:
            pick.append(rng.choice(cand, size=min(cand.size, n_sample // 4), replace=False))
    rows = np.unique(np.concatenate(pick))
    # star meshes from the global lattice: the cells around the node, every Kuhn tetrahedron of them, isolated numbering per star
    tets, mids = np.array(PM._TET), np.array(PM._MID[3])
    conn_l, coord_l, gid_l, base = [], [], [], 0
    row_gid, row_loc = [], []
    h2 = 0.5 / M
    for I in rows:
        x = X[I]
        lo = [max(0, (x[d] - 1) // 2) for d in range(3)]
        hi = [min(dims[d] * M - 1, x[d] // 2) for d in range(3)]
        cells = np.array([[cx, cy, cz] for cz in range(lo[2], hi[2] + 1) for cy in range(lo[1], hi[1] + 1) for cx in range(lo[0], hi[0] + 1)])
        corner = np.stack([(tets >> d) & 1 for d in range(3)], axis=-1)              # [6, 4, 3]
        v = 2 * (cells[:, None, None, :] + corner[None])                             # [nc, 6, 4, 3] lattice coordinates
        m = (v[:, :, mids[:, 0]] + v[:, :, mids[:, 1]]) // 2
        lat = np.concatenate([v, m], axis=2).reshape(-1, 10, 3)                      # [ne, 10, 3]
        lat = lat[(lat == x).all(axis=2).any(axis=1)]                                # tetrahedra that contain the node
        g = lat[..., 0] + ng[0] * (lat[..., 1] + ng[1] * lat[..., 2])
        ug, inv = np.unique(g, return_inverse=True)
        conn_l.append(inv.reshape(-1, 10) + base)
        latu = np.stack([ug % ng[0], (ug // ng[0]) % ng[1], ug // (ng[0] * ng[1])], axis=1)
        coord_l.append(latu * h2)
        gid_l.append(ug)
        row_gid.append(g_all[I]); row_loc.append(base + int(np.searchsorted(ug, g_all[I])))
        base += ug.size
    conn_s = np.concatenate(conn_l).astype(np.int32)
    coords_s = np.concatenate(coord_l).astype(np.float64)
    gid_true = np.concatenate(gid_l)
    if R.available():
        rp, ci, va = R.assemble("linelas", dim, "P2", conn_s, coords_s, lam=LAM, mu=MU)
        kind = "reference"
    else:
        A = O.Matrix(dim * coords_s.shape[0], 240)
        O.assembly_linelas(dim, "P2", conn_s, coords_s, np.arange(coords_s.shape[0]), LAM, MU, A)
        rp, ci, va = A.csr()
        kind = "port"
    # our rows
    nrp, nci = pat.nodes()
    vals_rows = []
    num = den = 0.0
    ok = True
    import torch
    idx = []
    meta = []
    for I, rl in zip(rows, row_loc):
        b0, L = int(nrp[I]), int(nrp[I + 1] - nrp[I])
        cols = nci[b0:b0 + L].astype(np.int64)
        cg = cols if colmap is None else colmap[cols]
        meta.append((rl, b0, L, cg))
        idx.append(np.arange(9 * b0, 9 * (b0 + L), dtype=np.int64))
    got_all = values[torch.from_numpy(np.concatenate(idx)).to(values.device)].cpu().numpy()
    at = 0
    for rl, b0, L, cg in meta:
        blk = got_all[at:at + 9 * L].reshape(3, L, 3)            # (a, p, b)
        at += 9 * L
        for a in range(3):
            r = dim * rl + a
            seg = slice(rp[r], rp[r + 1])
            ref_cols = dim * gid_true[ci[seg] // dim] + ci[seg] % dim
            mine_cols = (dim * cg[:, None] + np.arange(3)[None, :]).ravel()
            order = np.argsort(mine_cols)
            o2 = np.argsort(ref_cols)
            if mine_cols.size != ref_cols.size or not np.array_equal(mine_cols[order], ref_cols[o2]):
                ok = False
                continue
            d = blk[a].ravel()[order] - va[seg][o2]
            num += float((d * d).sum()); den += float((va[seg][o2] ** 2).sum())
    rel = float(np.sqrt(num / den)) if den > 0 else float("nan")
    return {"rows_checked": int(rows.size) * 3, "by_position": {str(k): int((on_plane[rows] == k).sum()) for k in range(4)},
            "pattern": "identical" if ok else "DIFFERS", "rel_frobenius": rel, "tolerance": 1e-12, "checker": kind,
            "result": "pass" if ok and rel <= 1e-12 else "FAIL"}


def _worker(M):
    return oracle_elasticity_time(M)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; FEDDLib itself needs Trilinos/MPI, not in
    this image) on all host cores, one independent element partition (sub-cube) per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    M = args.ref_M
    oracle_elasticity_time(2)  # build/load
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_worker, [M] * cores)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_worker, [M] * cores)
            times.append(time.perf_counter() - t0)
    ne = res[0][1] * cores
    dt = float(np.mean(times))
    value = ne / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"P2 tet linear elasticity, structured cube, {cores} partitions of 6*{M}^3 tets "
                                   "(bounded sample of config 3)", "lambda": LAM, "mu": MU},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                             "sample": f"{cores} x (6*{M}^3 = {res[0][1]}) tets per step, one process per core, "
                                       "FEDDLib's assemblyLinElasXDim loop (one entry per insert) + per-row sort/merge "
                                       "standing in for Tpetra fillComplete; ghost-row merge not included"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--M", type=int, default=70, help="H/h of the cube per GPU (70 = config 3)")
    ap.add_argument("--mode", default="gather", choices=["gather", "coloured", "atomic"])
    ap.add_argument("--ref-M", dest="ref_M", type=int, default=14)
    ap.add_argument("--cpu-M", dest="cpu_M", type=int, default=16)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: exchange after the whole assembly")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU ghost rows: 'fused' = stored straight into the owners' buffers over NVLink peer memory by the "
                         "ghost-row kernels; 'nccl' = all-to-all-v of the ghost values on a side stream")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true", help="skip the sampled-row parity check of the full-size matrix")
    ap.add_argument("--no-ns", action="store_true", help="skip the secondary Navier-Stokes block timings")
    ap.add_argument("--no-ns-mgpu", dest="no_ns_mgpu", action="store_true", help="several GPUs: skip the Navier-Stokes block on the partition")
    ap.add_argument("--ns-M", dest="ns_M", type=int, default=50, help="H/h of the P2-P1 cube of the secondary timings")
    ap.add_argument("--cfg4-levels", dest="cfg4_levels", type=int, default=2, help="regular refinements of DFG3DCylinder_6k (config 4)")
    ap.add_argument("--all-modes", action="store_true", help="also time the other scatter modes (extra keys)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from feddlib_b200 import BLOCK_FULL, Context, Mesh, Pattern
    from feddlib_b200 import mesh as PM

    dim, fe, M = 3, "P2", args.M
    ctx = Context(local_rank)
    numa_node = ctx.bind_host_numa()      # pinned staging buffers next to this rank's GPU (e2e at N > 1)
    ctx.set_scatter_mode(args.mode)
    if world == 1:
        conn, coords, gid = PM.build_structured(dim, fe, 1, M)
        t0 = time.perf_counter()
        mesh = Mesh(ctx, dim, conn, coords)
        pat = Pattern(ctx, mesh)
        ctx.synchronize()
        t_pattern = time.perf_counter() - t0
        runner, t_nccl_warm = None, None
    else:
        from feddlib_b200 import dist as fdist
        # NCCL sets its peer-to-peer channels up on the first all-to-all (seconds at 8 ranks): do that before the plan is timed
        t0 = time.perf_counter()
        w = torch.zeros(world, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(torch.empty_like(w), w)
        torch.cuda.synchronize()
        t_nccl_warm = time.perf_counter() - t0
        t0 = time.perf_counter()
        runner = fdist.DistributedElasticity(ctx, dim, fe, M, rank, world)
        mesh, pat, conn, coords = runner.mesh, runner.pat, runner.conn, runner.coords
        ctx.synchronize()
        t_pattern = time.perf_counter() - t0
    ne = conn.shape[0]
    nnz = pat.nnz(dim, dim, BLOCK_FULL)
    values = ctx.empty_values(nnz)

    overlap = runner is not None and args.mode == "gather" and not args.no_overlap

    fused = overlap and args.exchange == "fused" and world <= 8
    if fused:
        # the peer-memory exchange needs CUDA IPC between the ranks of the box: set it up now (first assembly) and let
        # every rank fall back to the NCCL exchange together if any of them cannot map its peers
        ok = torch.ones(1, dtype=torch.int32, device=f"cuda:{local_rank}")
        try:
            runner.assemble_linelas_fused(values, LAM, MU)
            ctx.synchronize()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer-memory exchange unavailable ({exc}); using the NCCL exchange", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        fused = bool(int(ok.item()))

    def step():
        if fused:        # ghost rows first, stored by their kernels into the owners' receive buffers (peer memory)
            runner.assemble_linelas_fused(values, LAM, MU)
        elif overlap:    # ghost rows first, NCCL exchange on a side stream while the owned rows are assembled
            runner.assemble_linelas_overlapped(values, LAM, MU)
        else:
            pat.assemble_linelas_d(values, LAM, MU)
            if runner is not None:
                runner.exchange(values)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (ctx.launches - l0) // steps

    # parity of the full-size workload (outside every timed region): sampled rows against the reference on their element stars; on
    # several GPUs also the peer-memory exchange against the NCCL exchange, bitwise, on every rank
    parity = None
    if args.mode == "gather" and not args.no_parity:
        step()
        ctx.synchronize()
        parity = parity_check(values, pat, M, runner.dims if runner is not None else (1, 1, 1), rank, runner.plan if runner is not None else None)
        if runner is not None:
            n_owned_vals = pat.nnz_owned(dim, dim, BLOCK_FULL)
            v_nccl = values.clone()
            runner.assemble_linelas_overlapped(v_nccl, LAM, MU)
            ctx.synchronize()
            parity["nccl_exchange_equals_timed_path_bitwise"] = bool(torch.equal(v_nccl[:n_owned_vals], values[:n_owned_vals]))
            del v_nccl
            flags = torch.tensor([1 if parity["result"] == "pass" and parity["nccl_exchange_equals_timed_path_bitwise"] else 0], device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            parity["all_ranks"] = "pass" if int(flags.item()) == 1 else "FAIL"
    with ClockSampler(local_rank) as clk:
        ms, launches = timed(step, args.steps, args.warmup)
    clocks = clk.summary()

    # several GPUs: where a step's time goes (outside the timed region: three more steps with events around the phases, max over ranks)
    phases = None
    if fused:
        runner.phase_timing = True
        acc = None
        for _ in range(3):
            step()
            ctx.synchronize()
            pm = runner.phase_ms()
            acc = pm if acc is None else {k: acc[k] + pm[k] for k in pm}
        runner.phase_timing = False
        keys = sorted(acc)
        t = torch.tensor([acc[k] / 3 for k in keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = {k: float(v) for k, v in zip(keys, t.tolist())}
        phases["what"] = "device ms per phase of one step, max over ranks; the barrier runs under the owned rows"

    # several GPUs: the fused Navier-Stokes (0,0) block on the same partition (same node pattern and value layout as the
    # elasticity matrix): velocity on the unique map -> repeated map on the device (MultiVector::importFromVector,
    # NavierStokes_def.hpp:294), assembly with the NCCL ghost-row exchange under the owned rows, unpack-add
    ns_mgpu = None
    if runner is not None and args.mode == "gather" and not args.no_ns_mgpu:
        u_unique = torch.from_numpy(np.random.default_rng(1234 + rank).uniform(-1, 1, dim * runner.plan.n_owned)).cuda()

        def ns_step():
            u_rep = runner.import_vector(u_unique, dim)
            runner.assemble_overlapped(values, dim, dim, BLOCK_FULL, lambda: pat.assemble_ns_jacobian_d(values, u_rep, 1.0, 1.0e-3, True))

        ns_ms, ns_launches = timed(ns_step, max(3, args.steps // 2), 3)
        ns_mgpu = {"workload": f"fused (0,0) block rho*nu*A + rho*N(u) + rho*W(u), P2 cube H/h={M} per GPU, {ne * world} tets on {world} GPUs, "
                               "velocity import + NCCL ghost-row exchange inside the step",
                   "ms_per_step": ns_ms, "elements_per_s": ne * world / (ns_ms * 1e-3), "launches_per_step": ns_launches,
                   "hbm_frac_per_gpu": (ne * conn.shape[1] * 4 + (M + 1) ** 3 * dim * 8 + nnz * 8) / (ns_ms * 1e-3) / 1e9 / peaks()[0]}
        step(); ctx.synchronize()   # leave the elasticity matrix in `values` for the checksum / e2e below

    ne_total = ne * world
    value = ne_total / (ms * 1e-3)

    # roofline of the assembly pass (all launches of one step on this rank): algorithmic bytes of SURVEY.md 8(d)
    n_vertices = (M + 1) ** 3
    alg_bytes = ne * conn.shape[1] * 4 + n_vertices * dim * 8 + nnz * 8
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(M) if args.mode == "gather" and world == 1 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes,
                "kernel": f"assembly pass = k_geom + k_ring (edge-node rows, ring order, TMA bulk stores) + k_task "
                          f"(vertex-node rows, block tasks) bucket launches ({launches} launches/step); time = whole pass, CUDA events"
                if args.mode == "gather" else f"{args.mode} scatter pass ({launches} launches/step)"}

    extra = {}
    if args.all_modes and world == 1:
        for m in ("gather", "coloured", "atomic"):
            if m == args.mode:
                continue
            ctx.set_scatter_mode(m)
            ms_m, l_m = timed(step, max(2, args.steps // 3), 3)
            extra[m] = {"ms_per_step": ms_m, "value": ne / (ms_m * 1e-3), "launches": l_m,
                        "hbm_frac": alg_bytes / (ms_m * 1e-3) / 1e9 / peak}
        ctx.set_scatter_mode(args.mode)

    # e2e: host buffers through the C ABI's host-pointer entry points; H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        pinned_vals = torch.empty(nnz, dtype=torch.float64, pin_memory=True)
        pinned_xyz = torch.from_numpy(coords).pin_memory()
        out_np, xyz_np = pinned_vals.numpy(), pinned_xyz.numpy()

        def e2e_step():
            mesh.update_coords(xyz_np)                      # H2D of this step's points (pinned)
            pat.assemble_linelas(LAM, MU, out=out_np)       # assemble + D2H of the CSR values (pinned)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the device->host copy of the CSR values alone (all ranks at the same time): what the platform gives for that copy
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            pinned_vals.copy_(values, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        d2h = (time.perf_counter() - t0) / 2
        if world > 1:
            t = torch.tensor([d2h], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d2h = float(t.item())
        e2e = {"value": ne_total / dt, "unit": UNIT, "h2d_bytes_per_step": int(coords.nbytes),
               "d2h_bytes_per_step": int(nnz * 8), "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "host_numa_node": numa_node,
               "d2h_copy_alone_ms": d2h * 1e3, "d2h_copy_alone_GBs_per_gpu": nnz * 8 / d2h / 1e9,
               "d2h_copy_alone_GBs_all_gpus": world * nnz * 8 / d2h / 1e9}
        checksum = float(out_np[: min(nnz, 1 << 20)].sum())
        del pinned_vals
        # the same end-to-end step through the compiled C++ host layer (FEDD::FE_b200 over the C ABI: host containers in,
        # fill-complete host CSR in a pooled page-locked buffer out); N = 1 only, the binary is built by build()
        exe = os.path.join(ROOT, "feddlib_b200", "bench_fe_b200")
        if world == 1 and os.path.exists(exe):
            try:
                r = subprocess.run([exe, str(M), str(max(2, args.e2e_steps)), str(local_rank)], capture_output=True, text=True, timeout=600)
                cpp = json.loads(r.stdout.strip().splitlines()[-1])
                e2e["cpp_host"] = {"value": ne / (cpp["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": cpp["ms_per_step"],
                                   "ms_best": cpp["ms_best"], "update_points_ms": cpp.get("update_points_ms"),
                                   "first_call_s": cpp["first_call_s"], "addFE_s": cpp["addFE_s"],
                                   "checksum_first_1Mi_values": cpp["checksum_first_1Mi_values"],
                                   "what": "FEDD::FE_b200::updatePoints + assemblyLinElasXDim (C++): the reference's vector-of-vectors points "
                                           "flattened into a page-locked staging buffer + H2D, assembly, CSR values D2H into a pooled "
                                           "page-locked buffer owned by the returned matrix"}
            except Exception as exc:  # noqa: BLE001
                e2e["cpp_host"] = {"error": str(exc)[:200]}
    else:
        checksum = float(values[: min(nnz, 1 << 20)].sum().item())

    # secondary numbers (not the headline): the Navier-Stokes blocks of config 4 on a structured P2-P1 cube, same engine
    ns_extra = cfg4 = cfg12 = None
    if world == 1 and not args.no_ns:
        ns_extra = ns_jacobian_numbers(ctx, args.ns_M, peak)
        cfg4 = config4_numbers(ctx, args.cfg4_levels, peak)
        cfg12 = {"config1": laplace_numbers(ctx, 2, "P1", 224, "config 1", peak, steps=50),
                 "config2": laplace_numbers(ctx, 3, "P2", 100, "config 2", peak)}

    cpu_baseline = None
    if rank == 0:
        t_cpu, ne_cpu = oracle_elasticity_time(args.cpu_M)
        cpu_baseline = {"value": ne_cpu / t_cpu, "unit": UNIT, "cores": 1, "kind": cpu_kind(),
                        "sample": f"6*{args.cpu_M}^3 = {ne_cpu} P2 tets of the same cube workload; "
                                  + ("FEDDLib's own assemblyLinElasXDim compiled against mock Trilinos containers"
                                     if cpu_kind() == "reference" else "oracle restatement of assemblyLinElasXDim")
                                  + f" (one entry per insert + per-row sort/merge), {t_cpu:.1f} s on 1 core"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{'config 3' if M == 70 else ('config 5' if M == 101 else 'config 3-like')}{' per GPU (weak-scaling box of sub-cubes)' if world > 1 else ''}: "
                                       f"linear elasticity P2 3D cube H/h={M} per GPU, {ne} tets/GPU, "
                                       f"{nnz} CSR values/GPU (30x30 local blocks)",
                           "lambda": LAM, "mu": MU, "scatter_mode": args.mode,
                           "l2_policy": "outputs (5.7 GB at M=70) exceed the 126 MB L2; no flush needed",
                           "parallelism": f"element partition over {world} GPU(s)" + (
                               "; ghost rows assembled first and stored by their kernels into the owners' receive buffers over NVLink "
                               "peer memory (CUDA IPC), one-element all-reduce as barrier under the owned rows, unpack-add"
                               if fused else "; ghost rows assembled first, NCCL ghost-row exchange overlapped with the owned rows"
                               if overlap else ("; NCCL ghost-row exchange after the assembly" if world > 1 else "")),
                           "pattern_build_s": t_pattern, "pattern_build_breakdown": getattr(runner, "timing", None),
                           "nccl_first_alltoall_s": t_nccl_warm if world > 1 else None},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "checksum_first_1Mi_values": checksum, "parity_check": parity}
        if phases:
            line["multi_gpu_phase_ms"] = phases
        if ns_mgpu:
            line["navier_stokes_block_multi_gpu"] = ns_mgpu
        if extra:
            line["other_scatter_modes"] = extra
        if ns_extra:
            line["navier_stokes_blocks"] = ns_extra
        if cfg4:
            line["config4_navier_stokes_dfg3d"] = cfg4
        if cfg12:
            line["laplace_configs_1_2"] = cfg12
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
