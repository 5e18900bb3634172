"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
per-rank device assembly + NCCL ghost-row exchange + unpack-add must reproduce the oracle's global matrix.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from feddlib_b200 import BLOCK_FULL, Context
    from feddlib_b200 import mesh as PM
    from feddlib_b200.dist import DistributedElasticity
    from oracle import oracle as O

    dim, fe, M = 3, "P2", int(os.environ.get("T_M", "3"))
    lam, mu = 8e6, 2e6
    ctx = Context(local)
    for mode in ("gather", "coloured", "atomic"):
        ctx.set_scatter_mode(mode)
        run = DistributedElasticity(ctx, dim, fe, M, rank, world)
        plan, pat = run.plan, run.pat
        values = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        pat.assemble_linelas_d(values, lam, mu)
        run.exchange(values)
        ctx.synchronize()
        vals = values.cpu().numpy()
        # oracle: global matrix over all ranks' elements
        dims = run.dims
        nglob = 1
        for d in range(dim):
            nglob *= dims[d] * 2 * M + 1
        G = O.Matrix(dim * nglob, 64)
        for r in range(world):
            c2, x2, g2, _ = PM.build_structured_box(dim, fe, dims, M, r)
            O.assembly_linelas(dim, fe, c2, x2, g2, lam, mu, G)
        Gs = G.scipy().tocsr()
        rp, ci = plan.rowptr, plan.colind
        rpd, cid = pat.expand(dim, dim, BLOCK_FULL)
        n_owned_dofs = dim * plan.n_owned
        col_gid_dof = (dim * plan.colmap_gids[:, None] + np.arange(dim)[None, :]).ravel()
        err = ref = 0.0
        checked = 0
        for row in range(n_owned_dofs):
            grow = dim * plan.unique_gids[row // dim] + row % dim
            seg, segv = Gs.indices[Gs.indptr[grow]:Gs.indptr[grow + 1]], Gs.data[Gs.indptr[grow]:Gs.indptr[grow + 1]]
            mine_c = col_gid_dof[cid[rpd[row]:rpd[row + 1]]]
            mine_v = vals[rpd[row]:rpd[row + 1]]
            order = np.argsort(mine_c)
            assert np.array_equal(mine_c[order], seg), f"pattern mismatch rank {rank} row {row}"
            err += ((mine_v[order] - segv) ** 2).sum(); ref += (segv ** 2).sum(); checked += seg.size
        rel = float(np.sqrt(err / ref))
        assert rel <= 1e-12, f"rank {rank} mode {mode}: relative Frobenius error {rel:.3e}"
        tot = torch.tensor([checked], device="cuda"); dist.all_reduce(tot)
        assert int(tot) == Gs.nnz
        # overlapped variant (ghost rows first, exchange concurrent with the owned rows): same owned values, bitwise
        values2 = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        values2.fill_(float("nan"))
        run.assemble_linelas_overlapped(values2, lam, mu)
        ctx.synchronize()
        n_owned_vals = pat.nnz_owned(dim, dim, BLOCK_FULL)
        if mode == "atomic":   # atomics: summation order not fixed
            assert torch.allclose(values2[:n_owned_vals], values[:n_owned_vals], rtol=1e-12, atol=1e-6), f"rank {rank} mode {mode}: overlapped exchange differs"
        else:
            assert torch.equal(values2[:n_owned_vals], values[:n_owned_vals]), f"rank {rank} mode {mode}: overlapped exchange differs"
        if mode == "gather":
            # fused variant: ghost rows stored straight into the owners' receive buffers over NVLink peer memory; run it
            # several times (the receive buffers alternate) -- every result must equal the NCCL exchange bitwise
            for it in range(4):
                values3 = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
                values3.fill_(float("nan"))
                run.assemble_linelas_fused(values3, lam, mu)
                ctx.synchronize()
                assert torch.equal(values3[:n_owned_vals], values[:n_owned_vals]), f"rank {rank}: fused peer-memory exchange differs (pass {it})"
            dist.barrier()
            run.close_peer()
        if mode == "gather":
            # load vector: FE::assemblyRHS on the device + export/ADD of the ghost rows (NCCL) = global oracle vector
            f = np.array([1.5, -2.0, 0.25])
            for vec_field in (False, True):
                dofs = dim if vec_field else 1
                got = run.assemble_rhs(f, 1, vec_field)[: dofs * plan.n_owned].cpu().numpy()
                glob = np.zeros(dofs * nglob)
                for r in range(world):
                    c2, x2, g2, _ = PM.build_structured_box(dim, fe, dims, M, r)
                    np.add.at(glob, (dofs * g2[:, None] + np.arange(dofs)[None, :]).ravel(), O.assembly_rhs(dim, fe, c2, x2, f, 1, vec_field))
                want = glob[(dofs * plan.unique_gids[:, None] + np.arange(dofs)[None, :]).ravel()]
                assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max(), f"rank {rank}: load vector export/ADD"
        if mode == "gather":
            # Navier-Stokes (0,0) block on several GPUs: the velocity lives on the unique map, is imported to the repeated map on
            # the device (MultiVector::importFromVector before the advection assemblies, NavierStokes_def.hpp:294), the fused block
            # rho*nu*A + rho*N(u) + rho*W(u) is assembled with the NCCL ghost-row exchange overlapped; N(u) alone as well
            import scipy.sparse as sp
            rho, nu = 1.3, 1.0e-3
            ufun = lambda x: np.stack([np.sin(2 * x[:, 1]) + x[:, 2], -x[:, 0] ** 2, 0.3 + x[:, 0] * x[:, 1]], axis=1).ravel()
            owned = plan.rep_of_row[: plan.n_owned]
            u_unique = torch.from_numpy(ufun(run.coords[owned])).cuda()
            u_rep = run.import_vector(u_unique, dim)
            assert np.array_equal(u_rep.cpu().numpy(), ufun(run.coords)), f"rank {rank}: unique -> repeated import"
            vN = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
            run.assemble_overlapped(vN, dim, dim, BLOCK_FULL, lambda: pat.assemble_ns_jacobian_d(vN, u_rep, rho, nu, True))
            ctx.synchronize()
            GA, GN, GW = O.Matrix(dim * nglob, 64), O.Matrix(dim * nglob, 64), O.Matrix(dim * nglob, 64)
            for r in range(world):
                c2, x2, g2, _ = PM.build_structured_box(dim, fe, dims, M, r)
                O.assembly_laplace_vecfield(dim, fe, c2, x2, g2, GA)
                O.assembly_advection(dim, fe, c2, x2, g2, ufun(x2), GN)
                O.assembly_advection_in_u(dim, fe, c2, x2, g2, ufun(x2), GW)
            J = (rho * nu * GA.scipy() + rho * GN.scipy() + rho * GW.scipy()).tocsr()
            rows = (dim * plan.unique_gids[:, None] + np.arange(dim)[None, :]).ravel()
            got = sp.csr_matrix((vN.cpu().numpy()[: pat.nnz_owned(dim, dim, BLOCK_FULL)], col_gid_dof[cid[: rpd[n_owned_dofs]]], rpd[: n_owned_dofs + 1]),
                                shape=(n_owned_dofs, dim * nglob))
            diff = got - J[rows]
            relJ = float(np.sqrt(diff.multiply(diff).sum() / J[rows].multiply(J[rows]).sum()))
            assert relJ <= 1e-12, f"rank {rank}: multi-GPU Navier-Stokes block, relative Frobenius error {relJ:.3e}"
            # ghost rows NEXT TO the owned rows (row phases ROWS_GEOM / ROWS_GHOST_ONLY, side stream): same matrix, bitwise
            if world > 1:
                vA = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL)); vB = ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
                run.concurrent_ghost = False
                run.assemble_linelas_fused(vA, lam, mu)
                ctx.synchronize(); dist.barrier()
                run.concurrent_ghost = True
                run.assemble_linelas_fused(vB, lam, mu)
                ctx.synchronize(); dist.barrier()
                run.concurrent_ghost = False
                nown = pat.nnz_owned(dim, dim, BLOCK_FULL)
                assert torch.equal(vA[:nown], vB[:nown]), f"rank {rank}: concurrent ghost rows differ from the serial phases"
            # the peer-memory path must refuse operators whose kernels do not store ghost rows through it
            from feddlib_b200 import LogicError
            try:
                run.assemble_fused(vN, dim, dim, BLOCK_FULL, lambda: pat.assemble_ns_jacobian_d(vN, u_rep, rho, nu, True))
                refused = world == 1
            except LogicError:
                refused = True
            ctx.synchronize()
            dist.barrier()
            run.close_peer()
            assert refused, f"rank {rank}: assemble_fused accepted an operator without peer-memory ghost rows"
            print(f"[dist_gpu_check] rank {rank}/{world}: Navier-Stokes block on {world} GPUs rel. error {relJ:.2e} OK", flush=True)
        print(f"[dist_gpu_check] rank {rank}/{world} mode {mode}: owned rows {plan.n_owned}, ghost rows {plan.n_ghost}, "
              f"rel. error {rel:.2e} OK (overlapped exchange equal)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
