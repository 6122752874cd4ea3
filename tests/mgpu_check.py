"""Multi-GPU parity check, launched by tests/test_multigpu_gpu.py (or by hand) as
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
Every rank owns one brick of a periodic 3D3V k=3 lattice, exchanges upwind ghost faces over NCCL
(hyperdeal_b200.partition), applies the operator in two parts (interior overlapped with the exchange, then the
boundary layer; split and fused-halo variants; three complete fused rk45 steps) and compares its brick with the same
operator / integrator applied to the WHOLE lattice on its own GPU (single brick, periodic wrap) — which tests/test_apply_gpu.py pins against the oracle.  Prints one 'MGPU OK' line."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hyperdeal_b200 import api  # noqa: E402
from hyperdeal_b200.partition import BrickPartition, HaloExchange  # noqa: E402

VEL = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nloc = (3, 2, 2, 2, 2, 2)
    # kernel 6 = three-round 3D3V kernel (the automatic choice), 2 = two-role pipelined kernel, 1 = generic kernel;
    # "x24" = the layout bench.py uses at 8 GPUs (x_2 cut in four, x_1 in two, rows of cells along x_0 whole), at world 4
    # its 4-GPU restriction (x_2 in four); bricks one cell thick in the cut direction included
    cases = [(6, (2, 1, 0), None, nloc), (2, (2, 1, 0), None, nloc), (6, (0, 1, 2), None, nloc), (1, (2, 1, 0), None, nloc)]
    if world == 8:
        cases += [(6, None, (1, 2, 4, 1, 1, 1), (4, 2, 1, 2, 2, 2)), (2, None, (1, 2, 4, 1, 1, 1), (4, 2, 2, 2, 2, 2))]
    elif world == 4:
        cases += [(6, None, (1, 1, 4, 1, 1, 1), (4, 2, 1, 2, 2, 2)), (6, None, (1, 2, 2, 1, 1, 1), (4, 2, 2, 2, 2, 2))]
    for kernel, split_order, grid, nloc in cases:
        part = BrickPartition(world, rank, nloc, split_order=split_order) if grid is None else BrickPartition(world, rank, nloc, grid=grid)
        ctx = api.Context(local)
        left, right = (0.0,) * 6, (1.0,) * 6
        mf = api.MatrixFree(ctx, 3, 3, 3, nloc, left, right, n_cells_global=part.n_cells_global, cell_offset=part.cell_offset, side_kind=part.side_kind)
        op = api.AdvectionOperation(mf, VEL, 0.5)
        op.set_kernel(kernel)
        # whole lattice on this GPU
        mf_all = api.MatrixFree(ctx, 3, 3, 3, part.n_cells_global, left, right)
        op_all = api.AdvectionOperation(mf_all, VEL, 0.5)
        op_all.set_kernel(kernel)
        u = np.random.default_rng(1234).standard_normal(mf_all.n_dofs)
        a_src, a_dst = mf_all.initialize_dof_vector(), mf_all.initialize_dof_vector()
        mf_all.copy_in(a_src, u)
        op_all.apply(a_dst, a_src, 0.0)
        ref_all = mf_all.copy_out(a_dst).reshape(tuple(reversed(part.n_cells_global)) + (4096,))
        sl = tuple(slice(part.cell_offset[d], part.cell_offset[d] + nloc[d]) for d in reversed(range(6)))
        mine = np.ascontiguousarray(u.reshape(ref_all.shape)[sl]).reshape(-1)
        expect = np.ascontiguousarray(ref_all[sl]).reshape(-1)
        src = torch.from_numpy(mine).cuda()
        dst = torch.zeros_like(src)
        halo = mf.halo_total
        send = torch.zeros(max(halo, 1), dtype=torch.float64, device="cuda")
        ghost = torch.full((max(halo, 1),), float("nan"), dtype=torch.float64, device="cuda")
        offsets = {(d, s): mf.halo_offset(d, s) for d in range(6) for s in range(2)}
        sizes = {(d, s): mf.ghost_size(d, s) for d in range(6) for s in range(2)}
        ex = HaloExchange(part, offsets, sizes, op.ghost_sides())
        for it in range(2):
            mf.halo_pack(src.data_ptr(), send.data_ptr(), send_mask=ex.send_mask())
            works = ex.start(send, ghost)
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_INTERIOR)
            HaloExchange.finish(works)
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_BOUNDARY)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        rel = float(np.max(np.abs(got - expect)) / np.max(np.abs(expect)))
        # direct variant: the pack kernel stores into the neighbours' ghost buffers over NVLink (peer-mapped pointers)
        from hyperdeal_b200.partition import PeerHaloExchange

        peer = PeerHaloExchange(part, offsets, sizes, halo, op.ghost_sides(), torch.device("cuda", local), ctx=ctx)
        dst.zero_()
        for it in range(4):  # split variant: pack kernel, stream flags, two operator launches (all on one stream)
            g, m = peer.start(mf, ctx, src.data_ptr())
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_INTERIOR)
            peer.wait_ready(ctx, m)
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_BOUNDARY)
            peer.consumed(ctx)
        torch.cuda.synchronize()
        rel = max(rel, float(np.max(np.abs(dst.cpu().numpy() - expect)) / np.max(np.abs(expect))))
        if op.kernel_name.startswith(("advect_3d3v", "rounds_3d3v")):
            # fused variant: ONE kernel per step packs, sends over NVLink, does the interior, waits, does the boundary
            dst.zero_()
            for it in range(5):
                g, sends, counters, epoch = peer.begin_fused(ctx, op)
                op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), sends, counters, epoch)
                peer.consumed(ctx)
            torch.cuda.synchronize()
            assert not op.overlap_timed_out()
            rel = max(rel, float(np.max(np.abs(dst.cpu().numpy() - expect)) / np.max(np.abs(expect))))
            # complete rk45 steps: per stage ONE kernel per GPU (pack + NVLink stores + operator + stage update,
            # hd_lsrk_stage_overlapped) against the single-GPU fused integrator on the whole lattice
            a_ki, a_ti = mf_all.initialize_dof_vector(), mf_all.initialize_dof_vector()
            rk_all = api.LowStorageRungeKuttaIntegrator(mf_all, a_ki, a_ti, "rk45")
            sol, ki, ti = src.clone(), torch.zeros_like(src), torch.zeros_like(src)
            rk = api.LowStorageRungeKuttaIntegrator(mf, ki.data_ptr(), ti.data_ptr(), "rk45")
            time, dt = 0.0, 2e-3
            for it in range(3):
                rk_all.perform_time_step(a_src, time, dt, op_all)
                rk.perform_time_step_partitioned(sol.data_ptr(), time, dt, op, peer, ctx)
                time += dt
            torch.cuda.synchronize()
            assert not op.overlap_timed_out()
            exp_sol = np.ascontiguousarray(mf_all.copy_out(a_src).reshape(ref_all.shape)[sl]).reshape(-1)
            rel = max(rel, float(np.max(np.abs(sol.cpu().numpy() - exp_sol)) / np.max(np.abs(exp_sol))))
        t = torch.tensor([rel], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("MGPU %s world=%d grid=%s kernel=%s rel=%.3e" % ("OK" if t.item() <= 1e-13 else "FAIL", world, part.grid, op.kernel_name, t.item()), flush=True)
        assert t.item() <= 1e-13, t.item()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
