#!/usr/bin/env python
"""bench.py — advection-operator throughput (GDoF/s) on B200, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one application  dst = M^-1 A(src)  of the 3D3V degree-3 FP64 advection operator
(the loop body of performance/operators_advection_01.likwid.cc:205-239) on a Cartesian periodic
phase-space lattice with 8^6 cells = 1 073 741 824 DoFs per GPU (BASELINE.json configs[1],
performance/operators_advection_01/node_level_basic.json) and synthetic data (sin*cos wave
sampled at the GLL nodes, velocity (1, .15, -.05, .1, -.15, .5), SkewFactor 0.5).

  value       GDoF/s with src/dst resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e         the same metric through hd_advection_apply_host on pinned HOST buffers: H2D of src and
              D2H of dst inside the timed region (N=1: rank 0's 8 GiB + 8 GiB per step)
  roofline    algorithmic 16 B/DoF (read src once + write dst once) over the measured
              kernel time against MEASURED_PEAKS.json's copy bandwidth
  cpu_baseline  the CPU restatement of the reference's literal ECL algorithm (oracle/, "port") on
              the host cores, on a bounded sample (4^6-cell lattice of the same discretisation)
  --impl reference   times only that CPU restatement (the hyper.deal binary itself cannot be built
              here: deal.II and MPI are absent)

N > 1: weak scaling, one brick of 8^6 cells per GPU; the lattice is doubled along x_2, x_1, x_0
(examples/advection/performance/weak.py:95-101 doubles the x-directions first) and the bricks
exchange the upwind ghost faces over NCCL each step (pack -> send/recv overlapped with the interior cells
-> boundary-layer cells), all inside the timed region.
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

VELOCITY = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
SKEW = 0.5
DEGREE = 3
CELLS_PER_DIR = 8
BYTES_PER_DOF = 16.0  # FP64: read src + write dst


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel_name):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_name)
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line): NVML polled
    every ~2 ms from a thread (the timed region is only tens of milliseconds long); nvidia-smi -lms as the fallback."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index
        self.sm, self.reasons, self.mx, self.nvml, self.stop_flag, self.thread = [], set(), None, None, False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = None
            try:
                import torch

                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if uuid else b"")
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_baseline(seconds_hint=15.0):
    """CPU restatement of the reference's ECL kernel (oracle/hd_oracle.cpp), all host cores, on a
    4^6-cell 3D3V k=3 lattice (same discretisation, 1/64 of the cells)."""
    import numpy as np

    from oracle import oracle as O

    cores = os.cpu_count() or 1
    nc = (4,) * 6
    mesh = O.Mesh(3, 3, nc, (0.0,) * 6, (1.0,) * 6, (True,) * 6)
    orc = O.Oracle(mesh, DEGREE, skew=SKEW, velocity=VELOCITY, nthreads=cores)
    src = orc.interpolate(O.hyperrectangle_exact, 0.0)
    dst = np.zeros_like(src)
    orc.apply(src, dst=dst)  # warm-up (also builds the library)
    t0, n = time.perf_counter(), 0
    while True:
        orc.apply(src, dst=dst)
        n += 1
        el = time.perf_counter() - t0
        if el > seconds_hint or n >= 50:
            break
    return {"value": orc.ndofs * n / el / 1e9, "unit": "GDoF/s", "cores": cores, "kind": "port", "sample": "%d applies of the 4^6-cell (%.1fM DoF) 3D3V k=3 lattice, %d threads" % (n, orc.ndofs / 1e6, cores), "seconds": el}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    import numpy as np

    from oracle import oracle as O

    cores = os.cpu_count() or 1
    nc = (4,) * 6
    mesh = O.Mesh(3, 3, nc, (0.0,) * 6, (1.0,) * 6, (True,) * 6)
    orc = O.Oracle(mesh, DEGREE, skew=SKEW, velocity=VELOCITY, nthreads=cores)
    src = orc.interpolate(O.hyperrectangle_exact, 0.0)
    dst = np.zeros_like(src)
    for _ in range(max(1, min(args.warmup, 2))):
        orc.apply(src, dst=dst)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.apply(src, dst=dst)
    el = time.perf_counter() - t0
    v = orc.ndofs * steps / el / 1e9
    sample = "each step = one apply on a 4^6-cell (%.1fM DoF) 3D3V k=3 lattice, %d threads; CPU restatement of hyper.deal's ECL kernel (the hyper.deal binary needs deal.II+MPI, absent)" % (orc.ndofs / 1e6, cores)
    print(json.dumps({
        "impl": "reference", "metric": "advection operator throughput (3D3V, k=3, FP64)", "value": v, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": "3D3V k=3 FP64 advection apply, Cartesian periodic, skew 0.5 (CPU sample: 4^6 cells)"},
        "cpu_baseline": {"value": v, "unit": "GDoF/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=CELLS_PER_DIR, help="cells per direction per GPU (default 8 = the BASELINE workload)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 fused 3D3V kernel")
    ap.add_argument("--layout", default="x24", choices=["x", "xv", "x24"],
                    help="how the N = 8 lattice (16,16,16,8,8,8) is cut: x = x_2,x_1,x_0 in two (bricks of 8^6 cells); xv = x_2,x_1,v_2 in two (16x8x8x8x8x4); "
                         "x24 (default) = x_2 in four, x_1 in two (16x8x4x8x8x8): two cut directions instead of three, rows of cells (along x_0) stay whole")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N > 1: direct peer-memory stores over NVLink (default) or NCCL send/recv")
    ap.add_argument("--overlap", default="kernel", choices=["kernel", "parts"], help="N > 1: one launch that waits for the halo flag in-kernel (default) or interior/boundary launches")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    from hyperdeal_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    from hyperdeal_b200.partition import BrickPartition, HaloExchange

    # global lattice of the weak-scaling recipe (x_2, x_1, x_0 doubled in turn); how it is cut into equal bricks is ours
    recipe = BrickPartition(world, rank, [args.cells] * 6, split_order=(2, 1, 0))
    if args.layout == "x" or world < 8:
        part = recipe
    else:
        # never cut x_0 (rows of cells are walked along x_0): the third cut goes through v_2 ("xv") or through x_2 again ("x24")
        cut = BrickPartition(world, rank, [args.cells] * 6, split_order=(2, 1, 5) if args.layout == "xv" else (2, 1, 2))
        nloc = [g // c for g, c in zip(recipe.n_cells_global, cut.grid)]
        part = BrickPartition(world, rank, nloc, grid=cut.grid)
        assert part.n_cells_global == recipe.n_cells_global
    nglob, p = list(part.n_cells_global), list(part.grid)
    ctx = api.Context(local_rank)
    mf = api.MatrixFree(ctx, 3, 3, DEGREE, part.n_cells, (0.0,) * 6, (1.0,) * 6, n_cells_global=part.n_cells_global, cell_offset=part.cell_offset, side_kind=part.side_kind)
    op = api.AdvectionOperation(mf, VELOCITY, SKEW)
    op.set_kernel(args.kernel)
    n_dofs = mf.n_dofs
    src = torch.empty(n_dofs, dtype=torch.float64, device="cuda")
    dst = torch.empty(n_dofs, dtype=torch.float64, device="cuda")
    api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
    dst.zero_()
    halo = mf.halo_total
    send = torch.empty(max(halo, 16), dtype=torch.float64, device="cuda")
    ghost = torch.zeros(max(halo, 16), dtype=torch.float64, device="cuda")
    offsets = {(d, s): mf.halo_offset(d, s) for d in range(6) for s in range(2)}
    sizes = {(d, s): mf.ghost_size(d, s) for d in range(6) for s in range(2)}
    # upwind flux: only the inflow ghost side of every cut direction is read -> only that one is exchanged
    exch = HaloExchange(part, offsets, sizes, op.ghost_sides())
    send_mask = exch.send_mask()
    halo_bytes = exch.bytes_per_exchange[0] * 8

    # N > 1: direct NVLink variant (pack kernel stores into the neighbours' ghost buffers) unless --halo nccl
    peer, halo_mode = None, "none"
    if world > 1:
        halo_mode = "nccl"
        if args.halo == "peer":
            from hyperdeal_b200.partition import PeerHaloExchange

            try:
                peer = PeerHaloExchange(part, offsets, sizes, halo, op.ghost_sides(), torch.device("cuda", local_rank))
                halo_mode = "peer"
            except Exception as e:  # symmetric memory unavailable on this box: NCCL send/recv
                if rank == 0:
                    sys.stderr.write("bench: peer-memory halo unavailable (%s: %s), using NCCL send/recv\n" % (type(e).__name__, e))
            ok = torch.tensor([1 if peer is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                peer, halo_mode = None, "nccl"
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream() if world > 1 else None
    ev_src, ev_halo = torch.cuda.Event(), torch.cuda.Event()

    def step():
        """one operator application; N > 1: start the halo (side stream) -> interior cells -> halo arrived -> boundary cells
        (the reference's overlapping levels, matrix_free.templates.h:1516-1566)"""
        if world == 1:
            op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
            return
        if peer is not None and args.overlap == "kernel":
            # ONE launch: a warp per CTA stores the boundary layers into the neighbours' ghost buffers over NVLink while the
            # others do the interior cells; the boundary layer starts when the neighbours' arrival counters are complete
            g, sends, counters, epoch = peer.begin_fused(ctx, op)
            op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), sends, counters, epoch)
            peer.consumed(ctx)
            return
        ev_src.record(main_stream)
        side_stream.wait_event(ev_src)
        ctx.set_stream(side_stream.cuda_stream)
        m = 0
        with torch.cuda.stream(side_stream):
            if peer is not None:
                g, m = peer.start(mf, ctx, src.data_ptr())
            else:
                mf.halo_pack(src.data_ptr(), send.data_ptr(), send_mask=send_mask)
                HaloExchange.finish(exch.start(send, ghost))
                g = ghost
                ev_halo.record(side_stream)
        ctx.set_stream(main_stream.cuda_stream)
        op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_INTERIOR)
        if peer is not None:
            peer.wait_ready(ctx, m)
        else:
            main_stream.wait_event(ev_halo)
        op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_BOUNDARY)
        if peer is not None:
            peer.consumed(ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = op.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for _ in range(args.steps):
        if world == 1:
            ctx.timer_start()
            step()
            kernel_ms.append(ctx.timer_stop())
        else:
            step()
    ev1.record()
    barrier()
    launches = op.launch_count - launches0
    if peer is not None and args.overlap == "kernel" and op.overlap_timed_out():
        raise SystemExit("bench: the in-kernel halo wait timed out on rank %d (halo never signalled)" % rank)
    if world > 1:
        # per-launch kernel time for the roofline entry: one un-split launch on this rank's brick (outside the timed region)
        ctx.timer_start()
        op.apply(dst.data_ptr(), src.data_ptr(), 0.0, ghosts=ghost.data_ptr())
        kernel_ms = [ctx.timer_stop()]
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    if world > 1 and not (peer is not None and args.overlap == "kernel"):
        launches += args.steps * sum(send_mask)  # pack kernels
    value = n_dofs * world * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end to end with HOST buffers: H2D of src and D2H of dst inside the timed region
    e2e = None
    if not args.no_e2e and world == 1:
        # N = 1: the host-buffer entry point of the C ABI (hd_advection_apply_host: copy-in / kernel / copy-out pipelined over slabs)
        h_src = torch.empty(n_dofs, dtype=torch.float64, pin_memory=True)
        h_dst = torch.empty(n_dofs, dtype=torch.float64, pin_memory=True)
        h_src.copy_(src)
        e_steps = max(1, min(args.steps, 3))
        op.apply_host_ptr(h_dst.data_ptr(), h_src.data_ptr(), 0.0)  # warm-up (allocates staging)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            op.apply_host_ptr(h_dst.data_ptr(), h_src.data_ptr(), 0.0)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        e2e = {"value": n_dofs * e_steps / el / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": n_dofs * 8, "d2h_bytes_per_step": n_dofs * 8, "steps": e_steps,
               "checksum": float(h_dst[:: max(1, n_dofs // 4096)].sum())}
        del h_src, h_dst
    elif not args.no_e2e:
        # N > 1: every rank copies its brick of src in from pinned host memory, runs the step (halo exchange included) and
        # copies its brick of dst out; wall clock between barriers, max over ranks.  Skipped if the box is short of host memory.
        import psutil

        need = 2 * n_dofs * 8 * world
        ok = torch.tensor([1 if psutil.virtual_memory().available > 2 * need else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        h_src = h_dst = None
        if ok.item() == 1:
            try:  # the pinned allocation itself may still fail on one rank: agree on the outcome before anyone enters a barrier
                h_src = torch.empty(n_dofs, dtype=torch.float64, pin_memory=True)
                h_dst = torch.empty(n_dofs, dtype=torch.float64, pin_memory=True)
            except RuntimeError:
                h_src = h_dst = None
            ok = torch.tensor([1 if h_dst is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 1:
            h_src.copy_(src)

            def e2e_step():
                src.copy_(h_src, non_blocking=True)
                step()
                h_dst.copy_(dst, non_blocking=True)

            e_steps = max(1, min(args.steps, 2))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                e2e_step()
            barrier()
            el_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(el_t, op=dist.ReduceOp.MAX)
            el = float(el_t.item())
            e2e = {"value": n_dofs * world * e_steps / el / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": n_dofs * 8 * world, "d2h_bytes_per_step": n_dofs * 8 * world,
                   "steps": e_steps, "checksum": float(h_dst[:: max(1, n_dofs // 4096)].sum())}
            del h_src, h_dst
        else:
            e2e = {"value": None, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "not enough free host memory for %d pinned bricks" % world}

    if rank == 0:
        peak, peak_src = _peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)
        achieved = n_dofs * BYTES_PER_DOF / (k_ms * 1e-3) / 1e9
        name = op.kernel_name
        out = {
            "metric": "advection operator throughput (3D3V, k=3, FP64)", "value": value, "unit": "GDoF/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D3V k=3 FP64 advection apply, Cartesian periodic, %d^6 cells (%.3g DoFs) per GPU, skew 0.5, ECL" % (args.cells, n_dofs),
                       "cells_global": nglob, "cells_per_gpu": list(part.n_cells), "gpu_grid": p, "halo_bytes_sent_per_gpu_per_step": halo_bytes, "halo": halo_mode, "overlap": (args.overlap if peer is not None else "parts") if world > 1 else "none", "l2": "vectors (%.1f GiB each) are larger than L2; no flush needed" % (n_dofs * 8 / 2**30),
                       "kernel": name},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(name),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": n_dofs * BYTES_PER_DOF, "kernel_ms": k_ms},
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if e2e is not None:
            out["e2e"] = e2e
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
