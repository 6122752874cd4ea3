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
  cpu_baseline  the CPU restatement of the reference's literal ECL algorithm (oracle/hd_ecl_simd.cpp, "port":
              vectorised over cells like the reference, one pinned thread per core) on the host cores, a bounded
              number of applies on the SAME 8^6-cell lattice (a smaller one only if the host lacks the memory)
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


def _traffic(kernel_name, n_gpus=1):
    """measured DRAM bytes of one launch (ncu), keyed by kernel name; the N > 1 launches (ghost reads, interior/boundary
    phases, halo stores) have their own entries "<kernel>@N<n>" — never the N = 1 constant"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_name if n_gpus == 1 else "%s@N%d" % (kernel_name, n_gpus))
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

    def begin(self):
        """forget what was sampled so far: the timed region starts now (the polling thread keeps running)"""
        del self.sm[:]
        self.reasons.clear()
        del self.samples[:]

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


CPU_NOTE = ("CPU restatement of hyper.deal's literal ECL algorithm (advection_operation.h:221-566), vectorised over 8 cells per "
            "SIMD batch like the reference's VectorizedArray<double>, one pinned thread per core (oracle/hd_ecl_simd.cpp); "
            "the hyper.deal binary itself needs deal.II + MPI, which this image does not have")


def _cpu_problem():
    """(FastECL, src, dst, same_config): the benchmark lattice itself — 8^6 cells, 8 GiB per vector — if the host has the
    memory for two such vectors, else the largest lattice of the same family that fits (halved along v)."""
    import numpy as np
    import psutil

    from oracle import oracle as O

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nc = [CELLS_PER_DIR] * 6
    avail = psutil.virtual_memory().available
    d = 5
    while 2 * 8 * 4096 * int(np.prod(nc)) > 0.6 * avail and int(np.prod(nc)) > 4096:
        nc[d] //= 2
        d = 3 + (d - 3 - 1) % 3
    ecl = O.FastECL(nc, (0.0,) * 6, (1.0,) * 6, VELOCITY, skew=SKEW, nthreads=cores, pin=True)
    block = np.random.default_rng(20240229).standard_normal(1 << 22)
    src = np.empty(ecl.ndofs)
    for i in range(0, ecl.ndofs, block.size):  # (a standard normal block repeated: filling 10^9 values from the generator takes longer than the run)
        src[i:i + block.size] = block[: min(block.size, ecl.ndofs - i)]
    dst = np.empty_like(src)
    return ecl, src, dst, nc == [CELLS_PER_DIR] * 6, cores


def cpu_baseline(seconds_hint=15.0):
    """the reference algorithm on the host cores, on the benchmark lattice (bounded number of applies)"""
    ecl, src, dst, same, cores = _cpu_problem()
    ecl.apply(src, dst)  # warm-up (also builds the library)
    t0, n = time.perf_counter(), 0
    while True:
        ecl.apply(src, dst)
        n += 1
        el = time.perf_counter() - t0
        if el > seconds_hint or n >= 20:
            break
    return {"value": ecl.ndofs * n / el / 1e9, "unit": "GDoF/s", "cores": cores, "kind": "port", "same_config": bool(same),
            "sample": "%d applies of the %s-cell (%.3g DoF) 3D3V k=3 FP64 lattice, %d threads; %s" % (n, "x".join(str(c) for c in ecl.n_cells), ecl.ndofs, cores, CPU_NOTE),
            "seconds": el}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    ecl, src, dst, same, cores = _cpu_problem()
    for _ in range(max(1, min(args.warmup, 2))):
        ecl.apply(src, dst)
    t0 = time.perf_counter()
    for _ in range(steps):
        ecl.apply(src, dst)
    el = time.perf_counter() - t0
    v = ecl.ndofs * steps / el / 1e9
    cells = "x".join(str(c) for c in ecl.n_cells)
    sample = "each step = one apply on the %s-cell (%.3g DoF) 3D3V k=3 FP64 lattice, %d threads; %s" % (cells, ecl.ndofs, cores, CPU_NOTE)
    print(json.dumps({
        "impl": "reference", "metric": "advection operator throughput (3D3V, k=3, FP64)", "value": v, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D3V k=3 FP64 advection apply, Cartesian periodic, %s cells (%.3g DoFs), skew 0.5, ECL (on the host cores)" % (cells, ecl.ndofs),
                   "same_config": bool(same)},
        "cpu_baseline": {"value": v, "unit": "GDoF/s", "cores": cores, "kind": "port", "same_config": bool(same), "sample": sample},
        "e2e": {"value": v, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_aux_workload(args):
    """Secondary bench lines (N = 1), same JSON contract, for the other BASELINE.json configurations:
         k5f32  — configs[2]: 3D3V advection, degree 5, FP32, periodic, 6x6x6x4x4x4 cells (6.4e8 DoFs)
         vp2d2v — configs[3]: one Vlasov-Poisson LSRK stage in 2D2V (degree 3, FP64, 32^4 cells): rho = int f dv, Poisson solve
                  (CG), grad(phi) table, general-velocity operator + stage update in ONE kernel
         lsrk   — the headline lattice (3D3V, degree 3, FP64, 8^6 cells) inside its time integrator: one low-storage Runge-Kutta
                  stage = operator + both vector updates in ONE kernel (hd_lsrk_step, rk45; a step = one stage, 32 B/DoF)
    `value` = device-resident throughput; `e2e` = the same step with the vectors starting and ending in pinned host memory."""
    import numpy as np
    import torch

    from hyperdeal_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    ctx = api.Context(0)
    peak, peak_src = _peaks()
    vp = args.workload == "vp2d2v"
    lsrk = args.workload == "lsrk"
    if lsrk:
        dx, k, nc, np_dtype, t_dtype = 3, DEGREE, (CELLS_PER_DIR,) * 6, np.float64, torch.float64
        left, right = (0.0,) * 6, (1.0,) * 6
        bytes_per_dof, skew = 32, SKEW
    elif vp:
        dx, k, nc, np_dtype, t_dtype = 2, 3, (32, 32, 32, 32), np.float64, torch.float64
        left, right = (0.0,) * 2 + (-6.0,) * 2, (4.0 * np.pi,) * 2 + (6.0,) * 2
        bytes_per_dof, skew = 32, 0.0  # read Ti and sol, write sol and Ti_next (the rho pass re-reads Ti: + 8, counted as overhead)
    else:
        dx, k, nc, np_dtype, t_dtype = 3, 5, (6, 6, 6, 4, 4, 4), np.float32, torch.float32
        left, right = (0.0,) * 6, (1.0,) * 6
        bytes_per_dof, skew = 8, SKEW
    dim = 2 * dx
    mf = api.MatrixFree(ctx, dx, dx, k, nc, left, right, dtype=np_dtype)
    op = api.AdvectionOperation(mf, VELOCITY[:dim] if not vp else (1.0,) * dim, skew)
    n = mf.n_dofs
    torch.manual_seed(11)
    src = torch.empty(n, dtype=t_dtype, device="cuda").normal_()
    dst = torch.zeros_like(src)
    if vp:
        src.mul_(0.01).add_(1.0)
        ps = api.PoissonSolver(mf)
        a_v = torch.zeros(int(np.prod(nc[:dx])) * (k + 1) ** dx * dx, dtype=torch.float64, device="cuda")
        d_rho = mf.initialize_dof_vector_x()
        op.set_phase_space_velocity(a_v.data_ptr())
        sol, ti_next = src.clone(), torch.zeros_like(src)
        rk = api.LowStorageRungeKuttaIntegrator(mf, ti_next.data_ptr(), src.data_ptr(), "rk45")
        L = api.lib()
        cg = []

        def step():
            api.VectorTools.velocity_space_integration(mf, d_rho, src.data_ptr())
            cg.append(ps.solve(d_rho, a_v.data_ptr(), rel_tol=1e-7, max_iterations=10000))
            api._check(L.hd_lsrk_stage_fused(rk._h, op._h, 1, api.c_void_p(sol.data_ptr()), api.c_void_p(src.data_ptr()), api.c_void_p(ti_next.data_ptr()), None, 0.0, 1e-9))
    elif lsrk:
        ki = torch.zeros_like(src)
        rk = api.LowStorageRungeKuttaIntegrator(mf, ki.data_ptr(), dst.data_ptr(), "rk45")
        n_stages = rk.n_stages()

        def step():  # one complete rk45 step = n_stages fused stages; reported per stage below
            rk.perform_time_step(src.data_ptr(), 0.0, 1e-6, op)
    else:

        def step():
            op.apply(dst.data_ptr(), src.data_ptr(), 0.0)

    per_step = n_stages if lsrk else 1
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = op.launch_count
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps / per_step
    # the dominant kernel alone (events on the library's stream)
    k_ms = []
    for _ in range(5):
        ctx.timer_start()
        if lsrk:
            step()
        elif vp:
            api._check(L.hd_lsrk_stage_fused(rk._h, op._h, 1, api.c_void_p(sol.data_ptr()), api.c_void_p(src.data_ptr()), api.c_void_p(ti_next.data_ptr()), None, 0.0, 1e-9))
        else:
            op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
        k_ms.append(ctx.timer_stop() / per_step)
    k_ms = sum(k_ms) / len(k_ms)
    launches = op.launch_count - launches0
    name = op.kernel_name
    result = sol if vp else (src if lsrk else dst)
    checksum = float(result[:: max(1, n // 65536)].double().abs().sum().item())
    # end to end: vectors start and end in pinned host memory
    es = src.element_size()
    h_in = torch.empty(n, dtype=t_dtype).pin_memory()
    h_out = torch.empty(n, dtype=t_dtype).pin_memory()
    h_in.copy_(src)
    e_steps = 2
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        src.copy_(h_in, non_blocking=True)
        step()
        h_out.copy_(result, non_blocking=True)
        torch.cuda.synchronize()
    el = (time.perf_counter() - t0) / per_step
    achieved = n * bytes_per_dof / (k_ms * 1e-3) / 1e9
    out = {
        "metric": "Vlasov-Poisson LSRK stage throughput (2D2V, k=3, FP64)" if vp else ("fused LSRK stage throughput (3D3V, k=3, FP64)" if lsrk else "advection operator throughput (3D3V, k=5, FP32)"),
        "value": n / (ms * 1e-3) / 1e9, "unit": "GDoF/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if (vp or lsrk) else "f32", "data": "synthetic",
        "config": {"workload": ("2D2V k=3 FP64 Vlasov-Poisson stage (rho, CG field solve, general-velocity operator + LSRK update), %s cells (%.3g DoFs)" if vp else
                                ("3D3V k=3 FP64 low-storage Runge-Kutta stage (rk45; operator + both vector updates in one kernel), %s cells (%.3g DoFs), skew 0.5; a step = one stage" if lsrk else
                                 "3D3V k=5 FP32 advection apply, Cartesian periodic, %s cells (%.3g DoFs), skew 0.5, ECL")) % ("x".join(map(str, nc)), n),
                   "kernel": name, "l2": "vectors (%.1f GiB each) are larger than L2; no flush needed" % (n * es / 2**30)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(name, 1), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": n * bytes_per_dof, "kernel_ms": k_ms,
                     "note": ("FP64-bound, not HBM-bound: ~77 DFMA per DoF (velocity varies inside the cell), see DESIGN.md" if vp else
                              ("32 B/DoF: Ti read, sol read + written, Ti_next written; K is never stored (DESIGN.md section 5)" if lsrk else
                               "global-memory tile kernel (kernel_tile_global.cu): a degree-5 cell and its partial sums do not fit into shared memory together, see DESIGN.md"))},
        "clocks": clocks, "gpu_launches": int(launches) + (args.steps * 2 if vp else 0), "checksum": checksum,
        "e2e": {"value": n * e_steps / el / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": n * es / per_step, "d2h_bytes_per_step": n * es / per_step, "steps": e_steps * per_step},
    }
    if vp:
        out["config"]["cg_iterations_per_solve"] = sum(cg[-args.steps:]) / max(1, args.steps)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="headline", choices=["headline", "k5f32", "vp2d2v", "lsrk"],
                    help="headline = the BASELINE.json metric (default); k5f32 / vp2d2v = secondary lines for configs[2] / configs[3]; lsrk = one fused "
                         "Runge-Kutta stage on the headline lattice (N = 1)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=CELLS_PER_DIR, help="cells per direction per GPU (default 8 = the BASELINE workload)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 two-role pipelined 3D3V kernel, 6 three-round 3D3V kernel")
    ap.add_argument("--layout", default="x24", choices=["x", "xv", "x24"],
                    help="how the N = 8 lattice (16,16,16,8,8,8) is cut: x = x_2,x_1,x_0 in two (bricks of 8^6 cells); xv = x_2,x_1,v_2 in two (16x8x8x8x8x4); "
                         "x24 (default) = x_2 in four, x_1 in two (16x8x4x8x8x8): two cut directions instead of three, rows of cells (along x_0) stay whole")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="N > 1: direct peer-memory stores over NVLink (default) or NCCL send/recv")
    ap.add_argument("--overlap", default="kernel", choices=["kernel", "parts"], help="N > 1: one launch that waits for the halo flag in-kernel (default) or interior/boundary launches")
    ap.add_argument("--sustain", type=float, default=3.0, help="also time the step back to back for at least this many seconds (0 = off)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "headline":
        if rank == 0:
            run_aux_workload(args)
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

    ctx = api.Context(local_rank)

    def make_partition(cells):
        # global lattice of the weak-scaling recipe (x_2, x_1, x_0 doubled in turn); how it is cut into equal bricks is ours
        recipe = BrickPartition(world, rank, list(cells), split_order=(2, 1, 0))
        if args.layout == "x" or world < 8:
            return recipe
        # never cut x_0 (rows of cells are walked along x_0): the third cut goes through v_2 ("xv") or through x_2 again ("x24")
        cut = BrickPartition(world, rank, list(cells), split_order=(2, 1, 5) if args.layout == "xv" else (2, 1, 2))
        nloc = [max(1, g // c) for g, c in zip(recipe.n_cells_global, cut.grid)]
        part = BrickPartition(world, rank, nloc, grid=cut.grid)
        return part

    class Problem:
        """one brick per rank of a periodic lattice + everything one operator application needs (halo plan, peer buffers)"""

        def __init__(self, part):
            self.part = part
            self.mf = api.MatrixFree(ctx, 3, 3, DEGREE, part.n_cells, (0.0,) * 6, (1.0,) * 6, n_cells_global=part.n_cells_global, cell_offset=part.cell_offset,
                                     side_kind=part.side_kind)
            self.op = api.AdvectionOperation(self.mf, VELOCITY, SKEW)
            self.op.set_kernel(args.kernel)
            mf, op = self.mf, self.op
            self.n_dofs = mf.n_dofs
            self.src = torch.empty(self.n_dofs, dtype=torch.float64, device="cuda")
            self.dst = torch.zeros(self.n_dofs, dtype=torch.float64, device="cuda")
            halo = mf.halo_total
            self.send = torch.empty(max(halo, 16), dtype=torch.float64, device="cuda")
            self.ghost = torch.zeros(max(halo, 16), dtype=torch.float64, device="cuda")
            offsets = {(d, s): mf.halo_offset(d, s) for d in range(6) for s in range(2)}
            sizes = {(d, s): mf.ghost_size(d, s) for d in range(6) for s in range(2)}
            # upwind flux: only the inflow ghost side of every cut direction is read -> only that one is exchanged
            self.exch = HaloExchange(part, offsets, sizes, op.ghost_sides())
            self.send_mask = self.exch.send_mask()
            self.halo_bytes = self.exch.bytes_per_exchange[0] * 8
            # N > 1: direct NVLink variant (sender CTAs store into the neighbours' ghost buffers) unless --halo nccl
            self.peer, self.halo_mode = None, "none"
            if world > 1:
                self.halo_mode = "nccl"
                if args.halo == "peer":
                    from hyperdeal_b200.partition import PeerHaloExchange

                    try:
                        self.peer = PeerHaloExchange(part, offsets, sizes, halo, op.ghost_sides(), torch.device("cuda", local_rank), ctx=ctx)
                        self.halo_mode = "peer"
                    except Exception as e:  # symmetric memory unavailable on this box: NCCL send/recv
                        if rank == 0:
                            sys.stderr.write("bench: peer-memory halo unavailable (%s: %s), using NCCL send/recv\n" % (type(e).__name__, e))
                    ok = torch.tensor([1 if self.peer is not None else 0], device="cuda")
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    if ok.item() == 0:
                        self.peer, self.halo_mode = None, "nccl"
            self.main_stream = torch.cuda.current_stream()
            self.side_stream = torch.cuda.Stream() if world > 1 else None
            self.ev_src, self.ev_halo = torch.cuda.Event(), torch.cuda.Event()
            self.fused = self.peer is not None and args.overlap == "kernel"

        def step(self):
            """one operator application; N > 1: start the halo -> interior cells -> halo arrived -> boundary cells
            (the reference's overlapping levels, matrix_free.templates.h:1516-1566)"""
            mf, op, src, dst, peer = self.mf, self.op, self.src, self.dst, self.peer
            if world == 1:
                op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
                return
            if self.fused:
                # ONE launch: the sender CTAs store the boundary layers into the neighbours' ghost buffers over NVLink while the
                # others do the interior cells; the boundary layer starts when the neighbours' arrival counters are complete
                g, sends, counters, epoch = peer.begin_fused(ctx, op)
                op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), sends, counters, epoch)
                peer.consumed(ctx)
                return
            self.ev_src.record(self.main_stream)
            self.side_stream.wait_event(self.ev_src)
            ctx.set_stream(self.side_stream.cuda_stream)
            m = 0
            with torch.cuda.stream(self.side_stream):
                if peer is not None:
                    g, m = peer.start(mf, ctx, src.data_ptr())
                else:
                    mf.halo_pack(src.data_ptr(), self.send.data_ptr(), send_mask=self.send_mask)
                    HaloExchange.finish(self.exch.start(self.send, self.ghost))
                    g = self.ghost
                    self.ev_halo.record(self.side_stream)
            ctx.set_stream(self.main_stream.cuda_stream)
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_INTERIOR)
            if peer is not None:
                peer.wait_ready(ctx, m)
            else:
                self.main_stream.wait_event(self.ev_halo)
            op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_BOUNDARY)
            if peer is not None:
                peer.consumed(ctx)

    # ---- N > 1: parity of THIS code path (same layout, halo mode and overlap) on a small lattice, before anything is timed:
    # every rank's brick against the whole small lattice applied as one periodic brick on its own GPU (the single-brick
    # operator is pinned against the oracle by tests/test_apply_gpu.py).  Random field; all ranks must agree to 1e-13.
    parity = None
    if world > 1:
        full_part = make_partition([args.cells] * 6)
        small_local = [max(1, c // 4) for c in full_part.n_cells]
        spart = BrickPartition(world, rank, small_local, grid=full_part.grid)
        sp = Problem(spart)
        mf_all = api.MatrixFree(ctx, 3, 3, DEGREE, spart.n_cells_global, (0.0,) * 6, (1.0,) * 6)
        op_all = api.AdvectionOperation(mf_all, VELOCITY, SKEW)
        op_all.set_kernel(args.kernel)
        u = np.random.default_rng(20240229).standard_normal(mf_all.n_dofs)
        a_src = torch.from_numpy(u).cuda()
        a_dst = torch.zeros_like(a_src)
        op_all.apply(a_dst.data_ptr(), a_src.data_ptr(), 0.0)
        shape = tuple(reversed(spart.n_cells_global)) + (4096,)
        sl = tuple(slice(spart.cell_offset[d], spart.cell_offset[d] + spart.n_cells[d]) for d in reversed(range(6)))
        expect = a_dst.reshape(shape)[sl].contiguous().reshape(-1)
        sp.src.copy_(a_src.reshape(shape)[sl].contiguous().reshape(-1))
        rel = 0.0
        for it in range(3):  # (three applications: both ghost buffers of the double buffer and the epoch counters are exercised)
            sp.dst.fill_(float("nan"))
            sp.step()
            torch.cuda.synchronize()
            rel = max(rel, float(((sp.dst - expect).abs().max() / expect.abs().max()).item()) if bool(torch.isfinite(sp.dst).all()) else float("inf"))
        if sp.fused and sp.op.overlap_timed_out():
            rel = float("inf")
        t = torch.tensor([rel], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity = {"parity_rel": float(t.item()), "tolerance": 1e-13, "layout": args.layout if world >= 8 else "recipe", "gpu_grid": list(spart.grid),
                  "cells_per_gpu": list(spart.n_cells), "cells_global": list(spart.n_cells_global), "halo": sp.halo_mode,
                  "overlap": ("kernel" if sp.fused else "parts"), "kernel": sp.op.kernel_name, "field": "standard normal, seed 20240229",
                  "against": "the whole small lattice as one periodic brick on every rank's own GPU"}
        if not (parity["parity_rel"] <= 1e-13):
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed", **parity}))
            dist.barrier()
            dist.destroy_process_group()
            raise SystemExit(3)
        del sp, mf_all, op_all, a_src, a_dst, expect
        torch.cuda.empty_cache()

    part = make_partition([args.cells] * 6)
    nglob, p = list(part.n_cells_global), list(part.grid)
    prob = Problem(part)
    mf, op, src, dst, ghost, peer = prob.mf, prob.op, prob.src, prob.dst, prob.ghost, prob.peer
    n_dofs, halo_bytes, halo_mode, send_mask = prob.n_dofs, prob.halo_bytes, prob.halo_mode, prob.send_mask
    step = prob.step
    # synthetic data: a standard normal field (every face term and every rounding path is exercised; zeros — what the
    # reference benchmark streams — or a smooth wave would let a wrong kernel look right in the checksums below)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    src.normal_(generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # The clock sampler is set up BEFORE the barrier: nvmlInit takes tens of milliseconds on an 8-GPU box and a different
    # time in every process; done after the barrier it let the ranks enter the timed loop milliseconds apart, and since
    # every kernel waits for its neighbours' halo the whole skew landed in the 10 timed steps (N = 8: 5.4 instead of 4.9 ms).
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler is not None:
        sampler.begin()
    launches0 = op.launch_count
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
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        # per-launch kernel time for the roofline entry: one un-split launch on this rank's brick (outside the timed region)
        ctx.timer_start()
        op.apply(dst.data_ptr(), src.data_ptr(), 0.0, ghosts=ghost.data_ptr())
        kernel_ms = [ctx.timer_stop()]
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    if world > 1 and not (peer is not None and args.overlap == "kernel"):
        launches += args.steps * sum(send_mask)  # pack kernels
    value = n_dofs * world * args.steps / (total_ms * 1e-3) / 1e9
    # result fingerprint of the device-resident run (random field): compared with the end-to-end result below
    dst_sample = dst[:: max(1, n_dofs // 65536)].clone()
    checksum = float(dst_sample.abs().sum().item())

    # ---- sustained regime: the same step back to back for >= --sustain seconds (an LSRK run lives there; the burst number
    # above is taken at boost clocks)
    sustained = None
    if args.sustain > 0:
        n_s = max(args.steps, int(args.sustain * 1e3 / max(total_ms / args.steps, 1e-3)) + 1)
        s_sampler = ClockSampler(local_rank) if rank == 0 else None
        if s_sampler is not None:
            s_sampler.start()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if s_sampler is not None:
            s_sampler.begin()
        es0.record()
        for _ in range(n_s):
            step()
        es1.record()
        barrier()
        s_clocks = s_sampler.stop() if rank == 0 else None
        s_ms = torch.tensor([es0.elapsed_time(es1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(s_ms, op=dist.ReduceOp.MAX)
        s_total = float(s_ms.item())
        sustained = {"value": n_dofs * world * n_s / (s_total * 1e-3) / 1e9, "unit": "GDoF/s", "steps": n_s, "seconds": s_total * 1e-3, "ms_per_step": s_total / n_s,
                     "clocks": s_clocks}

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
        h_sample = h_dst[:: max(1, n_dofs // 65536)].cuda()
        e2e = {"value": n_dofs * e_steps / el / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": n_dofs * 8, "d2h_bytes_per_step": n_dofs * 8, "steps": e_steps,
               "checksum": float(h_sample.abs().sum().item()), "max_abs_dev_vs_resident": float((h_sample - dst_sample).abs().max().item())}
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
            h_sample = h_dst[:: max(1, n_dofs // 65536)].cuda()
            e2e = {"value": n_dofs * world * e_steps / el / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": n_dofs * 8 * world, "d2h_bytes_per_step": n_dofs * 8 * world,
                   "steps": e_steps, "checksum": float(h_sample.abs().sum().item()), "max_abs_dev_vs_resident": float((h_sample - dst_sample).abs().max().item())}
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
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(name, world),
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this N (profiles/ncu_traffic.json); null = not captured",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": n_dofs * BYTES_PER_DOF, "kernel_ms": k_ms},
            "clocks": clocks, "gpu_launches": int(launches), "checksum": checksum,
        }
        if sustained is not None:
            out["sustained"] = sustained
        if parity is not None:
            out["parity"] = parity
            out["parity_rel"] = parity["parity_rel"]
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
