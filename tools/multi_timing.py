"""Throughput of the single-process multi-GPU route (hd_multi_*, the C++ hosts' route): operator applications and complete
fused rk45 steps on N GPUs driven from ONE process, 8^6 cells per GPU (weak scaling recipe: x_2, x_1 doubled), wall clock around
a synchronised batch (every brick has its own stream; hd_multi_synchronize waits for all of them).
   python tools/multi_timing.py [n_gpus ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api  # noqa: E402

V = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
GRIDS = {1: (1, 1, 1, 1, 1, 1), 2: (1, 1, 2, 1, 1, 1), 4: (1, 2, 2, 1, 1, 1), 8: (1, 2, 4, 1, 1, 1)}


def run(n):
    grid = GRIDS[n]
    cells = [8, 8, 8, 8, 8, 8]
    # weak scaling: the global lattice grows with the grid (8^6 cells per GPU up to 4 GPUs; 16x8x4 bricks of the x24 layout at 8)
    glob = [8, 8 * (2 if n >= 4 else 1), 8 * (2 if n >= 2 else 1), 8, 8, 8] if n < 8 else [16, 16, 16, 8, 8, 8]
    mg = api.MultiGpu(n, 3, 3, 3, glob, (0.0,) * 6, (1.0,) * 6, grid)
    try:
        op = mg.advection(V, 0.5)
        s, k, t = (mg.initialize_dof_vector() for _ in range(3))
        mg.interpolate(s, api.FN_HYPERRECTANGLE, 0.0)

        def timeit(fn, reps):
            for _ in range(2):
                fn()
            mg.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            mg.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3

        ms = timeit(lambda: mg.apply(op, k, s, 0.0), 10)
        print("hd_multi N=%d grid %s: apply            %8.3f ms  %8.1f GDoF/s  (%s)" % (n, "x".join(map(str, grid)), ms, mg.n_dofs / ms / 1e6, mg.kernel_name(op)), flush=True)
        rk = mg.lsrk("rk45")
        ms = timeit(lambda: mg.lsrk_step(rk, op, s, k, t, 0.0, 1e-6), 3) / 5
        print("hd_multi N=%d grid %s: fused rk45 stage %8.3f ms  %8.1f GDoF/s  (%s)" % (n, "x".join(map(str, grid)), ms, mg.n_dofs / ms / 1e6, mg.kernel_name(op)), flush=True)
    finally:
        mg.close()
    torch.cuda.empty_cache()


if __name__ == "__main__":
    have = torch.cuda.device_count()
    for n in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
        if n <= have:
            run(n)
