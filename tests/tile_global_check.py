"""Child process of tests/test_zz_vp_device_gpu.py: the global-memory tile kernel (kernel_tile_global.cu, hd_advection_set_kernel 5)
against the literal oracle — degree 5 (the 3D3V FP32 case of BASELINE.json configs[2] and smaller relatives) and degree 3, incl.
the fused LSRK step.  Own process: this kernel has not run on a GPU yet."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hyperdeal_b200 import api  # noqa: E402
from oracle import oracle as O  # noqa: E402

VEL = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
CASES = [
    # dx dv cells                 k  skew dtype
    (3, 3, (2, 1, 1, 1, 1, 2), 5, 0.5, np.float32),
    (3, 3, (2, 1, 1, 1, 1, 2), 5, 0.5, np.float64),
    (2, 2, (2, 3, 2, 2), 5, 0.0, np.float64),
    (1, 1, (3, 2), 5, 0.5, np.float64),
    (3, 3, (2, 2, 1, 2, 1, 2), 3, 0.5, np.float64),
    (2, 2, (3, 2, 2, 3), 3, 0.0, np.float32),
]


def main():
    ctx = api.Context(0)
    bad = 0
    for dx, dv, nc, k, skew, dtype in CASES:
        dim = dx + dv
        left, right = (-1.0,) * dim, (1.0,) * dim
        mesh = O.Mesh(dx, dv, nc, left, right, (True,) * dim)
        orc = O.Oracle(mesh, k, skew=skew, velocity=VEL[:dim], nthreads=4)
        f = np.random.default_rng(3).standard_normal(orc.ndofs)
        if dtype == np.float32:
            f = f.astype(np.float32).astype(np.float64)
        ref = orc.apply(f)
        mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right, dtype=dtype)
        op = api.AdvectionOperation(mf, VEL[:dim], skew)
        op.set_kernel(5)
        d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.copy_in(d_src, f)
        op.apply(d_dst, d_src, 0.0)
        out = mf.copy_out(d_dst).astype(np.float64)
        rel = float(np.max(np.abs(out - ref)) / np.max(np.abs(ref)))
        ok = rel <= (1e-12 if dtype == np.float64 else 1e-5) and op.kernel_name == "tile_global"
        # fused LSRK step through the same kernel
        sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
        mf.copy_in(sol, f)
        integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk33")
        integ.perform_time_step(sol, 0.0, 1e-3, op)
        ref2 = O.lsrk_step(lambda v, t: orc.apply(v, t), f, 0.0, 1e-3, "rk33")
        rel2 = float(np.max(np.abs(mf.copy_out(sol).astype(np.float64) - ref2)) / np.max(np.abs(ref2)))
        ok = ok and rel2 <= (1e-12 if dtype == np.float64 else 1e-5)
        bad += not ok
        print("TG %s dx=%d dv=%d cells=%s k=%d %s rel=%.3e fused rk33 rel=%.3e kernel=%s" % ("OK" if ok else "FAIL", dx, dv, nc, k, np.dtype(dtype).name, rel, rel2, op.kernel_name), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
