"""Child process of tests/test_zz_vp_device_gpu.py: the general-velocity kernel (kernel_vp.cu) against the literal oracle with
the same separable velocity tables.  Runs in its own process so that a fault in this not-yet-validated kernel cannot take
the CUDA context of the main test session with it.  Prints one 'VPK OK|FAIL' line per case."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hyperdeal_b200 import api  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import oracle_vp as V  # noqa: E402

CASES = [
    # dx dv cells                   nq    skew  dtype
    (1, 1, (3, 4), None, 0.0, np.float64),
    (1, 1, (2, 3), None, 0.5, np.float64),
    (1, 1, (131, 5), None, 0.3, np.float64),            # more than one CTA of the thread-per-cell kernel, ragged tail
    (2, 2, (2, 3, 2, 2), None, 0.0, np.float64),
    (2, 2, (3, 1, 1, 3), None, 0.5, np.float64),        # odd number of cells: half a warp without a cell
    (2, 2, (5, 4, 3, 4), None, 1.0, np.float64),
    (2, 2, (2, 2, 3, 2), 5, 0.3, np.float64),
    (3, 3, (2, 1, 2, 2, 2, 1), None, 0.0, np.float64),
    (3, 3, (2, 2, 2, 2, 2, 2), None, 0.5, np.float64),
    (3, 3, (1, 2, 1, 2, 1, 2), None, 0.3, np.float32),
    (2, 2, (3, 2, 2, 2), None, 0.0, np.float32),
    (1, 1, (7, 3), None, 0.0, np.float32),
]


def expected_kernel(dx, nq):
    """degree 3 with 4 quadrature points: the register-tile kernels (kernel_vp_tile.cuh); else the generic one"""
    if nq in (None, 4):
        return {1: "vp_tile_1d1v", 2: "vp_tile_2d2v", 3: "vp_tile_3d3v"}[dx]
    return "vp_generic"


def main():
    ctx = api.Context(0)
    bad = 0
    for dx, dv, nc, nq, skew, dtype in CASES:
        dim = dx + dv
        left, right = (0.0,) * dx + (-1.3,) * dv, (2.0,) * dx + (1.7,) * dv  # v = 0 lies inside a cell
        vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, n_points=nq, nthreads=2)
        rng = np.random.default_rng(17)
        a_v = rng.standard_normal(vp.adv.a_v_table.shape)
        orc = O.Oracle(vp.mesh, 3, nq=nq, skew=skew, a_x_table=vp.v_at_q, a_v_table=a_v, nthreads=2)
        f = rng.standard_normal(orc.ndofs)
        if dtype == np.float32:
            f = f.astype(np.float32).astype(np.float64)
        ref = orc.apply(f)
        mf = api.MatrixFree(ctx, dx, dv, 3, nc, left, right, n_points=nq, dtype=dtype)
        op = api.AdvectionOperation(mf, (1.0,) * dim, skew)
        import torch

        d_av = torch.from_numpy(np.ascontiguousarray(a_v)).cuda()
        op.set_phase_space_velocity(d_av.data_ptr())
        d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.copy_in(d_src, f)
        tol = 1e-12 if dtype == np.float64 else 1e-5
        for choice in (0, 1):  # automatic choice, then the generic kernel
            op.set_kernel(choice)
            mf.copy_in(d_dst, np.zeros(orc.ndofs))
            op.apply(d_dst, d_src, 0.0)
            out = mf.copy_out(d_dst).astype(np.float64)
            rel = float(np.max(np.abs(out - ref)) / np.max(np.abs(ref)))
            want = expected_kernel(dx, nq) if choice == 0 else "vp_generic"
            ok = rel <= tol and op.kernel_name == want
            bad += not ok
            print("VPK %s dx=%d dv=%d cells=%s nq=%s skew=%g %s kernel=%s rel=%.3e" % ("OK" if ok else "FAIL", dx, dv, nc, nq, skew, np.dtype(dtype).name, op.kernel_name, rel), flush=True)
        op.set_kernel(0)
        # back to the constant velocity: the shipped kernels again
        op.set_phase_space_velocity(None)
        op.apply(d_dst, d_src, 0.0)
        assert not op.kernel_name.startswith("vp_")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
