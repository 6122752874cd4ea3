"""hd_multi_*: the single-process multi-GPU entry points of the C ABI (C++ hosts; include/hyperdeal_b200.h) against the
single-GPU entry points on the same lattice — operator and complete LSRK steps, 3D3V k=3 (three-round kernel, fused stage)
and a 2D2V Dirichlet lattice (generic kernel).  Needs >= 2 GPUs; the single-GPU side is pinned against the oracle by
tests/test_apply_gpu.py / test_pipeline_gpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


VEL6 = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
CASES = [
    # (n_gpus, dim_x, dim_v, degree, n_cells_global, grid, periodic, velocity)
    (2, 3, 3, 3, (3, 2, 4, 2, 2, 2), (1, 1, 2, 1, 1, 1), True, VEL6),
    (2, 3, 3, 3, (4, 2, 2, 2, 2, 2), (2, 1, 1, 1, 1, 1), True, VEL6),  # rows of cells along x_0 cut
    (2, 3, 3, 3, (2, 2, 2, 2, 2, 2), (1, 1, 1, 1, 1, 2), True, tuple(-v for v in VEL6)),
    (4, 3, 3, 3, (3, 2, 4, 2, 2, 2), (1, 2, 2, 1, 1, 1), True, VEL6),
    (4, 3, 3, 3, (3, 2, 4, 2, 2, 2), (1, 1, 4, 1, 1, 1), True, VEL6),  # bricks one cell thick
    (8, 3, 3, 3, (4, 2, 4, 2, 2, 2), (1, 2, 4, 1, 1, 1), True, VEL6),  # the x24 layout of bench.py at 8 GPUs
    (2, 2, 2, 3, (4, 4, 4, 4), (1, 2, 1, 1), False, (1.0, 0.15, -0.05, 0.1)),
    (4, 2, 2, 3, (4, 4, 4, 4), (1, 2, 1, 2), False, (1.0, 0.15, -0.05, 0.1)),
    (2, 2, 2, 2, (4, 6, 4, 2), (1, 2, 1, 1), True, (1.0, 0.15, -0.05, 0.1)),
]


# fused = the operator kernels pack and send the halo themselves (3D3V degree-3 FP64 lattices; the default there);
# HD_MULTI_FUSED=0 = pack kernels + CUDA events (every other lattice, and the cross-check)
@pytest.mark.parametrize("fused", [True, False], ids=["fused_halo", "pack_kernels"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dd%dv_k%d_%s%s" % (c[0], c[1], c[2], c[3], "x".join(map(str, c[5])), "" if c[6] else "_dirichlet"))
def test_multi_matches_single(case, fused, monkeypatch):
    from hyperdeal_b200 import api

    if not fused:
        if not (case[1] == 3 and case[3] == 3):
            pytest.skip("the fused halo only exists for 3D3V degree 3: same code path as the other parametrisation")
        monkeypatch.setenv("HD_MULTI_FUSED", "0")

    n, dim_x, dim_v, degree, cells, grid, periodic, vel = case
    if _n_gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    dim = dim_x + dim_v
    left, right = (-1.0,) * dim, (1.0,) * dim
    ctx = api.Context(0)
    mf = api.MatrixFree(ctx, dim_x, dim_v, degree, cells, left, right, periodic=periodic)
    op = api.AdvectionOperation(mf, vel, 0.5)
    if not periodic:
        op.set_dirichlet_builtin(api.FN_HYPERRECTANGLE)
    u = np.random.default_rng(7).standard_normal(mf.n_dofs)
    s1, k1, t1 = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(s1, u)
    op.apply(k1, s1, 0.3)
    ref_apply = mf.copy_out(k1)

    mg = api.MultiGpu(n, dim_x, dim_v, degree, cells, left, right, grid, periodic=periodic)
    try:
        assert mg.n_dofs == mf.n_dofs
        mop = mg.advection(vel, 0.5)
        if not periodic:
            api._check(api.lib().hd_multi_advection_set_dirichlet_builtin(mop, api.FN_HYPERRECTANGLE))
        s, k, t = (mg.initialize_dof_vector() for _ in range(3))
        mg.copy_in(s, u)
        assert np.array_equal(mg.copy_out(s), u)  # scatter / gather round trip
        for _ in range(3):  # ghost buffers alternate: more than two exchanges
            mg.apply(mop, k, s, 0.3)
        got = mg.copy_out(k)
        scale = np.max(np.abs(ref_apply))
        assert np.max(np.abs(got - ref_apply)) <= 1e-13 * scale
        # three complete rk45 steps against the single-GPU fused integrator
        rk1 = api.LowStorageRungeKuttaIntegrator(mf, k1, t1, "rk45")
        rk = mg.lsrk("rk45")
        time, dt = 0.0, 1e-3
        for _ in range(3):
            rk1.perform_time_step(s1, time, dt, op)
            mg.lsrk_step(rk, mop, s, k, t, time, dt)
            time += dt
        mg.synchronize()
        a, b = mf.copy_out(s1), mg.copy_out(s)
        assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(a))
        # norms: the reduction over bricks equals the single-GPU reduction
        n1 = api.VectorTools.norm_and_error(mf, s1, api.FN_HYPERRECTANGLE, time)
        n2 = mg.norm_and_error(s, api.FN_HYPERRECTANGLE, time)
        assert abs(n1[0] - n2[0]) <= 1e-12 * n1[0] and abs(n1[1] - n2[1]) <= 1e-12 * n1[1]
        # interpolation of the built-in initial condition
        mg.interpolate(k, api.FN_HYPERRECTANGLE, 0.25)
        api.VectorTools.interpolate(mf, k1, api.FN_HYPERRECTANGLE, 0.25)
        assert np.array_equal(mg.copy_out(k), mf.copy_out(k1))
    finally:
        mg.close()


def test_multi_rejects_bad_grid():
    from hyperdeal_b200 import api

    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    with pytest.raises(api.HdError):
        api.MultiGpu(2, 2, 2, 3, (3, 3, 3, 3), (-1,) * 4, (1,) * 4, (1, 2, 1, 1))  # 3 cells cannot be cut in two
    with pytest.raises(api.HdError):
        api.MultiGpu(2, 2, 2, 3, (4, 4, 4, 4), (-1,) * 4, (1,) * 4, (2, 2, 1, 1))  # 4 bricks on 2 GPUs
