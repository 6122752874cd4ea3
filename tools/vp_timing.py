"""Vlasov-Poisson right-hand side on the device (BASELINE.json configs[3]), stage by stage: velocity-space integration
(rho = int f dv), the x-space Poisson solve (Jacobi-preconditioned CG; iterations reported), the general-velocity operator
(kernel_vp.cu) and the LSRK stage update; ms per piece with CUDA events, GDoF/s of the whole stage.
   python tools/vp_timing.py [1d1v|2d2v|3d3v ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api  # noqa: E402

CASES = {
    "1d1v": (1, 1, 3, (2048, 2048)),
    "2d2v": (2, 2, 3, (32, 32, 32, 32)),
    "2d2v_big": (2, 2, 3, (64, 64, 32, 32)),
    "3d3v": (3, 3, 3, (8, 8, 8, 4, 4, 4)),
}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name):
    dx, dv, k, nc = CASES[name]
    dim = dx + dv
    ctx = api.Context(0)
    left, right = (0.0,) * dx + (-6.0,) * dv, (4.0 * np.pi,) * dx + (6.0,) * dv
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right)
    op = api.AdvectionOperation(mf, (1.0,) * dim, 0.0)
    ps = api.PoissonSolver(mf)
    n = k + 1
    ncx = int(np.prod(nc[:dx]))
    a_v = torch.zeros(ncx * n ** dx * dv, dtype=torch.float64, device="cuda")
    d_rho = mf.initialize_dof_vector_x()
    op.set_phase_space_velocity(a_v.data_ptr())
    f, K, sol = (mf.initialize_dof_vector() for _ in range(3))
    # Landau-damping-like initial condition from the built-in function is not available: a smooth random field does for timing
    # (the CG iteration count depends on the right-hand side only weakly)
    rng = np.random.default_rng(3)
    h = rng.standard_normal(mf.n_dofs) * 0.01 + 1.0
    mf.copy_in(f, h)
    mf.copy_in(sol, h)
    del h
    t_rho = timeit(lambda: api.VectorTools.velocity_space_integration(mf, d_rho, f))
    def solve():
        try:
            ps.solve(d_rho, a_v.data_ptr(), rel_tol=1e-7, max_iterations=10000)
        except api.HdError as e:  # reported below through last_solve
            print("   (solve: %s)" % e)

    ps.solve(d_rho, a_v.data_ptr(), rel_tol=1e-7, max_iterations=10000)
    it_cold, res_cold = ps.last_solve  # first solve from a zero potential; the timed ones start from the previous potential
    t_poisson = timeit(solve)
    t_apply = timeit(lambda: op.apply(K, f, 0.0))
    L = api.lib()
    t_update = timeit(lambda: api._check(L.hd_lsrk_stage_update(mf._h, api.c_void_p(sol), api.c_void_p(f), api.c_void_p(K), 1e-6, 1e-6)))
    Ki, Ti = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    rk = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk45")
    mf.copy_in(Ti, np.full(mf.n_dofs, 1.0))
    t_fused = timeit(lambda: api._check(L.hd_lsrk_stage_fused(rk._h, op._h, 1, api.c_void_p(sol), api.c_void_p(Ti), api.c_void_p(Ki), None, 0.0, 1e-9)))
    total = t_rho + t_poisson + t_apply + t_update
    total_fused = t_rho + t_poisson + t_fused
    it, res = ps.last_solve
    es = 8
    print("VP %-9s %s cells, %.3e DoFs, kernel %s" % (name, "x".join(map(str, nc)), mf.n_dofs, op.kernel_name))
    print("   rho = int f dv        %9.3f ms  (%5.0f GB/s read)" % (t_rho, mf.n_dofs * es / t_rho / 1e6))
    print("   Poisson solve         %9.3f ms  (warm start: %d CG iterations, relative residual %.2e; cold start: %d iterations; %d x-space DoFs)" % (t_poisson, it, res, it_cold, mf.n_dofs_x))
    print("   operator (general a)  %9.3f ms  (%6.1f GDoF/s, %5.0f GB/s algorithmic)" % (t_apply, mf.n_dofs / t_apply / 1e6, mf.n_dofs * 2 * es / t_apply / 1e6))
    print("   LSRK stage update     %9.3f ms  (%5.0f GB/s)" % (t_update, mf.n_dofs * 4 * es / t_update / 1e6))
    print("   whole stage           %9.3f ms  = %6.2f GDoF/s" % (total, mf.n_dofs / total / 1e6))
    print("   operator + update, one kernel (hd_lsrk_stage_fused) %9.3f ms (%5.0f GB/s algorithmic at 32 B/DoF)" % (t_fused, mf.n_dofs * 4 * es / t_fused / 1e6))
    print("   whole stage, fused    %9.3f ms  = %6.2f GDoF/s" % (total_fused, mf.n_dofs / total_fused / 1e6), flush=True)


if __name__ == "__main__":
    for name in sys.argv[1:] or ["1d1v", "2d2v", "3d3v"]:
        run(name)
        torch.cuda.empty_cache()
