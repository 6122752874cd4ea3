"""Child process of tests/test_zz_vp_device_gpu.py: the Vlasov-Poisson right-hand side assembled from the device pieces
(hd_velocity_space_integration -> hd_poisson_solve -> hd_advection_set_phase_space_velocity -> hd_advection_apply) against the
oracle (oracle/oracle_vp.py), and the reference's 2D2V Landau-damping golden run with it (diagnostics evaluated by the oracle's
functions on copies of the device vectors).  Own process: none of this device code has been validated yet."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hyperdeal_b200 import api  # noqa: E402
from oracle import oracle_vp as V  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    bad = 0
    ctx = api.Context(0)
    dx = dv = 2
    nc = (4, 4, 4, 4)
    left, right = (0.0,) * dx + (-6.0,) * dv, (4.0 * np.pi,) * dx + (6.0,) * dv
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, nthreads=4)
    f0 = vp.adv.interpolate(lambda p, t: V.vp_initial_condition(p, dx), 0.0)
    mf = api.MatrixFree(ctx, dx, dv, 3, nc, left, right)
    op = api.AdvectionOperation(mf, (1.0,) * 4, 0.0)
    ps = api.PoissonSolver(mf)
    a_v = torch.zeros(vp.adv.a_v_table.size, dtype=torch.float64, device="cuda")
    d_rho = mf.initialize_dof_vector_x()
    op.set_phase_space_velocity(a_v.data_ptr())

    def rhs(src, dst, t):
        api.VectorTools.velocity_space_integration(mf, d_rho, src)
        ps.solve(d_rho, a_v.data_ptr(), rel_tol=1e-11, max_iterations=2000)
        op.apply(dst, src, t)

    # ---- one right-hand side
    d_f, d_k = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_f, f0)
    rhs(d_f, d_k, 0.0)
    ref = vp.rhs(f0)
    got = mf.copy_out(d_k)
    rel = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    grad = a_v.cpu().numpy().reshape(vp.adv.a_v_table.shape)
    relg = float(np.max(np.abs(grad - vp.adv.a_v_table)) / np.max(np.abs(vp.adv.a_v_table)))
    ok = rel <= 1e-9 and relg <= 1e-8
    bad += not ok
    print("VPS %s one rhs: rel=%.3e  grad(phi) rel=%.3e" % ("OK" if ok else "FAIL", rel, relg), flush=True)
    if not ok:
        sys.exit(1)  # no point in the long run

    # ---- fused stages (hd_lsrk_stage_fused with the field refreshed per stage) = the unfused std::function structure
    def refresh(src, t):
        api.VectorTools.velocity_space_integration(mf, d_rho, src)
        ps.solve(d_rho, a_v.data_ptr(), rel_tol=1e-11, max_iterations=2000)

    s1, s2, Ki, Ti = (mf.initialize_dof_vector() for _ in range(4))
    mf.copy_in(s1, f0)
    mf.copy_in(s2, f0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk45")
    for step in range(2):
        integ.perform_time_step(s1, step * 1e-2, 1e-2, rhs)
        integ.perform_time_step_staged(s2, step * 1e-2, 1e-2, op, refresh)
    a, b = mf.copy_out(s1), mf.copy_out(s2)
    rel = float(np.max(np.abs(a - b)) / np.max(np.abs(a)))
    ok = rel <= 1e-13
    bad += not ok
    print("VPS %s fused stages against the unfused integrator: rel=%.3e (kernel %s)" % ("OK" if ok else "FAIL", rel, op.kernel_name), flush=True)
    for v in (s1, s2, Ki, Ti):
        api._check(api.lib().hd_vector_free(mf._h, api.c_void_p(v)))

    # ---- the golden run (rk45, 104 steps): fused stages with the device right-hand side
    rows, _ = V.run_vlasov_poisson_example(os.path.join(GOLDEN, "vp_2D_2D_k3.hyperrectangle_01.json"), n_points=4, nthreads=4, max_steps=0)
    gold = V.parse_vp_golden(os.path.join(GOLDEN, "vp_2D_2D_k3.hyperrectangle_01.out"))
    T, n_steps = 0.5, 104
    dt = T / n_steps
    sol, Ki, Ti = mf.initialize_dof_vector(), mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(sol, f0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk45")
    out_rows = []
    for step in range(1, n_steps + 1):
        integ.perform_time_step_staged(sol, (step - 1) * dt, dt, op, refresh)
        t = step * dt
        if int((t + 1e-11) / 0.1) != int((t + 1e-11 - dt) / 0.1):
            f = mf.copy_out(sol)
            vp.potential = np.zeros_like(vp.potential)
            vp.adv.a_v_table[...] = a_v.cpu().numpy().reshape(vp.adv.a_v_table.shape)
            g = vp.adv.a_v_table
            jxw = np.kron(vp.b.w * vp.h[1], vp.b.w * vp.h[0])
            en = [float(np.sum(g[:, :, d] ** 2 * jxw[None, :])) for d in range(2)]
            out_rows.append([t] + en + vp.phase_space_diagnostics(f))
            # the same numbers from the device diagnostics
            dev = api.VectorTools.field_energy(mf, a_v.data_ptr()) + api.VectorTools.phase_space_diagnostics(mf, sol)
            ref_row = out_rows[-1][1:]
            okd = all(abs(a - b) <= 1e-11 * max(abs(b), 1.0) for a, b in zip(dev, ref_row))
            bad += not okd
            print("VPD %s device diagnostics at t=%.3f" % ("OK" if okd else "FAIL", t), flush=True)
    for r, g in zip(out_rows, gold[1:]):
        ok = abs(r[1] - g[1]) <= 1e-7 * g[1] and abs(r[3] - g[3]) <= 1e-12 * g[3] and abs(r[4] - g[4]) <= 1e-10 * g[4] and abs(r[5] - g[5]) <= 1e-10 * g[5]
        bad += not ok
        print("VPS %s t=%.3f energy %.10e (gold %.10e) mass %.12e l2 %.10e kinetic %.10e" % ("OK" if ok else "FAIL", r[0], r[1], g[1], r[3], r[4], r[5]), flush=True)
    sys.exit(1 if bad or len(out_rows) != 5 else 0)


if __name__ == "__main__":
    main()
