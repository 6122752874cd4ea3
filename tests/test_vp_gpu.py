"""Validated device pieces of the Vlasov-Poisson rows: hd_velocity_space_integration against the oracle's
restatement of VectorTools::velocity_space_integration (numerics/vector_tools.h:238-315, quad_no_v = 2), and the general-velocity
advection kernel (kernel_vp.cu, hd_advection_set_phase_space_velocity) against the literal oracle with the same velocity tables."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from oracle import oracle_vp as V

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from hyperdeal_b200 import api as A

    return A


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("dx,dv,nc", [(1, 1, (6, 5)), (2, 2, (4, 3, 2, 3)), (3, 3, (2, 2, 1, 2, 2, 2)), (2, 2, (16, 16, 4, 4))])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_velocity_space_integration_matches_oracle(api, ctx, dx, dv, nc, dtype):
    dim = dx + dv
    left, right = (0.0,) * dx + (-6.0,) * dv, (4.0 * np.pi,) * dx + (6.0,) * dv
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, nthreads=2)
    mf = api.MatrixFree(ctx, dx, dv, 3, nc, left, right, dtype=dtype)
    f = np.random.default_rng(3).standard_normal(mf.n_dofs)
    if dtype == np.float32:
        f = f.astype(np.float32).astype(np.float64)
    ref = vp.velocity_space_integration(f)
    assert mf.n_dofs_x == ref.size
    d_f, d_rho = mf.initialize_dof_vector(), mf.initialize_dof_vector_x()
    mf.copy_in(d_f, f)
    for _ in range(2):  # the destination is overwritten, not accumulated
        api.VectorTools.velocity_space_integration(mf, d_rho, d_f)
    rho = mf.copy_out(d_rho, mf.n_dofs_x).astype(np.float64)
    tol = 1e-13 if dtype == np.float64 else 2e-6
    assert np.max(np.abs(rho - ref)) <= tol * np.max(np.abs(ref))
    mf.free_vector(d_f)
    mf.free_vector(d_rho)


def test_density_of_the_landau_initial_condition(api, ctx):
    """interpolated initial condition of examples/vlasov_poisson: the density is (1 + 0.01 cos(x_0 / 2)) times one constant"""
    dx = dv = 2
    nc = (4, 4, 4, 4)
    left, right = (0.0,) * dx + (-6.0,) * dv, (4.0 * np.pi,) * dx + (6.0,) * dv
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, nthreads=2)
    f0 = vp.adv.interpolate(lambda p, t: V.vp_initial_condition(p, dx), 0.0)
    mf = api.MatrixFree(ctx, dx, dv, 3, nc, left, right)
    d_f, d_rho = mf.initialize_dof_vector(), mf.initialize_dof_vector_x()
    mf.copy_in(d_f, f0)
    api.VectorTools.velocity_space_integration(mf, d_rho, d_f)
    rho = mf.copy_out(d_rho, mf.n_dofs_x)
    assert np.max(np.abs(rho - vp.velocity_space_integration(f0))) <= 1e-13 * np.max(np.abs(rho))


def test_general_velocity_kernel_matches_oracle():
    """six cases (1D1V, 2D2V incl. over-integration, 3D3V, FP32) in a child process (tests/vp_kernel_check.py); first GPU run:
    profiles/r01n_vp_kernel_gpu.txt"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_kernel_check.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPK OK") == 24 and "VPK FAIL" not in r.stdout
