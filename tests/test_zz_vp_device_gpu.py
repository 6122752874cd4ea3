"""Device runs of the x-space field solve (poisson_x.cu), the Vlasov-Poisson diagnostics (vp_diagnostics.cu), the drivers on top
of them and the global-memory tile kernel (kernel_tile_global.cu).  They were written after round 1's GPU budget was spent and
first ran — and passed — on the driver's box at the end of round 1 (GPUTEST_r01.json); since round 2 they are ordinary,
mandatory tests.  Everything here runs in child processes (a fault cannot touch the CUDA context of the other tests)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drivers():
    from hyperdeal_b200 import build, build_cpp

    build.build()
    return {os.path.basename(p): p for p in build_cpp.build()}


def test_vlasov_poisson_right_hand_side_and_golden_run_on_device():
    """density integration -> field solve -> general-velocity operator, one right-hand side against the oracle and then the
    reference's 2D2V Landau-damping golden (examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_step_check.py")], capture_output=True, text=True, timeout=180)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPS OK") == 7 and "VPS FAIL" not in r.stdout and "VPD FAIL" not in r.stdout


@pytest.mark.parametrize("env", [{}, {"HD_DRIVER_UNFUSED": "1"}], ids=["fused_stages", "reference_call_structure"])
def test_cpp_vlasov_poisson_driver_reproduces_golden(drivers, golden_dir, tmp_path, env):
    """examples/vlasov_poisson re-hosted (hyperdeal_b200/cpp/vlasov_poisson.cc) on the reference's 2D2V Landau-damping case:
    time_history_diagnostic.out against examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out — with the fused
    stages (field refresh + ONE kernel for operator and stage update) and with the reference's std::function call structure"""
    from oracle import oracle_vp as V

    r = subprocess.run([drivers["vlasov_poisson"], os.path.join(golden_dir, "vp_2D_2D_k3.hyperrectangle_01.json")], capture_output=True, text=True, timeout=180, cwd=str(tmp_path),
                       env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr
    rows = V.parse_vp_golden(str(tmp_path / "time_history_diagnostic.out"))
    gold = V.parse_vp_golden(os.path.join(golden_dir, "vp_2D_2D_k3.hyperrectangle_01.out"))
    assert len(rows) == len(gold) == 6
    for a, g in zip(rows, gold):
        assert abs(a[0] - g[0]) < 6e-4
        assert abs(a[1] - g[1]) <= 1e-7 * max(g[1], 1e-30) or g[1] == 0.0 == a[1]
        assert abs(a[3] - g[3]) <= 1e-12 * g[3] and abs(a[4] - g[4]) <= 1e-10 * g[4] and abs(a[5] - g[5]) <= 1e-10 * g[5]


def test_global_memory_tile_kernel_matches_oracle():
    """kernel_tile_global.cu (hd_advection_set_kernel 5): degree 5 and 3, FP32/FP64, plain apply and fused LSRK step"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tile_global_check.py")], capture_output=True, text=True, timeout=180)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("TG OK") == 6 and "TG FAIL" not in r.stdout
