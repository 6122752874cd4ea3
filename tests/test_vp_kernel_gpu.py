"""General-velocity advection kernel for the Vlasov-Poisson velocity field (kernel_vp.cu, hd_advection_set_phase_space_velocity;
SURVEY.md §8f 1).  The kernel was written after round 1's GPU budget was spent: its algebra is verified on the CPU
(tests/test_collapsed_general_velocity.py), the device code is not validated yet — hence the non-strict xfail and the separate
process (tests/vp_kernel_check.py), which keeps a possible fault away from the CUDA context of this test session."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="kernel_vp.cu has not run on a GPU yet (written after the round-1 GPU budget was spent)")
def test_general_velocity_kernel_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_kernel_check.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPK OK") == 6 and "VPK FAIL" not in r.stdout


@pytest.mark.xfail(strict=False, reason="kernel_vp.cu / poisson_x.cu have not run on a GPU yet (written after the round-1 GPU budget was spent)")
def test_vlasov_poisson_right_hand_side_and_golden_run_on_device():
    """density integration -> field solve -> general-velocity operator, one right-hand side against the oracle and then the
    reference's 2D2V Landau-damping golden (examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_step_check.py")], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPS OK") == 6 and "VPS FAIL" not in r.stdout
