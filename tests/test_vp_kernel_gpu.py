"""Device pieces of the Vlasov-Poisson rows.

* General-velocity advection kernel (kernel_vp.cu, hd_advection_set_phase_space_velocity): parity with the literal oracle on the
  GPU (first run: profiles/r01n_vp_kernel_gpu.txt, all six cases at round-off).
* The x-space field solve (poisson_x.cu) and the full right-hand side / golden run on the device pieces were written after round
  1's GPU budget was spent: their source is verified on the CPU (tests/test_poisson_emulation.py), the device run is pending —
  hence the non-strict xfail.
Both run in child processes (tests/vp_kernel_check.py, tests/vp_step_check.py), which keeps a possible fault of new device code
away from the CUDA context of this test session."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_general_velocity_kernel_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_kernel_check.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPK OK") == 6 and "VPK FAIL" not in r.stdout


@pytest.mark.xfail(strict=False, reason="kernel_vp.cu / poisson_x.cu have not run on a GPU yet (written after the round-1 GPU budget was spent)")
def test_vlasov_poisson_right_hand_side_and_golden_run_on_device():
    """density integration -> field solve -> general-velocity operator, one right-hand side against the oracle and then the
    reference's 2D2V Landau-damping golden (examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "vp_step_check.py")], capture_output=True, text=True, timeout=180)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("VPS OK") == 6 and "VPS FAIL" not in r.stdout
