"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the driver reads, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDoF/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "advection operator throughput (3D3V, k=3, FP64)" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "3D3V k=3 FP64 lattice" in d["cpu_baseline"]["sample"]
    assert isinstance(d["cpu_baseline"]["same_config"], bool)
    assert d["e2e"] == {"value": d["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank_zero_works():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_product_arm_needs_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
@pytest.mark.parametrize("workload", ["k5f32", "vp2d2v", "lsrk"])
def test_secondary_workloads_need_a_gpu(workload):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)


def test_tools_and_entry_points_compile():
    """every script under tools/, bench.py and __graft_entry__.py at least byte-compiles (they only run on a GPU box)"""
    import glob
    import py_compile

    for path in sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"),
                                                                          os.path.join(ROOT, "tests", "mgpu_check.py"), os.path.join(ROOT, "tests", "vp_kernel_check.py"),
                                                                          os.path.join(ROOT, "tests", "vp_step_check.py"), os.path.join(ROOT, "tests", "tile_global_check.py")]:
        py_compile.compile(path, doraise=True)
