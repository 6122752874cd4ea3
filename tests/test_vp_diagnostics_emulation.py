"""CPU check of the Vlasov-Poisson diagnostics kernels' source (hyperdeal_b200/csrc/vp_diagnostics.cu) through the host-emulation
harness: mass, squared L2 norm, kinetic energy, momentum and the field energy against the oracle's restatement of
examples/vlasov_poisson/include/diagnostics.h (oracle/oracle_vp.py, pinned by the reference's golden)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle_vp as V


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dgemu") / "libvpemu.so")
    csrc = os.path.join(ROOT, "hyperdeal_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", csrc, os.path.join(ROOT, "tests", "vp_emulation_harness.cpp"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.hd_diag_emulate.argtypes = [dp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, dp, dp]
    return lib


@pytest.mark.parametrize("dx,dv,nc,nq", [(1, 1, (3, 4), None), (2, 2, (2, 3, 2, 2), None), (2, 2, (2, 2, 3, 2), 5), (3, 3, (2, 1, 2, 2, 2, 1), None)])
def test_diagnostics_bodies_match_the_oracle(emu, dx, dv, nc, nq):
    dim = dx + dv
    left, right = (0.0,) * dx + (-1.3,) * dv, (2.0,) * dx + (1.7,) * dv
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, n_points=nq, nthreads=1)
    rng = np.random.default_rng(23)
    f = np.ascontiguousarray(rng.standard_normal(vp.adv.ndofs))
    a_v = np.ascontiguousarray(rng.standard_normal(vp.adv.a_v_table.shape))
    out6, en3 = np.zeros(6), np.zeros(3)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = emu.hd_diag_emulate(f.ctypes.data_as(dp), a_v.ctypes.data_as(dp), out6.ctypes.data_as(dp), en3.ctypes.data_as(dp), dx, dv, 3, nq or 4, (ctypes.c_int * dim)(*nc),
                             (ctypes.c_double * dim)(*left), (ctypes.c_double * dim)(*right))
    assert rc == 0
    ref = vp.phase_space_diagnostics(f)  # [mass, sqrt(sum f^2), kinetic, momentum...]
    scale = np.abs(f).max() * np.prod([r - l for l, r in zip(left, right)])
    assert abs(out6[0] - ref[0]) <= 1e-12 * scale
    assert abs(np.sqrt(out6[1]) - ref[1]) <= 1e-12 * ref[1]
    assert abs(out6[2] - ref[2]) <= 1e-11 * scale
    for d in range(dv):
        assert abs(out6[3 + d] - ref[3 + d]) <= 1e-11 * scale
    vp.adv.a_v_table[...] = a_v
    g = vp.adv.a_v_table
    jxw = np.array([1.0])
    for d in range(dx):
        jxw = np.kron(vp.b.w * vp.h[d], jxw)
    for d in range(dx):
        assert abs(en3[d] - np.sum(g[:, :, d] ** 2 * jxw[None, :])) <= 1e-12 * np.sum(g[:, :, d] ** 2 * jxw[None, :])
