"""CPU check of kernel_tile_global.cu (tile rounds with the partial sums kept in dst; the candidate for 3D3V degree 5 in FP32,
BASELINE.json configs[2]) through the host-emulation harness, against the oracle's literal ECL kernel."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tgemu") / "libvpemu.so")
    csrc = os.path.join(ROOT, "hyperdeal_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", csrc, os.path.join(ROOT, "tests", "vp_emulation_harness.cpp"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.hd_tg_emulate.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ip, dp, dp, dp, ctypes.c_double]
    lib.hd_tg_sp_emulate.argtypes = lib.hd_tg_emulate.argtypes
    return lib


VEL = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)


@pytest.mark.parametrize("dx,dv,nc,k,skew", [(3, 3, (2, 1, 1, 1, 1, 2), 5, 0.5), (2, 2, (2, 3, 2, 2), 5, 0.0), (1, 1, (3, 2), 5, 0.5), (3, 3, (2, 2, 1, 2, 1, 2), 3, 0.5),
                                             (2, 2, (3, 2, 2, 3), 3, 0.0)])
@pytest.mark.parametrize("smem_partials", [False, True], ids=["partials_in_dst", "partials_in_shared_memory"])
def test_tile_global_body_matches_literal_oracle(emu, dx, dv, nc, k, skew, smem_partials):
    dim = dx + dv
    left, right = (-1.0,) * dim, (1.0,) * dim
    mesh = O.Mesh(dx, dv, nc, left, right, (True,) * dim)
    orc = O.Oracle(mesh, k, skew=skew, velocity=VEL[:dim], nthreads=4)
    f = np.ascontiguousarray(np.random.default_rng(3).standard_normal(orc.ndofs))
    ref = orc.apply(f)
    out = np.zeros_like(f)
    dp = ctypes.POINTER(ctypes.c_double)
    vel = np.array(VEL[:dim])
    fn = emu.hd_tg_sp_emulate if smem_partials else emu.hd_tg_emulate
    out[:] = np.nan
    rc = fn(f.ctypes.data_as(dp), out.ctypes.data_as(dp), dim, k, (ctypes.c_int * dim)(*nc), (ctypes.c_double * dim)(*left), (ctypes.c_double * dim)(*right),
                           vel.ctypes.data_as(dp), float(skew))
    assert rc == 0
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
