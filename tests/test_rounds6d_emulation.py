"""CPU check of the three-round 3D3V degree-3 kernel's task bodies (hyperdeal_b200/csrc/rounds6d_tasks.cuh): compiled for the
host, run thread by thread on an emulated swizzled shared memory (tests/rounds6d_emulation.cpp), compared with the oracle's
literal ECL kernel (advection_operation.h:221-566)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("r6emu") / "libr6emu.so")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "rounds6d_emulation.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.hd_r6_emulate.argtypes = [dp, dp, ip, dp, dp, dp, ctypes.c_double, ctypes.c_int]
    return lib


@pytest.mark.parametrize("nc,vel,skew,ghost_mask", [
    ((3, 2, 2, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5, 0),
    ((2, 2, 1, 2, 2, 3), (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), 0.0, 0),
    ((2, 2, 2, 1, 2, 2), (1.0, 0.0, -0.05, 0.0, 0.0, 0.5), 0.5, 0),
    ((2, 2, 2, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5, 0b111111),
    ((3, 2, 1, 2, 2, 2), (-1.0, -0.15, 0.05, 0.1, 0.15, -0.5), 1.0, 0b101101),
])
def test_round_tasks_match_literal_oracle(emu, nc, vel, skew, ghost_mask):
    left, right = (-1.0,) * 6, (1.0,) * 6
    mesh = O.Mesh(3, 3, nc, left, right, (True,) * 6)
    orc = O.Oracle(mesh, 3, skew=skew, velocity=vel, nthreads=4)
    f = np.ascontiguousarray(np.random.default_rng(11).standard_normal(orc.ndofs))
    ref = orc.apply(f)
    out = np.zeros_like(f)
    dp = ctypes.POINTER(ctypes.c_double)
    v = np.array(vel)
    rc = emu.hd_r6_emulate(f.ctypes.data_as(dp), out.ctypes.data_as(dp), (ctypes.c_int * 6)(*nc), (ctypes.c_double * 6)(*left), (ctypes.c_double * 6)(*right),
                           v.ctypes.data_as(dp), float(skew), int(ghost_mask))
    assert rc == 0
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
