"""CPU check of the general-velocity kernel's source (hyperdeal_b200/csrc/kernel_vp.cu): the per-cell body is written against
(tid, nthreads) and a barrier macro, so the very same code compiles with g++ as one sequential "thread" per cell
(tests/vp_emulation_harness.cpp includes it with -DHD_VP_HOST_EMULATION).  A barrier-synchronised kernel computes the same thing in that mode, so its index logic, sweeps and
coefficients (basis.hpp, the product's own) can be compared with the oracle's literal kernel without a GPU.  This harness
exists only here; the product library has no CPU path.  What it cannot show: races, launch configuration, shared-memory
limits — those need the GPU run of tests/test_zz_vp_device_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from oracle import oracle_vp as V


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("vpemu") / "libvpemu.so")
    csrc = os.path.join(ROOT, "hyperdeal_b200", "csrc")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", csrc, os.path.join(ROOT, "tests", "vp_emulation_harness.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.hd_vp_emulate.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, dp, dp, ctypes.c_double]
    lib.hd_vp_tile_emulate.argtypes = [dp, dp, dp, ctypes.c_int, ip, dp, dp, ctypes.c_double]
    return lib


TILE_CASES = [
    # dx cells          skew
    (1, (3, 4), 0.0),
    (1, (2, 3), 0.5),
    (1, (1, 1), 0.0),
    (1, (5, 2), 0.3),
    (2, (2, 3, 2, 2), 0.0),
    (2, (3, 2, 2, 3), 0.5),
    (2, (1, 1, 1, 1), 0.0),       # one cell: every neighbour is the cell itself; the second half of the warp idles
    (2, (3, 1, 1, 3), 0.3),       # odd number of cells
    (2, (4, 2, 3, 2), 1.0),
    (3, (2, 1, 2, 2, 2, 1), 0.0),
    (3, (2, 3, 2, 1, 2, 2), 0.5),
    (3, (1, 1, 1, 1, 1, 1), 0.3),
]


@pytest.mark.parametrize("dx,nc,skew", TILE_CASES)
def test_tile_kernel_source_matches_literal_oracle(emu, dx, nc, skew):
    """kernel_vp_tile.cuh (degree 3, n_points 4; 1D1V thread-per-cell, 2D2V warp phases, 3D3V CTA phases) against the oracle's literal kernel"""
    dim = 2 * dx
    left, right = (0.0,) * dx + (-1.3,) * dx, (2.0,) * dx + (1.7,) * dx  # v = 0 lies inside a cell
    vp = V.VlasovPoissonOracle(dx, dx, 3, nc, left, right, nthreads=2)
    rng = np.random.default_rng(23)
    a_v = np.ascontiguousarray(rng.standard_normal(vp.adv.a_v_table.shape))
    orc = O.Oracle(vp.mesh, 3, skew=skew, a_x_table=vp.v_at_q, a_v_table=a_v, nthreads=2)
    f = np.ascontiguousarray(rng.standard_normal(orc.ndofs))
    ref = orc.apply(f)
    out = np.full_like(f, np.nan)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    rc = emu.hd_vp_tile_emulate(f.ctypes.data_as(dp), out.ctypes.data_as(dp), a_v.ctypes.data_as(dp), dx, (ctypes.c_int * dim)(*nc),
                                (ctypes.c_double * dim)(*left), (ctypes.c_double * dim)(*right), float(skew))
    assert rc == 0
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))


CASES = [
    # dx dv cells                   nq    skew
    (1, 1, (3, 4), None, 0.0),
    (1, 1, (2, 3), None, 0.5),
    (1, 1, (1, 1), None, 0.0),                # single cell: both neighbours are the cell itself
    (2, 2, (2, 3, 2, 2), None, 0.0),
    (2, 2, (2, 2, 3, 2), 5, 0.3),             # over-integration: the sweeps change extents (n -> nq -> n)
    (3, 3, (2, 1, 2, 2, 2, 1), None, 0.0),
]


@pytest.mark.parametrize("dx,dv,nc,nq,skew", CASES)
def test_kernel_source_matches_literal_oracle(emu, dx, dv, nc, nq, skew):
    dim = dx + dv
    left, right = (0.0,) * dx + (-1.3,) * dv, (2.0,) * dx + (1.7,) * dv  # v = 0 lies inside a cell
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, n_points=nq, nthreads=2)
    rng = np.random.default_rng(17)
    a_v = np.ascontiguousarray(rng.standard_normal(vp.adv.a_v_table.shape))
    orc = O.Oracle(vp.mesh, 3, nq=nq, skew=skew, a_x_table=vp.v_at_q, a_v_table=a_v, nthreads=2)
    f = np.ascontiguousarray(rng.standard_normal(orc.ndofs))
    ref = orc.apply(f)
    out = np.zeros_like(f)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    ncell = (ctypes.c_int * dim)(*nc)
    lo, hi = (ctypes.c_double * dim)(*left), (ctypes.c_double * dim)(*right)
    rc = emu.hd_vp_emulate(f.ctypes.data_as(dp), out.ctypes.data_as(dp), a_v.ctypes.data_as(dp), dx, dv, 3, nq or 4, ncell, lo, hi, float(skew))
    assert rc == 0
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_kernel_source_reduces_to_constant_velocity_when_the_table_is_constant(emu):
    """a_v constant and a v-mesh far from v = 0: compare with the constant-velocity literal kernel in the v-directions"""
    dx = dv = 1
    nc, left, right = (3, 2), (0.0, 2.0), (1.0, 3.0)
    mesh = O.Mesh(dx, dv, nc, left, right, (True, True))
    vp = V.VlasovPoissonOracle(dx, dv, 3, nc, left, right, nthreads=1)
    a_v = np.full(vp.adv.a_v_table.shape, -0.7)
    orc = O.Oracle(mesh, 3, a_x_table=vp.v_at_q, a_v_table=a_v, nthreads=1)
    f = np.ascontiguousarray(np.random.default_rng(5).standard_normal(orc.ndofs))
    out = np.zeros_like(f)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = emu.hd_vp_emulate(f.ctypes.data_as(dp), out.ctypes.data_as(dp), np.ascontiguousarray(a_v).ctypes.data_as(dp), dx, dv, 3, 4, (ctypes.c_int * 2)(*nc),
                           (ctypes.c_double * 2)(*left), (ctypes.c_double * 2)(*right), 0.0)
    assert rc == 0
    assert np.max(np.abs(out - orc.apply(f))) <= 1e-12 * np.max(np.abs(f))
