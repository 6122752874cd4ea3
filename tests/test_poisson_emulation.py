"""CPU check of the x-space field-solve kernels' source (hyperdeal_b200/csrc/poisson_x.cu) through the host-emulation harness
(tests/vp_emulation_harness.cpp, one sequential "thread" per cell): the matrix-free SIP-DG Laplacian, the mass matrix and the
gradient-at-quadrature-points bodies against the oracle's dense operators (oracle/oracle_vp.py: PoissonDG, gradient_at_q),
which are pinned by the reference's Vlasov-Poisson golden.  Plus: a CG iteration built on the emulated operator reproduces the
oracle's potential (the device loop of hd_poisson_solve has the same structure)."""
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
    so = str(tmp_path_factory.mktemp("xsemu") / "libvpemu.so")
    csrc = os.path.join(ROOT, "hyperdeal_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", csrc, os.path.join(ROOT, "tests", "vp_emulation_harness.cpp"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.hd_xs_emulate.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, dp, ctypes.c_double]
    return lib


def _call(emu, which, src, n_out, dim_x, nq, ncx, h, scale=1.0):
    dp = ctypes.POINTER(ctypes.c_double)
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.zeros(n_out)
    rc = emu.hd_xs_emulate(which, src.ctypes.data_as(dp), dst.ctypes.data_as(dp), dim_x, 3, nq, (ctypes.c_int * dim_x)(*ncx), (ctypes.c_double * dim_x)(*h), float(scale))
    assert rc == 0
    return dst


CASES = [(1, (5,), None), (1, (1,), None), (2, (4, 3), None), (2, (2, 2), 5), (3, (2, 3, 2), None)]


@pytest.mark.parametrize("dim_x,ncx,nq", CASES)
def test_operator_mass_and_gradient_bodies_match_the_oracle(emu, dim_x, ncx, nq):
    b = O.basis_1d(3, nq)
    h = [0.7 + 0.2 * d for d in range(dim_x)]
    ref = V.PoissonDG(b, ncx, h)
    N = ref.K.shape[0]
    u = np.random.default_rng(4).standard_normal(N)
    tol = 1e-12 * np.abs(ref.K).max() * np.abs(u).max()
    assert np.max(np.abs(_call(emu, 0, u, N, dim_x, b.nq, ncx, h) - ref.K @ u)) <= 50 * tol
    assert np.max(np.abs(_call(emu, 1, u, N, dim_x, b.nq, ncx, h, scale=-1.0) + ref.M @ u)) <= 1e-13 * np.abs(u).max()
    # gradient table [cell][q][d]
    ncell = int(np.prod(ncx))
    nd, nqx = b.n**dim_x, b.nq**dim_x
    g = _call(emu, 2, u, ncell * nqx * dim_x, dim_x, b.nq, ncx, h).reshape(ncell, nqx, dim_x)

    def kron(mats):
        r = np.array([[1.0]])
        for m in mats:
            r = np.kron(m, r)
        return r

    pc = u.reshape(ncell, nd)
    for d in range(dim_x):
        op = kron([(b.G / h[e]) if e == d else b.S for e in range(dim_x)])
        assert np.max(np.abs(g[:, :, d] - pc @ op.T)) <= 1e-12 * np.abs(u).max() / min(h)


def test_cg_on_the_emulated_operator_reproduces_the_oracle_potential(emu):
    """hd_poisson_solve's iteration (zero-mean right-hand side, plain CG from a zero start) on the 2-D x-mesh of the golden case"""
    b = O.basis_1d(3)
    ncx, h = (4, 4), [np.pi, np.pi]
    ref = V.PoissonDG(b, ncx, h)
    N = ref.K.shape[0]
    rho = np.random.default_rng(8).standard_normal(N)
    rho -= rho.mean()
    rhs = _call(emu, 1, rho, N, 2, 4, ncx, h, scale=-1.0)
    rhs -= rhs.mean()
    A = lambda v: _call(emu, 0, v, N, 2, 4, ncx, h)
    x = np.zeros(N)
    r = rhs - A(x)
    p = r.copy()
    rr, bb = r @ r, rhs @ rhs
    it = 0
    while rr > (1e-11) ** 2 * bb and it < 2 * N:
        Ap = A(p)
        alpha = rr / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        rr_new = r @ r
        p = r + (rr_new / rr) * p
        rr = rr_new
        it += 1
    assert it < 2 * N
    exact = ref.solve(rhs)
    assert np.max(np.abs((x - x.mean()) - (exact - exact.mean()))) <= 1e-8 * np.abs(exact).max()
