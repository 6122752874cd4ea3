"""The vectorised CPU baseline of bench.py (oracle/hd_ecl_simd.cpp: the reference's literal ECL algorithm, 8 cells per SIMD
batch) against the scalar oracle that the reference's golden files pin (oracle/hd_oracle.cpp)."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("nc,vel,skew", [
    ((8, 2, 2, 1, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5),
    ((3, 2, 1, 2, 2, 2), (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), 0.0),   # partial SIMD batch, all signs flipped
    ((11, 1, 2, 2, 1, 2), (1.0, 0.0, -0.05, 0.0, 0.0, 0.5), 1.0),        # one full + one partial batch per row, zero components
])
def test_simd_baseline_matches_scalar_oracle(nc, vel, skew):
    left, right = (-1.0, 0.0, -1.0, 0.0, -2.0, 0.0), (1.0, 1.0, 1.0, 2.0, 1.0, 3.0)
    orc = O.Oracle(O.Mesh(3, 3, nc, left, right, (True,) * 6), 3, skew=skew, velocity=vel, nthreads=4)
    f = np.ascontiguousarray(np.random.default_rng(5).standard_normal(orc.ndofs))
    ref = orc.apply(f)
    for threads in (1, 3):
        out = O.FastECL(nc, left, right, vel, skew=skew, nthreads=threads, pin=False).apply(f)
        assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
