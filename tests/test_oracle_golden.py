"""Pin the CPU oracle against the reference's own golden outputs.

Fixtures in tests/golden/ are verbatim copies of
/root/reference/examples/advection/tests/*.{json,out,configuration} (data, not code).
Tolerance: the reference compares with `numdiff -a 1e-5 -r 1e-8`
(cmake/macros/macro_hyper_deal_pickup_tests.cmake:57); the goldens print 11 digits,
so 2e-10 relative is the resolution of the files and is what is asserted here.
FCL goldens (UseECL=false: _03/_04 in 1D1V, even numbers in 2D2V) are byte-identical to
their ECL twins in the reference, and are compared against the ECL oracle.
"""
import os

import pytest

from oracle import oracle as O

CASES_2D = ["adv_2D_2D_k3.hyperrectangle_%02d" % i for i in range(1, 9)]
CASES_Q5 = ["adv_2D_2D_k3_q5.hyperrectangle_%02d" % i for i in range(1, 5)]
CASES_1D = ["adv_1D_1D_k3.hyperrectangle_01", "adv_1D_1D_k3.hyperrectangle_01_rk33", "adv_1D_1D_k3.hyperrectangle_01_rk47", "adv_1D_1D_k3.hyperrectangle_01_rk59", "adv_1D_1D_k3.hyperrectangle_02", "adv_1D_1D_k3.hyperrectangle_03", "adv_1D_1D_k3.hyperrectangle_04"]


def _n_points(golden_dir, name):
    conf = open(os.path.join(golden_dir, name.split(".")[0] + ".configuration")).read()
    return int(conf.split("N_POINTS=")[1].split()[0])


def _compare(golden_dir, name, max_lines=None):
    nq = _n_points(golden_dir, name)
    lines = O.run_advection_example(os.path.join(golden_dir, name + ".json"), n_points=nq, nthreads=4, max_lines=max_lines)
    gold = O.parse_golden(os.path.join(golden_dir, name + ".out"))
    if max_lines is None:
        assert len(lines) == len(gold)
    for (t1, n1, e1), (t2, n2, e2) in zip(lines, gold):
        assert abs(t1 - t2) <= 1e-3 * max(abs(t2), 1e-3)
        assert abs(n1 - n2) <= 2e-10 * abs(n2), (name, t1, n1, n2)
        if e2 > 1e-12:  # collocation goldens have a round-off-level error at t=0
            assert abs(e1 - e2) <= 2e-10 * abs(e2), (name, t1, e1, e2)
        else:
            assert e1 < 1e-12


@pytest.mark.parametrize("name", CASES_2D + CASES_Q5)
def test_golden_2d2v(golden_dir, name):
    _compare(golden_dir, name)


@pytest.mark.parametrize("name", CASES_1D)
def test_golden_1d1v(golden_dir, name):
    _compare(golden_dir, name)


def test_lsrk_stage_counts():
    # tests/time_discretization/time_integrators_01.output: 3 5 7 9
    assert [len(O.lsrk_coefficients(k)[0]) for k in ("rk33", "rk45", "rk47", "rk59")] == [3, 5, 7, 9]


def test_lsrk_scalar_ode():
    # tests/time_discretization/time_integrators_02.cc:96-124 with golden
    # time_integrators_02.mpirun=1.output: y' = y sin^2(t), y(0) = 1, rk45, dt = 0.1,
    # 100 steps -> every entry prints "118.127" (deallog, 6 significant digits)
    import math

    import numpy as np

    y = np.array([1.0])
    dt = 0.1
    for it in range(100):
        y = O.lsrk_step(lambda v, tt: v * math.sin(tt) ** 2, y, dt * it, dt, "rk45")
    assert "%.6g" % y[0] == "118.127"
