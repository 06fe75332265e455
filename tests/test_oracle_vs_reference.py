"""The C oracle against the live Python reference (build container only; skipped on the GPU box)."""
import numpy as np
import pytest

import oracle
from _reference import load_reference
from helpers import FIELDS, max_rel
from msgwam_b200 import scenarios

ref_available = load_reference() is not None
pytestmark = pytest.mark.skipif(not ref_available, reason="/root/reference not present")


@pytest.mark.parametrize("sheared,shuffled,seed", [(False, False, 1), (True, False, 2), (True, True, 3)])
def test_rk3_synthetic_column_bitwise(sheared, shuffled, seed):
    ref = load_reference()
    sc = scenarios.column_ensemble(3000, seed=seed, ngrid=201, sheared=sheared, shuffled=shuffled, amplitude=0.3)
    sc.install(ref)
    orc = oracle.Oracle(sc.oracle_cfg())
    vr = vo = sc.var()
    for step in range(3):
        vr = ref.RK3(sc.dt, vr)
        vo = orc.RK3(sc.dt, vo)
        for i, nm in enumerate(FIELDS):
            assert np.array_equal(np.asarray(vr[i], dtype=np.float64), vo[i]), (step, nm)


def test_default_column_restatement_is_the_drivers_ic(golden):
    d = golden("driver_history.npz")
    sc = scenarios.default_column()
    for a, k in zip(sc.state, ("dens", "lambda", "phi", "rr", "drr", "kk", "ll", "mm", "dmm")):
        assert np.array_equal(a, d[k][0]), k
    assert np.array_equal(sc.uu, d["uu"][0]) and np.array_equal(sc.rhobar, d["rhobar"])
    assert np.array_equal(sc.pressure_gradient, d["pressure_gradient"])


def test_openmp_mode_agrees_to_summation_order():
    sc = scenarios.column_ensemble(20000, seed=5, ngrid=201, sheared=True, amplitude=0.3)
    a = oracle.Oracle(sc.oracle_cfg(), nthreads=1).RK3(sc.dt, sc.var())
    b = oracle.Oracle(sc.oracle_cfg(), nthreads=4).RK3(sc.dt, sc.var())
    for i, nm in enumerate(FIELDS):
        assert max_rel(b[i], a[i], floor=1e-300) <= 1e-12, nm
