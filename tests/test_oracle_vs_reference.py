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


def test_frozen_background_oracle_is_the_composition_of_reference_functions():
    """Extension "M2" (frozen mean flow over a step): the oracle's RK3_frozen against the same composition of the live
    Python reference's own functions -- its RK3 through model_config['rhs'] (L:691) with du_st = dv_st = 0, then
    wave_projection + du_dt / dv_dt once per step -- bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_frozen import frozen_step, install_frozen
    ref = load_reference()
    sc = scenarios.column_ensemble(2500, seed=12, ngrid=151, sheared=True, amplitude=0.3)
    install_frozen(ref, sc)
    orc = oracle.Oracle(sc.oracle_cfg())
    vr = vo = sc.var()
    for step in range(3):
        vr = frozen_step(ref, sc.dt, vr)
        vo = orc.RK3_frozen(sc.dt, vo)
        for i, nm in enumerate(FIELDS):
            assert np.array_equal(np.asarray(vr[i], dtype=np.float64), np.asarray(vo[i], dtype=np.float64)), (step, nm)
    assert np.max(np.abs(vo[9] - sc.uu)) > 1e-3          # the deposit does move the wind
