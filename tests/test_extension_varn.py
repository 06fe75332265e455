"""The N(z) extension of the oracle (no counterpart in the reference; parity unpinned by the reference):
(1) a constant profile must reduce bit-for-bit to the reference's scalar-bvf results (golden fixtures),
(2) a varying profile must agree bit-for-bit with the independent numpy restatement of the extension."""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from ext_reference import ExtReference
from helpers import FIELDS, max_rel, scenario_from_npz

CASES = ["random_col.npz", "random_col_sat.npz", "random_hprop_sat.npz"]


def n_profile(grids):
    n2 = 1e-4 * (1 + 3 * .5 * (1 + np.tanh((grids - 15e3) / 3e3)))     # SURVEY.md 8(d), configs[2]
    return np.sqrt(n2)


@pytest.mark.parametrize("case", CASES)
def test_constant_profile_reduces_to_reference(case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    cfg = sc.oracle_cfg()
    cfg["bvf"] = np.full(len(sc.grids), sc.model["bvf"])
    orc = oracle.Oracle(cfg)
    out = orc.rhs_default(sc.dt, sc.var())
    tol = 0.0 if not sc.hprop else 1e-14
    for i, nm in enumerate(FIELDS):
        assert max_rel(out[i], d["rhs_" + nm]) <= tol, (case, nm)
    var = sc.var()
    for step in (1, 2, 3):
        var = orc.RK3(sc.dt, var)
        for i, nm in enumerate(FIELDS):
            assert max_rel(var[i], d["step%d_%s" % (step, nm)]) <= (0.0 if not sc.hprop else 1e-13), (case, step, nm)


@pytest.mark.parametrize("case", CASES)
def test_varying_profile_matches_numpy_restatement(case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    cfg = sc.oracle_cfg()
    cfg["bvf"] = n_profile(sc.grids) * (sc.model["bvf"] / 0.01)
    ext = ExtReference(cfg)
    orc = oracle.Oracle(cfg)
    a, b = orc.rhs_default(sc.dt, sc.var()), ext.rhs(sc.dt, sc.var())
    tol = 0.0 if not sc.hprop else 1e-14          # numpy's tan is not libm's
    for i, nm in enumerate(FIELDS):
        assert max_rel(a[i], b[i]) <= tol, (case, nm, max_rel(a[i], b[i]))
    assert np.abs(a[4]).max() > 0 and np.abs(a[8]).max() > 0          # extents now evolve: cgr_up != cgr_down
    va = vb = sc.var()
    for step in range(2):
        va, vb = orc.RK3(sc.dt, va), ext.RK3(sc.dt, vb)
    for i, nm in enumerate(FIELDS):
        assert max_rel(va[i], vb[i]) <= (0.0 if not sc.hprop else 1e-13), (case, nm)


# ---- CUDA path (general, stage-by-stage kernels) vs the oracle's extension branch -------------------------------
@pytest.fixture()
def lprop():
    import importlib
    import msgwam_b200.libprop as lp
    importlib.reload(lp)
    return lp


def _profile_scenario(case, constant=False):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    prof = np.full(len(sc.grids), sc.model["bvf"]) if constant else n_profile(sc.grids) * (sc.model["bvf"] / 0.01)
    sc.model = dict(sc.model, bvf=prof)
    return d, sc


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_constant_profile_reduces_to_reference_fixture(lprop, case):
    from test_gpu_parity import assert_state_close
    d, sc = _profile_scenario(case, constant=True)
    sc.install(lprop)
    out = lprop.rhs_default(sc.dt, sc.var())
    assert_state_close(out, [d["rhs_" + nm] for nm in FIELDS], tag=case)
    var = sc.var()
    for step in (1, 2, 3):
        var = lprop.RK3(sc.dt, var)
        assert_state_close(var, [d["step%d_%s" % (step, nm)] for nm in FIELDS], tag="%s step %d" % (case, step))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_varying_profile_vs_oracle(lprop, case):
    from test_gpu_parity import assert_state_close
    d, sc = _profile_scenario(case)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    # ddrr_st = cgr_up - cgr_down cancels ~3 digits under N(z); with phi != 0 CUDA's sin (<= 2 ulp from libm's)
    # enters both terms, so the difference is compared at 1e-10 there (bit-faithful when phi == 0)
    assert_state_close(lprop.rhs_default(sc.dt, sc.var()), orc.rhs_default(sc.dt, sc.var()), tag=case,
                       ray_tol=1e-10 if sc.hprop else 1e-13)
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    trig = 4e-16 if sc.model["phi0"] == 0.0 and not sc.hprop else 1e-13
    assert max_rel(lprop.omega(kk, ll, mm, phi, rr), orc.omega(kk, ll, mm, phi, rr)) <= trig
    assert max_rel(lprop.cg_rr(kk, ll, mm, lam, phi, rr), orc.cg_rr(kk, ll, mm, lam, phi, rr)) <= trig
    with pytest.raises(TypeError):
        lprop.omega(kk, ll, mm, phi)                      # a profile needs the height
    for v in (0, 1, 2, 3, 4):
        args = (dens, lam, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                sc.dkk, sc.dll, dmm, sc.grids)
        got, want = lprop.wave_projection(*args, var=v), orc.wave_projection(*args, var=v)
        assert np.max(np.abs(got - want)) <= 1e-12 * max(np.max(np.abs(want)), 1e-300), (case, v)
    va = vb = sc.var()
    start = sc.var()
    for step in range(3):
        va, vb = lprop.RK3(sc.dt, va), orc.RK3(sc.dt, vb)
        assert_state_close(va, vb, tag="%s step %d" % (case, step + 1), start=start)
    assert np.max(np.abs(np.asarray(va[4]) - sc.state[4])) > 0         # extents evolve under N(z)
    st = lprop.saturation(sc.dt, dens, rr, drr * 0 + .1, drr, drr * 0, kk, ll, mm, mm * 1e-6, direct=True)
    sw = orc.saturation(sc.dt, dens, rr, drr * 0 + .1, drr, drr * 0, kk, ll, mm, mm * 1e-6, direct=True)
    assert max_rel(st, sw) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("n,ngrid,shuffled", [(60011, 501, False), (60011, 1001, True), (257, 34, False)])
def test_gpu_fused_column_step_with_profile_vs_oracle(lprop, n, ngrid, shuffled):
    """The fused two-sweep column step with an N(z) profile (msgwam_column_step_nz: rr, drr, mm, dmm evolve) on a
    sheared, feeding-back ensemble, three steps, through lprop.RK3 and in place through RayEnsemble."""
    from msgwam_b200 import scenarios
    from msgwam_b200.ensemble import RayEnsemble
    from test_gpu_parity import assert_state_close
    sc = scenarios.column_ensemble(n, seed=77, ngrid=ngrid, sheared=True, shuffled=shuffled, amplitude=0.3)
    sc.model = dict(sc.model, bvf=n_profile(sc.grids))
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    got = want = sc.var()
    for step in range(3):
        got, want = lprop.RK3(sc.dt, got), orc.RK3(sc.dt, want)
        assert_state_close(got, want, ray_tol=1e-12, grid_tol=1e-11, tag="step %d" % (step + 1), start=sc.var())
    assert np.max(np.abs(np.asarray(got[4]) - sc.state[4])) > 0 and np.max(np.abs(np.asarray(got[8]) - sc.state[8])) > 0
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt, 3)
    assert_state_close(ens.to_var(), want, ray_tol=1e-12, grid_tol=1e-11, tag="ensemble", start=sc.var())
