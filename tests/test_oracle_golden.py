"""The C oracle against the fixtures generated from the unmodified Python reference.

Bit-for-bit wherever the reference's arithmetic is +,-,*,/,sqrt and libm sin/cos (which numpy
calls for float64 on this image); <= 4 ulp where numpy's own tan() is involved (HPROP-on dk/dl),
and a relative 1e-13 for var=3,4 projections (numpy pairwise sum vs sequential sum).
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from helpers import FIELDS, field_rel, max_rel, scenario_from_npz

CASES = ["random_col.npz", "random_col_sat.npz", "random_col_phi.npz", "random_hprop.npz", "random_hprop_sat.npz"]


def test_interp_matches_numpy_fixture():
    d = load_golden("interp.npz")
    orc = oracle.Oracle(dict(bvf=0.01, phi0=0.0, grid=np.linspace(0, 1, 3), grids=np.array([.25, .75]), dkk=1, dll=1, rr_mm_area=0))
    assert np.array_equal(orc.interp(d["x"], d["xp"], d["fp"]), d["y"])
    # and against numpy itself on fresh points (numpy is the third-party arithmetic of the reference)
    rng = np.random.default_rng(7)
    x = rng.uniform(-5e3, 45e3, 20000)
    assert np.array_equal(orc.interp(x, d["xp"], d["fp"]), np.interp(x, d["xp"], d["fp"]))


@pytest.mark.parametrize("case", CASES)
def test_rhs_default_matches_reference(case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    out = oracle.Oracle(sc.oracle_cfg()).rhs_default(sc.dt, sc.var())
    tol = 0.0 if not sc.hprop else 1e-14      # numpy tan() is not libm's
    for i, nm in enumerate(FIELDS):
        err = max_rel(out[i], d["rhs_" + nm])
        assert err <= tol, (case, nm, err)


@pytest.mark.parametrize("case", CASES)
def test_rk3_three_steps_match_reference(case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    orc = oracle.Oracle(sc.oracle_cfg())
    var = sc.var()
    tol = 0.0 if not sc.hprop else 1e-13
    for step in (1, 2, 3):
        var = orc.RK3(sc.dt, var)
        for i, nm in enumerate(FIELDS):
            err = max_rel(var[i], d["step%d_%s" % (step, nm)])
            assert err <= tol, (case, step, nm, err)


def test_rk3_default_column_bit_exact_through_720_steps():
    d = load_golden("rk3_default_column.npz")
    sc = scenario_from_npz(d)
    orc = oracle.Oracle(sc.oracle_cfg())
    var = sc.var()
    for step in range(1, 721):
        var = orc.RK3(sc.dt, var)
        if step in (1, 2, 10, 100, 360, 720):
            for i, nm in enumerate(FIELDS):
                assert np.array_equal(var[i], d["step%d_%s" % (step, nm)]), (step, nm)


@pytest.mark.parametrize("case", CASES)
def test_point_functions_and_saturation(case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    orc = oracle.Oracle(sc.oracle_cfg())
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    assert np.array_equal(orc.omega(kk, ll, mm, phi), d["omega"])
    assert np.array_equal(orc.omega(kk, ll, mm, sc.model["phi0"]), d["omega_phi0"])
    assert np.array_equal(orc.cg_rr(kk, ll, mm, lam, phi, rr), d["cg_rr"])
    for direct, key in ((False, "sat_tend"), (True, "sat_direct")):
        got = orc.saturation(sc.dt, d["sat_dens"], rr, d["sat_rr_st"], drr, d["sat_drr_st"], kk, ll, mm, d["sat_mm_st"], direct=direct)
        assert np.array_equal(got, d[key]), key
        assert (d["sat_direct"] != d["sat_dens"]).any()           # the fixture exercises the clamp


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("which", ["grids", "grid"])
def test_wave_projection_all_variants(case, which):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    orc = oracle.Oracle(sc.oracle_cfg())
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    g = d[which]
    for var in range(5):
        got = orc.wave_projection(dens, lam, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                                  sc.dkk, sc.dll, dmm, g, var=var)
        want = d["proj%d_%s" % (var, which)]
        assert got.shape == want.shape
        if var <= 2:
            assert np.array_equal(got, want), (var, which)
        else:
            assert field_rel(got, want) <= 1e-13, (var, which)


def test_projection_corner_cases():
    """half-cell offset, abs() weight, top cell never written, clamping."""
    d = load_golden("projection_corner_cases.npz")
    orc = oracle.Oracle(dict(bvf=float(d["bvf"]), phi0=float(d["phi0"]), grid=d["grid"], grids=d["grids"], dkk=1, dll=1, rr_mm_area=0))
    n = len(d["rr_low"])
    z = np.zeros(1)
    for i in range(n):
        s = slice(i, i + 1)
        for which in ("grids", "grid"):
            for var in (0, 1, 2):
                got = orc.wave_projection(d["dens"][s], z, z, d["rr_low"][s], d["rr_up"][s], d["kk"][s], d["ll"][s],
                                          d["mm"][s] - .5 * d["dmm"][s], d["mm"][s] + .5 * d["dmm"][s],
                                          d["dkk"][s], d["dll"][s], d["dmm"][s], d[which], var=var)
                assert np.array_equal(got, d["ray%d_proj%d_%s" % (i, var, which)]), (i, var, which)
    # the documented quirks, as weights of var=2 (dens * psv == 1e9 * 1e-15)
    w = lambda i: d["ray%d_proj2_grids" % i] / (1e9 * 1e-4 * 1e-4 * 1e-7)
    assert np.isclose(w(0)[1], 0.2) and np.count_nonzero(w(0)) == 1          # [1200,1700]: only 0.2 in cell 1
    assert np.isclose(w(1)[0], 0.15)                                         # [100,350]: spurious |350-500|/1000
    assert np.count_nonzero(w(2)) == 0                                       # [98600,99400]: top cell never written
    assert np.allclose(w(3)[[96, 97]], 1.0) and np.count_nonzero(w(3)) == 2  # clamped straddler


def test_driver_fixture_matches_the_known_answers_recorded_in_the_survey():
    """SURVEY.md section 4 lists values read off the unmodified raytracer.py run (numpy 2.3.5); the committed fixture
    must reproduce them, so the fixture itself is pinned to an independent record."""
    d = load_golden("driver_history.npz")
    steps = list(d["steps"])
    k1, kend = steps.index(1), steps.index(1440)
    assert np.array_equal(d["rr"][k1][:3], [219.07715034112215, 469.07715034112215, 719.0771503411221])
    assert np.allclose(d["mm"][k1][:3], -0.00125665223824475, rtol=1e-14, atol=0)       # the survey printed 15 digits
    assert np.array_equal(d["uu"][k1][38:42], [-1.3771329439311946, -0.587158015003212, 0.6489099624965707, 1.858935033568595])
    assert np.allclose(d["rr"][kend][:3], [92862.22980969641, 92862.41977331725, 92905.23567557707], rtol=1e-13, atol=0)
    assert np.allclose(d["mm"][kend][:3], [-0.00304211131909678, -0.00302430454016963, -0.00299897575710608], rtol=1e-14, atol=0)
    assert np.allclose(d["dens"][kend][:3], [8.0293785331725159e9, 1.2631985534866663e10, 1.9293364183825668e10], rtol=1e-13, atol=0)
    assert abs(float(d["wa_max"]) - 3.8994466052014998) <= 1e-12 and abs(float(d["flux_diag_absmax"]) - 1.8943901432301569) <= 1e-12


def test_frozen_background_oracle_vs_reference_fixture():
    """oracle.RK3_frozen against the fixture that tests/golden/make_golden_frozen.py composed from the unmodified
    reference's functions (extension "M2": frozen mean flow over the step, one deposit per step)."""
    d = load_golden("frozen_col.npz")
    sc = scenario_from_npz(d)
    orc = oracle.Oracle(sc.oracle_cfg())
    var = sc.var()
    for step in (1, 2, 3):
        var = orc.RK3_frozen(sc.dt, var)
        for i, nm in enumerate(FIELDS):
            assert np.array_equal(np.asarray(var[i], dtype=np.float64), d["step%d_%s" % (step, nm)]), (step, nm)
