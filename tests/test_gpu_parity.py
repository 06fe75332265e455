"""CUDA path vs the reference's golden fixtures and vs the CPU oracle (B200 box; -m gpu).

Every call goes through msgwam_b200.libprop -> ctypes -> libmsgwam_b200.so (the C ABI).
Tolerances
  * per-ray state: the contract is <= 1e-10 relative after N steps (BASELINE.json north_star); the
    arithmetic is the reference's operation for operation, so short runs are asserted at 1e-13.
    Ray arithmetic itself is bit-faithful; differences come only from the mean flow, whose deposit is
    summed in a different order (fp64 atomics), and from CUDA's sin/cos/tan (<= 2 ulp) when phi != 0.
  * deposited grid fields / uu, vv: max|a-b| <= 1e-12 * max|b|  (summation order of ~n/G terms per cell).
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from helpers import FIELDS, field_rel, max_rel, scenario_from_npz
from msgwam_b200 import scenarios

pytestmark = pytest.mark.gpu

CASES = ["random_col.npz", "random_col_sat.npz", "random_col_phi.npz", "random_hprop.npz", "random_hprop_sat.npz"]
RAY_TOL = 1e-13
GRID_TOL = 1e-12


@pytest.fixture()
def lprop():
    import importlib
    import msgwam_b200.libprop as lp
    importlib.reload(lp)            # fresh module globals per test, like a fresh `import lib.libprop`
    return lp


def assert_state_close(got, want, ray_tol=RAY_TOL, grid_tol=GRID_TOL, tag="", start=None):
    """Per-ray fields: |got - want| <= ray_tol * max(|want|, |want - start|): relative to the value, or -- where
    the step nearly cancels the value (a wavenumber driven through zero) -- to the size of the update."""
    for i, nm in enumerate(FIELDS):
        g, w = np.asarray(got[i], dtype=np.float64), np.asarray(want[i], dtype=np.float64)
        assert g.shape == w.shape, (tag, nm, g.shape, w.shape)
        if nm in ("uu", "vv"):
            err = field_rel(g, w)
            assert err <= grid_tol, (tag, nm, err)
        else:
            scale = np.abs(w) if start is None else np.maximum(np.abs(w), np.abs(w - np.asarray(start[i], dtype=np.float64)))
            diff = np.abs(g - w)
            err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale)))) if g.size else 0.0
            assert err <= ray_tol, (tag, nm, err)


@pytest.mark.parametrize("case", CASES)
def test_rhs_default_vs_reference_fixture(lprop, case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    sc.install(lprop)
    out = lprop.rhs_default(sc.dt, sc.var())
    assert_state_close(out, [d["rhs_" + nm] for nm in FIELDS], tag=case)


@pytest.mark.parametrize("case", CASES)
def test_rk3_three_steps_vs_reference_fixture(lprop, case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    sc.install(lprop)
    var = sc.var()
    for step in (1, 2, 3):
        var = lprop.RK3(sc.dt, var)
        assert_state_close(var, [d["step%d_%s" % (step, nm)] for nm in FIELDS], tag="%s step %d" % (case, step))


def test_rk3_default_column_vs_reference_fixture(lprop):
    """BASELINE configs[0]: the driver's wave packet, pure RK3, 360 steps (<= 1e-10 per the contract)."""
    d = load_golden("rk3_default_column.npz")
    sc = scenario_from_npz(d)
    sc.install(lprop)
    var = sc.var()
    worst = 0.0
    for step in range(1, 361):
        var = lprop.RK3(sc.dt, var)
        if step in (1, 2, 10, 100, 360):
            tol = RAY_TOL if step <= 10 else 1e-10
            assert_state_close(var, [d["step%d_%s" % (step, nm)] for nm in FIELDS], ray_tol=tol, grid_tol=1e-10,
                               tag="step %d" % step)
            worst = max(worst, max(max_rel(var[i], d["step%d_%s" % (step, nm)]) for i, nm in enumerate(FIELDS[:9])))
    print("default column, worst per-ray relative error through 360 steps: %.3e" % worst)


def test_driver_loop_with_offline_saturation_vs_reference_history(lprop):
    """The driver's own loop (R:157-188): RK3 then saturation(direct=True), against raytracer.py's history."""
    d = load_golden("driver_history.npz")
    sc = scenarios.default_column()
    sc.install(lprop)
    dt = sc.dt
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = [a.copy() for a in sc.state]
    uu, vv = sc.uu.copy(), sc.vv.copy()
    steps = list(d["steps"])
    for nt in range(1, 101):
        state_in = np.array([dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv], dtype=object)
        out = lprop.RK3(dt, state_in)
        dens_prop, lam_n, phi_n, rr_n, drr_n, kk_n, ll_n, mm_n, dmm_n, uu_n, vv_n = out
        dens_n = lprop.saturation(dt, dens_prop, rr, (rr_n - rr) / 1, drr, (drr_n - drr) / dt, kk_n, ll_n, mm,
                                  (mm_n - mm) / dt, direct=True)
        dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv = dens_n, lam_n, phi_n, rr_n, drr_n, kk_n, ll_n, mm_n, dmm_n, uu_n, vv_n
        if nt in steps:
            k = steps.index(nt)
            for nm, val in (("dens", dens), ("rr", rr), ("mm", mm), ("drr", drr), ("dmm", dmm)):
                assert max_rel(val, d[nm][k]) <= 1e-11, (nt, nm, max_rel(val, d[nm][k]))
            assert field_rel(uu, d["uu"][k]) <= 1e-11, nt


@pytest.mark.parametrize("case", CASES)
def test_point_functions_and_saturation_vs_fixture(lprop, case):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    sc.install(lprop)
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    trig = 4e-16 if float(d["phi0"]) == 0.0 and not sc.hprop else 1e-13
    assert max_rel(lprop.omega(kk, ll, mm, phi), d["omega"]) <= trig
    assert np.array_equal(lprop.omega(kk, ll, mm, sc.model["phi0"]), d["omega_phi0"])
    assert max_rel(lprop.cg_rr(kk, ll, mm, lam, phi, rr), d["cg_rr"]) <= trig
    for nm in ("cg_lambda", "cg_phi", "dk_dt", "dl_dt", "dm_dt"):
        got = getattr(lprop, nm)(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        assert got.shape == d[nm].shape
        assert max_rel(got, d[nm], floor=1e-300) <= 1e-12, (nm, max_rel(got, d[nm]))
    gr = lprop.gradients(lam, phi, rr, sc.uu, sc.vv)
    assert gr.shape == d["gradients"].shape and np.array_equal(gr, d["gradients"])
    assert np.array_equal(lprop.du_dt(sc.vv, d["flux_grad"]), d["du_dt"])
    assert np.array_equal(lprop.dv_dt(sc.uu, d["flux_grad"]), d["dv_dt"])
    for direct, key in ((False, "sat_tend"), (True, "sat_direct")):
        got = lprop.saturation(sc.dt, d["sat_dens"], rr, d["sat_rr_st"], drr, d["sat_drr_st"], kk, ll, mm,
                               d["sat_mm_st"], direct=direct)
        assert np.array_equal(got, d[key]), key


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("which", ["grids", "grid"])
def test_wave_projection_vs_fixture(lprop, case, which):
    d = load_golden(case)
    sc = scenario_from_npz(d)
    sc.install(lprop)
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    for var in range(5):
        got = lprop.wave_projection(dens, lam, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                                    sc.dkk, sc.dll, dmm, d[which], var=var)
        want = d["proj%d_%s" % (var, which)]
        assert got.shape == want.shape
        for c in range(want.shape[0] if want.ndim == 2 else 1):
            g, w = (got[c], want[c]) if want.ndim == 2 else (got, want)
            assert field_rel(g, w) <= GRID_TOL, (case, which, var, field_rel(g, w))
            assert np.array_equal(g == 0, w == 0), (case, which, var, "support differs")


def test_projection_corner_cases_single_rays_bit_exact(lprop):
    """One ray at a time there is no summation order: the deposit must match bit for bit, including the
    half-cell offset, the abs() weight and the never-written top cell."""
    d = load_golden("projection_corner_cases.npz")
    lprop.HPROP_GLOBAL = False
    lprop.set_model_setup(bvf=float(d["bvf"]), phi0=float(d["phi0"]))
    z = np.zeros(1)
    for i in range(len(d["rr_low"])):
        s = slice(i, i + 1)
        for which in ("grids", "grid"):
            for var in (0, 1, 2):
                got = lprop.wave_projection(d["dens"][s], z, z, d["rr_low"][s], d["rr_up"][s], d["kk"][s], d["ll"][s],
                                            d["mm"][s] - .5 * d["dmm"][s], d["mm"][s] + .5 * d["dmm"][s],
                                            d["dkk"][s], d["dll"][s], d["dmm"][s], d[which], var=var)
                assert np.array_equal(got, d["ray%d_proj%d_%s" % (i, var, which)]), (i, var, which)


@pytest.mark.parametrize("sheared,shuffled,n,ngrid", [(False, False, 100003, 1001), (True, False, 100003, 1001),
                                                      (True, True, 50021, 401), (True, False, 777, 101)])
def test_rk3_column_ensemble_vs_oracle(lprop, sheared, shuffled, n, ngrid):
    """Synthetic column ensembles with wave amplitudes that feed back on the wind:
    numpy-in/numpy-out path and torch device-tensor path against the oracle, 3 steps."""
    import torch
    sc = scenarios.column_ensemble(n, seed=11, ngrid=ngrid, sheared=sheared, shuffled=shuffled, amplitude=0.3)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    vo = vn = sc.var()
    vt = np.empty(11, dtype=object)
    for i in range(11):
        vt[i] = torch.from_numpy(np.ascontiguousarray(vn[i])).cuda()
    lprop.set_statics(dkk=torch.from_numpy(sc.dkk).cuda(), dll=torch.from_numpy(sc.dll).cuda(), rr_mm_area=torch.from_numpy(sc.rr_mm_area).cuda())
    for step in range(3):
        vo = orc.RK3(sc.dt, vo)
        vt = lprop.RK3(sc.dt, vt)
        assert_state_close([t.cpu().numpy() for t in vt], vo, tag="torch step %d" % step, start=sc.var())
    lprop.set_statics(dkk=sc.dkk, dll=sc.dll, rr_mm_area=sc.rr_mm_area)
    vo = sc.var()
    for step in range(2):
        vo = orc.RK3(sc.dt, vo)
        vn = lprop.RK3(sc.dt, vn)
        assert_state_close(vn, vo, tag="numpy step %d" % step, start=sc.var())
    assert np.abs(vo[9]).max() > 0 and not np.array_equal(vo[9], sc.uu)      # the deposit did feed back


def test_rk3_analytic_constant_background(lprop):
    """Config-2 known answer: constant N, zero wind, f = 0 => c_g constant, m constant, z(t) = z0 + c_g t
    (RK3 is exact for a constant right-hand side)."""
    sc = scenarios.column_ensemble(4096, seed=3, ngrid=1001)
    sc.install(lprop)
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    cg = lprop.cg_rr(kk, ll, mm, lam, phi, rr)
    var = sc.var()
    for _ in range(5):
        var = lprop.RK3(sc.dt, var)
    assert max_rel(var[3], rr + cg * sc.dt * 5) <= 1e-13
    assert max_rel(var[7], mm) <= 1e-13


def test_user_supplied_rhs_plugin(lprop):
    """model_config['rhs'] is the reference's plug-in point (L:691): any callable must be honoured."""
    sc = scenarios.default_column()
    sc.install(lprop)
    calls = []

    def my_rhs(dt, var):
        calls.append(1)
        return lprop.rhs_default(dt, var)
    lprop.set_model_setup(rhs=my_rhs)
    a = lprop.RK3(sc.dt, sc.var())
    lprop.set_model_setup(rhs=lprop.rhs_default)
    b = lprop.RK3(sc.dt, sc.var())
    assert len(calls) == 3
    assert_state_close(a, b, tag="plugin vs fused")


def test_missing_statics_raise_keyerror(lprop):
    sc = scenarios.default_column()
    sc.install(lprop)
    lprop.statics.pop("dkk")
    with pytest.raises(KeyError):
        lprop.RK3(sc.dt, sc.var())
    with pytest.raises(KeyError):
        lprop.rhs_default(sc.dt, sc.var())


def test_empty_and_tiny_ensembles(lprop):
    sc = scenarios.column_ensemble(1, seed=1, ngrid=101, sheared=True, amplitude=0.3)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    assert_state_close(lprop.RK3(sc.dt, sc.var()), orc.RK3(sc.dt, sc.var()), tag="n=1")
    # n = 0: the mean flow still evolves under Coriolis / pressure gradient
    empty = sc.var()
    for i in range(9):
        empty[i] = np.zeros(0)
    lprop.set_statics(dkk=np.zeros(0), dll=np.zeros(0), rr_mm_area=np.zeros(0))
    cfg = sc.oracle_cfg(); cfg.update(dkk=np.zeros(0), dll=np.zeros(0), rr_mm_area=np.zeros(0))
    got, want = lprop.RK3(sc.dt, empty), oracle.Oracle(cfg).RK3(sc.dt, empty)
    assert got[3].shape == (0,)
    assert np.array_equal(got[9], want[9]) and np.array_equal(got[10], want[10])


def test_fused_cg_rr_is_bit_identical_to_the_library_route(lprop):
    """The column kernels evaluate cg_rr with shared reciprocals and a hand-rolled IEEE sqrt/division
    (csrc/column_step.cu: cg_rr_fast).  It must round exactly like __ddiv_rn/__dsqrt_rn, i.e. like numpy."""
    import ctypes
    import torch
    from msgwam_b200._cabi import check, lib
    from msgwam_b200._engine import Engine
    eng = Engine.get()
    rng = np.random.default_rng(99)
    n = 2_000_000
    mm = -2 * np.pi / rng.uniform(50., 50e3, n) * rng.choice([-1., 1.], n)
    kh = 2 * np.pi / rng.uniform(2e3, 2000e3, n)
    th = rng.uniform(0, 2 * np.pi, n)
    kk, ll = kh * np.sin(th), kh * np.cos(th)
    phi = rng.uniform(-1.4, 1.4, n) * rng.choice([0., 1.], n)
    mm[:1000] *= 1e-9; kk[1000:2000] = 0.0; ll[1000:1500] = 0.0; mm[3000:3100] = 0.0    # degenerate wavenumbers
    ff = 2 * 7.2921e-5 * np.sin(phi)
    for bvf in (0.01, 0.02):
        want = oracle.Oracle(dict(bvf=bvf, phi0=0.0, grid=np.linspace(0, 1, 4), grids=np.array([.1, .2, .3]), dkk=1, dll=1,
                                  rr_mm_area=0)).lib
        t = [eng.dev(a) for a in (kk, ll, mm, ff)]
        out = eng.empty(n)
        check(lib.msgwam_debug_cg_rr_fast(*[eng.ptr(x) for x in t], bvf ** 2, eng.ptr(out), n, eng.stream))
        got = out.cpu().numpy()
        # the oracle's cg_rr1 takes ff directly: (-mm)*(om*om - f2)/om/vk
        f2 = ff * ff
        kh2 = kk * kk + ll * ll
        with np.errstate(all="ignore"):
            om = np.sqrt((bvf ** 2 * kh2 + f2 * (mm * mm)) / (kh2 + mm * mm))
            ref = (-mm) * (om * om - f2) / om / (kh2 + mm * mm)
        same = (got == ref) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (bvf, int((~same).sum()), got[~same][:3], ref[~same][:3])


@pytest.mark.parametrize("n,sheared,amplitude", [(1_000_000, False, None), (3_000_000, True, 0.3)])
def test_full_size_ensembles_vs_oracle(lprop, n, sheared, amplitude):
    """BASELINE configs[1] at its full size (1e6 ray volumes, constant N, zero wind, G = 1000: the benchmark workload,
    bit for bit the same ensemble) and a sheared, feeding-back 3e6-ray ensemble: the C oracle steps these in seconds,
    so parity at full size is checked directly, not through a proxy property."""
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=sheared, amplitude=amplitude)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
    got, want = lprop.RK3(sc.dt, sc.var()), orc.RK3(sc.dt, sc.var())
    assert_state_close(got, want, tag=sc.name, start=sc.var())
    # checksum of the deposit: the mean-flow increment summed over the column equals the oracle's
    du_g, du_w = np.sum(np.asarray(got[9]) - sc.uu), np.sum(want[9] - sc.uu)
    assert abs(du_g - du_w) <= 1e-12 * max(abs(du_w), np.max(np.abs(want[9])))
