"""Edge cases of the fused column step (csrc/column_step.cu, deposit.cuh) against the CPU oracle (B200 box; -m gpu).

Covered: the distributed mean-flow chain at every kind of grid size (fewer levels than CTAs, not a multiple of the
CTA count, the largest the kernels take), its rare-operand route (winds so small or so large that the exact
invariant-divisor division is not proven -> IEEE divisions), a non-uniform abscissa for gradients(), the deposit's
outlier route (a few rays of a warp far from the rest, at the low end, the high end, and everywhere), repeated steps
through the hand-over between the two sweeps, and empty ensembles.
"""
import numpy as np
import pytest

import oracle
from helpers import FIELDS, field_rel
from msgwam_b200 import scenarios
from test_gpu_parity import assert_state_close

pytestmark = pytest.mark.gpu


@pytest.fixture()
def lprop():
    import importlib
    import msgwam_b200.libprop as lp
    importlib.reload(lp)
    return lp


def run_both(lprop, sc, steps=2, ray_tol=1e-13, grid_tol=1e-12):
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    got = want = sc.var()
    start = sc.var()
    for s in range(steps):
        got, want = lprop.RK3(sc.dt, got), orc.RK3(sc.dt, want)
        assert_state_close(got, want, ray_tol=ray_tol, grid_tol=grid_tol, tag="%s step %d" % (sc.name, s + 1), start=start)
    return got, want


@pytest.mark.parametrize("ngrid", [4, 5, 8, 34, 149, 150, 297, 1001])
def test_chain_slices_for_every_kind_of_grid_size(lprop, ngrid):
    """G = ngrid - 1 levels are dealt out to 148 CTAs: G < 148 (most CTAs idle), G = 148, G = 149 (two levels in the first
    CTAs, the last ones idle), G = 296, G = 1000."""
    sc = scenarios.column_ensemble(20011, seed=100 + ngrid, ngrid=ngrid, sheared=True, amplitude=0.3)
    run_both(lprop, sc, steps=3)


def test_largest_grid_the_column_kernels_take(lprop):
    from msgwam_b200._cabi import lib
    gmax = int(lib.msgwam_column_max_levels())
    assert gmax >= 1000
    sc = scenarios.column_ensemble(30011, seed=5, ngrid=gmax + 1, sheared=True, amplitude=0.3)
    # ~2400 levels of 41 m: the noise of the deposit (fixed-point quantum, summation order) reaches dm/dt through
    # diff(u) / dz with u carrying dt / rhobar / dz times it -- 1 / dz^2 in all, 6x the 1000-level cases
    run_both(lprop, sc, steps=2, ray_tol=1e-12, grid_tol=1e-11)


def test_grid_taller_than_the_column_kernels_take_falls_back_to_the_general_path(lprop):
    from msgwam_b200._cabi import lib
    from msgwam_b200.ensemble import RayEnsemble
    gmax = int(lib.msgwam_column_max_levels())
    sc = scenarios.column_ensemble(20011, seed=6, ngrid=gmax + 60, sheared=True, amplitude=0.3)
    # ~2500 levels of 40 m: the summation-order noise of the deposit enters dm/dt through du/dz = diff(u) / dz, where
    # u itself carries dt / rhobar / dz times the noise -- 1 / dz^2 in all, 6x the noise of the 1000-level cases
    got, want = run_both(lprop, sc, steps=2, ray_tol=1e-12, grid_tol=1e-11)
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt, 2)
    assert_state_close(ens.to_var(), want, ray_tol=1e-12, grid_tol=1e-11, tag="ensemble", start=sc.var())


@pytest.mark.parametrize("scale", [1e-262, 1e-255, 1e252])
def test_chain_rare_operand_route(lprop, scale):
    """Winds of 1e-255 m/s (differences below 1e-250: the invariant-divisor form is not proven there) and 1e252 m/s
    (above 1e250) send the chain and the table build through the IEEE-division route.  The huge wind also drives the
    wavenumbers out of cg_rr's fast range (library route) and on to inf/nan, exactly like the reference."""
    sc = scenarios.column_ensemble(5003, seed=9, ngrid=201, sheared=True, amplitude=None)
    sc.uu = sc.uu * scale
    sc.vv = sc.vv * scale
    sc.pressure_gradient = sc.pressure_gradient * scale
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    with np.errstate(all="ignore"):
        got, want = lprop.RK3(sc.dt, sc.var()), orc.RK3(sc.dt, sc.var())
    for i, nm in enumerate(FIELDS):
        if nm not in ("rr", "mm", "uu", "vv") and scale > 1:
            # known deviation (arithmetic contract iv): once rr or mm of a ray is non-finite the reference
            # poisons its other slots too (0 / nan in lam_st, phi_st; nan - nan in ddrr_st; False * nan in dens_st),
            # while the column path returns those slots untouched
            continue
        g, w = np.asarray(got[i], dtype=np.float64), np.asarray(want[i], dtype=np.float64)
        assert np.array_equal(np.isnan(g), np.isnan(w)), nm
        ok = np.isfinite(w)
        assert np.array_equal(np.isinf(g), np.isinf(w)), nm
        if nm in ("uu", "vv"):
            # the deposit (summation order) is negligible against such winds or dominates tiny ones: compare to the field
            if ok.any():
                assert field_rel(g[ok], w[ok]) <= 1e-12, (nm, scale)
        elif ok.any():
            scale_i = np.maximum(np.abs(w[ok]), 1e-300)
            assert np.max(np.abs(g[ok] - w[ok]) / scale_i) <= 1e-13, (nm, scale)


def test_non_uniform_abscissa_for_gradients(lprop):
    """gradients() interpolates on grid[1:-1] (L:349-356); with unequal spacing the slopes need true divisions and the
    interval search has to walk.  (The deposit keeps using dz = grid[1] - grid[0], as the reference does.)"""
    sc = scenarios.column_ensemble(20011, seed=13, ngrid=301, sheared=True, amplitude=0.3)
    rng = np.random.default_rng(2)
    g = sc.grid.copy()
    g[2:-1] += rng.uniform(-20., 20., g.size - 3)          # keep grid[0], grid[1] (dz) and the top
    sc.grid = g
    run_both(lprop, sc, steps=2)


@pytest.mark.parametrize("where", ["high", "low", "both", "half"])
def test_deposit_outlier_lanes(lprop, where):
    """An ordered ensemble in which some rays of every warp sit many cells away from their neighbours: the warp window
    must keep serving the majority while the outliers go through the CTA histogram."""
    n = 40009
    sc = scenarios.column_ensemble(n, seed=17, ngrid=501, sheared=True, amplitude=0.3)
    rr = sc.state[3].copy()
    rng = np.random.default_rng(3)
    if where == "half":
        idx = np.arange(0, n, 2)
        shift = 9000.
    else:
        idx = rng.choice(n, n // 40, replace=False)
        shift = {"high": 7000., "low": -7000., "both": 7000.}[where]
    delta = np.full(idx.size, shift)
    if where == "both":
        delta *= rng.choice([-1., 1.], idx.size)
    rr[idx] = np.clip(rr[idx] + delta, 200., 95e3)
    sc.state[3] = rr
    run_both(lprop, sc, steps=2)


def test_ten_steps_through_the_hand_over(lprop):
    """Ten consecutive steps: the stage-1 hand-over buffer and the chain counter are reused every step."""
    sc = scenarios.column_ensemble(50021, seed=19, ngrid=401, sheared=True, amplitude=0.3)
    run_both(lprop, sc, steps=10, ray_tol=1e-12, grid_tol=1e-11)


def test_device_tensors_in_place_and_empty(lprop):
    """torch CUDA inputs (zero-copy path) and an ensemble without rays: the mean flow still advances (pressure
    gradient and Coriolis terms), rays stay empty."""
    import torch
    sc = scenarios.column_ensemble(7001, seed=23, ngrid=201, sheared=True, amplitude=0.3)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    var = sc.var()
    dev = np.empty(11, dtype=object)
    for i in range(11):
        dev[i] = torch.from_numpy(np.ascontiguousarray(var[i])).cuda()
    got = lprop.RK3(sc.dt, dev)
    want = orc.RK3(sc.dt, var)
    assert all(isinstance(x, torch.Tensor) and x.is_cuda for x in got)
    assert_state_close([x.cpu().numpy() for x in got], want, start=var)
    # no rays at all
    empty = sc.var()
    for i in range(9):
        empty[i] = empty[i][:0]
    lprop.set_statics(dkk=sc.dkk[:0], dll=sc.dll[:0], rr_mm_area=sc.rr_mm_area[:0])
    cfg = sc.oracle_cfg()
    cfg.update(dkk=sc.dkk[:0], dll=sc.dll[:0], rr_mm_area=sc.rr_mm_area[:0])
    got0, want0 = lprop.RK3(sc.dt, empty), oracle.Oracle(cfg).RK3(sc.dt, empty)
    assert got0[3].size == 0
    assert np.array_equal(np.asarray(got0[9]), want0[9]) and np.array_equal(np.asarray(got0[10]), want0[10])
