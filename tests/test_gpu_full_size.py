"""Round-2 parity tests at the sizes BASELINE.json names: the fused N(z) step at 1e7 rays, the
configs[4] critical-level case at 5e6 rays with deletions that actually happen, the error word of the bounded
device-side waits, the host-buffer path of the N(z) step, and the cache-invalidation cases of ADVICE r01."""
import ctypes

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


def _ray_err(got, want, start):
    scale = np.maximum(np.abs(want), np.abs(want - start))
    diff = np.abs(got - want)
    return float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale)))) if got.size else 0.0


def test_fused_profile_step_at_1e7_rays_vs_oracle():
    """BASELINE configs[2] at its full size: 1e7 rays, N^2(z) profile, sheared wind, G = 1000, one fused two-sweep step
    (msgwam_column_step_nz) against the multi-thread oracle; then the sizes-independent property that the column sum
    of the mean-flow increment equals the oracle's (a checksum of the whole deposit)."""
    from msgwam_b200.ensemble import RayEnsemble
    sc = scenarios.nz_sheared_ensemble(10_000_000, seed=1234, amplitude=0.3)
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt)
    got = ens.to_var()
    want = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads()).RK3(sc.dt, sc.var())
    start = sc.var()
    for i, nm in enumerate(FIELDS):
        if nm in ("uu", "vv"):
            assert field_rel(got[i], want[i]) <= 1e-11, (nm, field_rel(got[i], want[i]))
        else:
            err = _ray_err(np.asarray(got[i]), np.asarray(want[i]), np.asarray(start[i]))
            assert err <= 1e-12, (nm, err)
    assert np.max(np.abs(got[4] - sc.state[4])) > 0 and np.max(np.abs(got[8] - sc.state[8])) > 0     # extents evolve under N(z)
    du_g, du_w = np.sum(got[9] - sc.uu), np.sum(want[9] - sc.uu)
    assert abs(du_g - du_w) <= 1e-11 * max(abs(du_w), np.max(np.abs(want[9])))


def _keep_mask(sc, rr, drr, mm, m_crit):
    dz = sc.grids[1] - sc.grids[0]
    nzmax = len(sc.grids) - 2
    with np.errstate(invalid="ignore"):
        nlow = np.trunc((rr - .5 * drr) / dz); nup = np.trunc((rr + .5 * drr) / dz + 1.)
    ood = ((nlow >= nzmax) & (nup >= nzmax)) | ((nlow <= 0) & (nup <= 0))
    return ~ood & (np.abs(mm) < m_crit)


def test_critical_level_stress_case_5e6_rays_with_real_deletions():
    """BASELINE configs[4] at 5e6 rays in the variant where deletion HAPPENS (sharp jet, rays on its flank and under
    the top): every cycle of 10 steps must delete rays through both predicates' union, 0 < survivors < n, and the
    survivors must equal the oracle's state filtered with the same mask (L:129-130 predicate or |m| >= m_crit)."""
    from msgwam_b200.ensemble import RayEnsemble, STATE
    n = 5_000_000
    sc = scenarios.critical_level_ensemble(n, ngrid=1001, stress=True)
    m_crit = scenarios.M_CRIT_STRESS
    ens = RayEnsemble.from_scenario(sc)
    state = [a.copy() for a in sc.state]
    stat = [sc.dkk.copy(), sc.dll.copy(), sc.rr_mm_area.copy()]
    uu, vv = sc.uu.copy(), sc.vv.copy()
    threads = oracle.max_threads()
    for cycle in range(3):
        cfg = sc.oracle_cfg(); cfg.update(dkk=stat[0], dll=stat[1], rr_mm_area=stat[2])
        orc = oracle.Oracle(cfg, nthreads=threads)
        var = np.empty(11, dtype=object)
        for i in range(9):
            var[i] = state[i]
        var[9], var[10] = uu, vv
        start = [a.copy() for a in var[:9]]
        for _ in range(10):
            var = orc.RK3(sc.dt, var)
        ens.step(sc.dt, 10)
        got = ens.to_var()
        for i, nm in enumerate(FIELDS):
            if nm in ("uu", "vv"):
                assert field_rel(got[i], var[i]) <= 1e-10, (cycle, nm, field_rel(got[i], var[i]))
            else:
                err = _ray_err(np.asarray(got[i]), np.asarray(var[i]), start[i])
                assert err <= 1e-10, (cycle, nm, err)
        before = ens.n
        keep = _keep_mask(sc, var[3], var[4], var[7], m_crit)
        survivors = ens.compact(m_crit=m_crit)
        assert 0 < survivors < before, (cycle, before, survivors)
        assert survivors == int(keep.sum()), (cycle, survivors, int(keep.sum()))
        state = [np.asarray(var[i])[keep] for i in range(9)]
        stat = [a[keep] for a in stat]
        uu, vv = np.asarray(var[9]), np.asarray(var[10])
        for i, nm in enumerate(STATE):
            if nm in ("rr", "mm", "dens"):                      # the changed slots, bit for bit the pre-deletion values of the survivors
                assert np.array_equal(ens.field(nm).cpu().numpy(), np.asarray(got[i])[keep]), (cycle, nm)
    assert ens.n < n


def test_error_word_raises_at_the_next_synchronising_call():
    """A bounded device-side wait that times out sets the error word in the work buffer; every method that synchronises
    anyway (to_var, compact, History.to_host, check_errors) must raise instead of returning invalid results."""
    from msgwam_b200 import _cabi
    from msgwam_b200.ensemble import History, RayEnsemble
    sc = scenarios.column_ensemble(5003, seed=3, ngrid=201, sheared=True, amplitude=0.3)
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt)
    off = int(_cabi.lib.msgwam_column_error_offset(ens.G))
    for call in (ens.to_var, lambda: ens.compact(sc.dt), ens.check_errors, lambda: History(ens, 1).to_host()):
        ens.work[off] = 1.0
        with pytest.raises(_cabi.MsgwamError):
            call()
        assert float(ens.work[off].item()) == 0.0            # cleared: the ensemble stays usable
    ens.step(sc.dt)
    assert np.isfinite(ens.to_var()[9]).all()


@pytest.mark.parametrize("path", ["device", "host", "host_frozen"])
def test_changed_statics_between_same_size_calls_take_effect(lprop, path):
    """two RK3 calls with the same n but different dkk must not share derived statics (pkl = dkk * dll).
    device: torch CUDA tensors in the 11 slots; host: numpy arrays, default semantics (statics re-read every call,
    even after an in-place edit); host_frozen: after freeze_statics() a new set_statics() is what refreshes them."""
    import torch
    sc = scenarios.column_ensemble(20011, seed=5, ngrid=301, sheared=True, amplitude=0.3)
    sc.install(lprop)
    var = sc.var()
    if path == "device":
        host, var = var, np.empty(11, dtype=object)
        for i in range(11):
            var[i] = torch.as_tensor(host[i], device="cuda")
    orc = oracle.Oracle(sc.oracle_cfg())
    if path == "host_frozen":
        lprop.freeze_statics()
    first = lprop.RK3(sc.dt, var)
    want = orc.RK3(sc.dt, sc.var())
    assert field_rel(np.asarray(first[9].cpu() if path == "device" else first[9]), want[9]) <= 1e-12
    dkk2 = sc.dkk * 3.0
    if path == "host":
        lprop.statics["dkk"][:] = dkk2                     # in place: the reference would see it at its next rhs call
    else:
        lprop.set_statics(dkk=dkk2.copy())
        if path == "host_frozen":
            lprop.freeze_statics()
    second = lprop.RK3(sc.dt, var)
    cfg2 = sc.oracle_cfg(); cfg2["dkk"] = dkk2
    want2 = oracle.Oracle(cfg2).RK3(sc.dt, sc.var())
    got_u = np.asarray(second[9].cpu() if path == "device" else second[9])
    assert field_rel(got_u, want2[9]) <= 1e-12, field_rel(got_u, want2[9])
    assert field_rel(got_u, want[9]) > 1e-9                    # and it differs from the first call's result


def test_profile_step_through_host_buffers_vs_oracle(lprop):
    """lprop.RK3 with numpy inputs and an N(z) profile goes through msgwam_rk3_column_nz_host: inputs up, the fused
    profile step, rr, drr, mm, dmm, uu, vv back; unchanged slots come back as views of the inputs."""
    sc = scenarios.nz_sheared_ensemble(150_001, seed=11, amplitude=0.3)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
    got = want = sc.var()
    for step in range(2):
        got, want = lprop.RK3(sc.dt, got), orc.RK3(sc.dt, want)
        assert_state_close(got, want, ray_tol=1e-12, grid_tol=1e-11, tag="step %d" % (step + 1), start=sc.var())
    assert isinstance(got[3], np.ndarray) and not got[0].flags.writeable


def test_compaction_reads_any_non_zero_flag_byte_as_keep():
    """msgwam_compact's tile counts and its scatter must agree on what a flag byte means (non-zero keeps)."""
    import torch
    from msgwam_b200._cabi import check, lib
    n = 70_003
    rng = np.random.default_rng(9)
    flags = rng.choice(np.array([0, 1, 2, 128, 255], dtype=np.uint8), n)
    keep = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    keep[:n] = torch.as_tensor(flags, device="cuda")
    src = torch.arange(n, dtype=torch.float64, device="cuda")
    dst = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    scratch = torch.empty(int(lib.msgwam_compact_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
    vp = ctypes.c_void_p
    check(lib.msgwam_compact(n, vp(keep.data_ptr()), 1, (vp * 1)(src.data_ptr()), (vp * 1)(dst.data_ptr()), vp(count.data_ptr()),
                             vp(scratch.data_ptr()), vp(torch.cuda.current_stream().cuda_stream)), "msgwam_compact")
    want = np.flatnonzero(flags != 0).astype(np.float64)
    assert int(count.item()) == len(want)
    assert np.array_equal(dst[:len(want)].cpu().numpy(), want)


@pytest.mark.parametrize("profile", [False, True])
def test_fixed_point_histogram_bounds_protocol(profile):
    """The CTA histogram of the deposit accumulates in 64-bit fixed point, scaled by bounds of the deposited flux
    (msgwam_rays_t.bounds): before the first step of an ensemble a pre-pass measures them, every step leaves the bounds
    of its own deposits; steps stay within the grid tolerance on a shuffled ensemble; an edit of the store through torch
    triggers a new pre-pass; and a flux that grows 20-fold BEHIND the ensemble's back (stale bounds) still gives the
    right answer -- rays that no longer fit the scale go to the global deposit in fp64, no accumulator can overflow."""
    from msgwam_b200.ensemble import RayEnsemble
    mk = scenarios.nz_sheared_ensemble if profile else (lambda n, **kw: scenarios.column_ensemble(n, ngrid=1001, sheared=True, **kw))
    sc = mk(150_007, seed=21, amplitude=0.02, shuffled=True)    # 0.02 -> 0.09 -> 0.4 after the two edits below: never chaotic
    ens = RayEnsemble.from_scenario(sc)
    assert float(ens._bounds.abs().sum()) == 0.0
    ens.step(sc.dt)
    b1 = ens._bounds.cpu().numpy().copy()
    assert (b1[:6] > 0).all() and (b1[6:12] == 0).all() and b1[12] == 1.0, b1
    ens.step(sc.dt, 3)                                          # fixed-point steps
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
    want = sc.var()
    for _ in range(4):
        want = orc.RK3(sc.dt, want)
    assert_state_close(ens.to_var(), want, ray_tol=1e-12, grid_tol=1e-11, tag="fixed point", start=sc.var())
    # an edit through torch invalidates the bounds: the next step measures them again
    ens.field("dens").mul_(20.0)
    want[0] = want[0] * 20.0
    ens.step(sc.dt)
    want = orc.RK3(sc.dt, want)
    b2 = ens._bounds.cpu().numpy()
    assert np.all(b2[:6] > 10.0 * b1[:6]), (b1, b2)
    assert_state_close(ens.to_var(), want, ray_tol=1e-11, grid_tol=1e-10, tag="after an edit", start=sc.var())
    # the same growth behind the ensemble's back (version counter restored by hand: no pre-pass, stale bounds)
    ens.field("dens").mul_(20.0)
    want[0] = want[0] * 20.0
    ens._slab_version = ens._slab._version
    ens.step(sc.dt, 2)
    want = orc.RK3(sc.dt, orc.RK3(sc.dt, want))
    ens.check_errors()
    assert_state_close(ens.to_var(), want, ray_tol=1e-10, grid_tol=1e-10, tag="stale bounds", start=sc.var())


@pytest.mark.parametrize("profile", [False, True])
def test_step_does_not_need_all_ctas_resident(profile):
    """the mean-flow chain of the second sweep waits on a grid-wide arrival counter; that must not assume that
    every CTA of the grid is resident (kernels of other streams or MPS clients may hold SMs).  With the test hook
    msgwam_debug_grid_mult(3) every sweep is launched with three CTAs per SM, of which one fits: two thirds of the grid
    have not started while the first third waits for the chain.  The slices of the chain are handed out by ticket to
    whichever CTAs run, so the steps complete, the error word stays clear and the result is the oracle's."""
    from msgwam_b200 import _cabi
    from msgwam_b200.ensemble import RayEnsemble
    mk = scenarios.nz_sheared_ensemble if profile else (lambda n, **kw: scenarios.column_ensemble(n, ngrid=1001, sheared=True, **kw))
    sc = mk(300_011, seed=5, amplitude=0.3)
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
    want = orc.RK3(sc.dt, orc.RK3(sc.dt, sc.var()))
    assert _cabi.lib.msgwam_debug_grid_mult(3) == 0
    try:
        ens = RayEnsemble.from_scenario(sc)
        ens.step(sc.dt, 2)
        got = ens.to_var()                      # synchronises and raises on a set error word (code 2: chain wait timed out)
    finally:
        assert _cabi.lib.msgwam_debug_grid_mult(1) == 0
    assert_state_close(got, want, ray_tol=1e-12, grid_tol=1e-11, tag="3 CTAs per SM", start=sc.var())


@pytest.mark.parametrize("profile,amplitude", [(False, 1.0), (True, 1.0), (True, 0.05)])
def test_fused_advance_matches_the_driver_loop_through_the_oracle(profile, amplitude):
    """RayEnsemble.advance in the column modes = msgwam_column_advance / _nz: the RK3 step with the driver's post-step
    clamp saturation(direct=True) (R:182-188, `/ 1` included) fused into the second sweep.  Oracle: the same loop
    through RK3 and saturation of the CPU restatement; with amplitude 1 a good part of the rays is clamped."""
    from msgwam_b200.ensemble import RayEnsemble
    mk = scenarios.nz_sheared_ensemble if profile else (lambda n, **kw: scenarios.column_ensemble(n, ngrid=1001, sheared=True, **kw))
    sc = mk(120_013, seed=33, amplitude=amplitude)
    ens = RayEnsemble.from_scenario(sc)
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
    var = sc.var()
    clamped = 0
    steps = 3
    for _ in range(steps):
        out = orc.RK3(sc.dt, var)
        dens = orc.saturation(sc.dt, out[0], var[3], (out[3] - var[3]) / 1, var[4], (out[4] - var[4]) / sc.dt, out[5], out[6],
                              var[7], (out[7] - var[7]) / sc.dt, direct=True)
        clamped += int(np.count_nonzero(dens != out[0]))
        out[0] = dens
        var = out
    assert (clamped > 1000) == (amplitude >= 1.0), clamped
    ens.advance(sc.dt, steps, saturate=True)
    assert_state_close(ens.to_var(), var, ray_tol=1e-10, grid_tol=1e-10, tag="advance", start=sc.var())


def test_frozen_background_mode_vs_reference_fixture_and_oracle(lprop):
    """Extension "M2" (msgwam_column_step_frozen: all three RK stages in registers, mean flow frozen over the step, one
    deposit and one mean-flow update per step) against the fixture composed from the unmodified reference's functions
    through model_config['rhs'] (tests/golden/make_golden_frozen.py), and in place at 3e5 rays against the oracle."""
    from conftest import load_golden
    from helpers import scenario_from_npz
    from msgwam_b200.ensemble import RayEnsemble
    d = load_golden("frozen_col.npz")
    sc = scenario_from_npz(d)
    sc.install(lprop)
    var = sc.var()
    for step in (1, 2, 3):
        var = lprop.RK3_frozen(sc.dt, var)
        assert_state_close(var, [d["step%d_%s" % (step, nm)] for nm in FIELDS], ray_tol=1e-12, grid_tol=1e-11,
                           tag="frozen fixture step %d" % step, start=sc.var())
    # the plug-in route: reference-style RK3 with model_config['rhs'] = rhs_frozen advances the rays identically
    lprop.set_model_setup(rhs=lprop.rhs_frozen)
    rays_only = lprop.RK3(sc.dt, sc.var())
    for i in (3, 7):
        assert np.max(np.abs(np.asarray(rays_only[i]) - d["step1_" + FIELDS[i]]) / np.abs(d["step1_" + FIELDS[i]])) <= 1e-12
    assert np.array_equal(np.asarray(rays_only[9]), sc.uu)
    big = scenarios.column_ensemble(300_007, seed=8, ngrid=1001, sheared=True, amplitude=0.3)
    ens = RayEnsemble.from_scenario(big)
    ens.step_frozen(big.dt, 3)
    orc = oracle.Oracle(big.oracle_cfg(), nthreads=oracle.max_threads())
    want = big.var()
    for _ in range(3):
        want = orc.RK3_frozen(big.dt, want)
    assert_state_close(ens.to_var(), want, ray_tol=1e-12, grid_tol=1e-11, tag="frozen ensemble", start=big.var())
    # and it is a different scheme: the coupled step gives a different wind
    ens2 = RayEnsemble.from_scenario(big)
    ens2.step(big.dt, 3)
    assert field_rel(ens2.to_var()[9], want[9]) > 1e-8
