"""Device-resident ray store: in-place multi-step advance, general (non-column) mode, and ray deletion by
stream compaction, against the CPU oracle / a numpy restatement of the deletion predicate."""
import numpy as np
import pytest

import oracle
from helpers import FIELDS, field_rel
from msgwam_b200 import scenarios

pytestmark = pytest.mark.gpu


def close(got, want, start, ray_tol=1e-13, grid_tol=1e-12):
    for i, nm in enumerate(FIELDS):
        g, w = np.asarray(got[i]), np.asarray(want[i])
        if nm in ("uu", "vv"):
            assert field_rel(g, w) <= grid_tol, (nm, field_rel(g, w))
        else:
            scale = np.maximum(np.abs(w), np.abs(w - start[i]))
            diff = np.abs(g - w)
            err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale))))
            assert err <= ray_tol, (nm, err)


@pytest.mark.parametrize("shuffled", [False, True])
def test_ensemble_in_place_steps_match_oracle(shuffled):
    from msgwam_b200.ensemble import RayEnsemble
    sc = scenarios.column_ensemble(60011, seed=21, ngrid=501, sheared=True, shuffled=shuffled, amplitude=0.3)
    ens = RayEnsemble.from_scenario(sc)
    orc = oracle.Oracle(sc.oracle_cfg())
    want = sc.var()
    for _ in range(4):
        want = orc.RK3(sc.dt, want)
    ens.step(sc.dt, 4)
    close(ens.to_var(), want, sc.var(), ray_tol=1e-12)


def test_ensemble_general_mode_with_online_saturation():
    from conftest import load_golden
    from helpers import scenario_from_npz
    from msgwam_b200.ensemble import RayEnsemble
    d = load_golden("random_col_sat.npz")
    sc = scenario_from_npz(d)
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt, 3)
    got = ens.to_var()
    for i, nm in enumerate(FIELDS):
        w = d["step3_" + nm]
        if nm in ("uu", "vv"):
            assert field_rel(got[i], w) <= 1e-12, nm
        else:
            assert np.max(np.abs(got[i] - w) / np.maximum(np.abs(w), 1e-300)) <= 1e-12, nm


def _keep_mask(sc, rr, drr, mm, m_crit):
    """numpy restatement of the deletion predicate: out_of_domain of wave_projection (L:124-130, grid := grids)
    or |m| >= m_crit."""
    dz = np.diff(sc.grids[:2])[0]
    nlow = ((rr - .5 * drr) / dz).astype(int)
    nup = ((rr + .5 * drr) / dz + 1.).astype(int)
    nzmax = len(sc.grids) - 2
    ood = ((nlow >= nzmax) & (nup >= nzmax)) | ((nlow <= 0) & (nup <= 0))
    return (~ood) & (np.abs(mm) < m_crit)


@pytest.mark.parametrize("n", [1, 4095, 4096, 4097, 250013])
def test_compaction_is_stable_and_exact(n):
    from msgwam_b200.ensemble import RayEnsemble, STATE, STATICS
    sc = scenarios.column_ensemble(n, seed=33, ngrid=201, sheared=True, shuffled=True, ztop_rays=120e3)   # ~1/6 above the top
    m_crit = float(np.quantile(np.abs(sc.state[7]), 0.8))
    ens = RayEnsemble.from_scenario(sc)
    keep = _keep_mask(sc, sc.state[3], sc.state[4], sc.state[7], m_crit)
    survivors = ens.compact(m_crit=m_crit)
    assert survivors == int(keep.sum()) and (n < 100 or 0 < survivors < n)
    ref = dict(zip(STATE + STATICS, list(sc.state) + [sc.dkk, sc.dll, sc.rr_mm_area]))
    for nm in STATE + STATICS:
        assert np.array_equal(ens.field(nm).cpu().numpy(), ref[nm][keep]), nm
    # the compacted store keeps stepping: same result as the oracle on the filtered arrays
    if survivors > 0:
        cfg = sc.oracle_cfg()
        cfg.update(dkk=sc.dkk[keep], dll=sc.dll[keep], rr_mm_area=sc.rr_mm_area[keep])
        var = np.empty(11, dtype=object)
        for i in range(9):
            var[i] = sc.state[i][keep]
        var[9], var[10] = sc.uu, sc.vv
        want = oracle.Oracle(cfg).RK3(sc.dt, var)
        ens.step(sc.dt, 1)
        close(ens.to_var(), want, var, ray_tol=1e-12)


def test_compaction_all_and_none():
    from msgwam_b200.ensemble import RayEnsemble
    sc = scenarios.column_ensemble(5000, seed=5, ngrid=201)
    ens = RayEnsemble.from_scenario(sc)
    assert ens.compact(m_crit=float("inf")) == int(_keep_mask(sc, sc.state[3], sc.state[4], sc.state[7], np.inf).sum())
    assert ens.compact(m_crit=0.0) == 0
    assert ens.compact(m_crit=1.0) == 0


def test_critical_level_run_with_periodic_deletion():
    """BASELINE configs[4] in miniature: jet, rays refracted towards critical levels / out of the top, deletion
    every 10 steps.  Oracle = reference RK3 on arrays (and statics) filtered by the same predicate."""
    from msgwam_b200.ensemble import RayEnsemble, STATE
    sc = scenarios.critical_level_ensemble(30011, ngrid=401)
    m_crit = 6.3e-3            # just above the initial maximum of |m| (2 pi / 1 km): refracted rays cross it within a few steps
    ens = RayEnsemble.from_scenario(sc)
    state = [a.copy() for a in sc.state]
    stat = [sc.dkk.copy(), sc.dll.copy(), sc.rr_mm_area.copy()]
    uu, vv = sc.uu.copy(), sc.vv.copy()
    deleted = 0
    for cycle in range(4):
        cfg = sc.oracle_cfg(); cfg.update(dkk=stat[0], dll=stat[1], rr_mm_area=stat[2])
        orc = oracle.Oracle(cfg)
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
                assert field_rel(got[i], var[i]) <= 1e-11, (cycle, nm, field_rel(got[i], var[i]))
            else:
                scale = np.maximum(np.abs(var[i]), np.abs(var[i] - start[i]))
                diff = np.abs(got[i] - var[i])
                err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale))))
                assert err <= 1e-11, (cycle, nm, err)
        keep = _keep_mask(sc, var[3], var[4], var[7], m_crit)
        survivors = ens.compact(m_crit=m_crit)
        assert survivors == int(keep.sum())
        deleted += int((~keep).sum())
        state = [np.asarray(var[i])[keep] for i in range(9)]
        stat = [a[keep] for a in stat]
        uu, vv = np.asarray(var[9]), np.asarray(var[10])
        for i, nm in enumerate(STATE):
            assert np.array_equal(ens.field(nm).cpu().numpy(), np.asarray(got[i])[keep]), (cycle, nm)
    assert deleted > 0 and ens.n > 0


def test_driver_loop_on_the_device_vs_reference_history():
    """RayEnsemble.advance = raytracer.py's loop (R:157-188: RK3, then saturation(direct=True)) without leaving the
    device, with a device-side history; compared with the history the unmodified driver produced (golden fixture)."""
    from conftest import load_golden
    from helpers import field_rel, max_rel
    from msgwam_b200.ensemble import History, RayEnsemble
    d = load_golden("driver_history.npz")
    sc = scenarios.default_column()
    ens = RayEnsemble.from_scenario(sc)
    hist = History(ens, nsnap=400, every=1)
    ens.advance(sc.dt, 360, saturate=True, history=hist)
    h = hist.to_host()
    assert list(h["steps"]) == list(range(361))
    steps = list(d["steps"])
    for nt in (0, 1, 2, 10, 100, 360):
        k = steps.index(nt)
        tol = 1e-13 if nt <= 10 else 1e-10
        for nm in ("dens", "rr", "mm", "drr", "dmm"):
            assert max_rel(h[nm][nt], d[nm][k]) <= tol, (nt, nm, max_rel(h[nm][nt], d[nm][k]))
        assert field_rel(h["uu"][nt], d["uu"][k]) <= max(tol, 1e-12), nt
    # advance(saturate=False) is plain stepping
    a, b = RayEnsemble.from_scenario(sc), RayEnsemble.from_scenario(sc)
    a.advance(sc.dt, 5, saturate=False); b.step(sc.dt, 5)
    for nm in ("rr", "mm", "dens"):
        assert np.array_equal(a.field(nm).cpu().numpy(), b.field(nm).cpu().numpy())


@pytest.mark.parametrize("amplitude,steps,tol", [(8.0, 1, 1e-10), (1.0, 3, 1e-10)])
def test_post_step_clamp_when_the_packet_saturates(amplitude, steps, tol):
    """The driver's packet never reaches saturation, so the clamp is exercised here: ensembles at and far above the
    static-instability amplitude, driver loop on the device vs the same loop through the oracle (R:157-188).  At eight
    times the threshold two thirds of the rays are clamped and the flow is violently unstable (wavenumbers change sign
    within a step; any 1e-16 difference grows ~1e3-fold per step, in the host-call loop just the same), hence one step
    (the deposit's fixed-point sums differ from the oracle's sequential fp64 sums by a few 1e-14 of the grid fields; the
    contract tolerance of 1e-10 is what is asserted)."""
    from helpers import max_rel
    from msgwam_b200.ensemble import RayEnsemble
    sc = scenarios.column_ensemble(30011, seed=41, ngrid=401, sheared=True, amplitude=amplitude)
    ens = RayEnsemble.from_scenario(sc)
    orc = oracle.Oracle(sc.oracle_cfg())
    var = sc.var()
    clamped = 0
    for _ in range(steps):
        out = orc.RK3(sc.dt, var)
        dens = orc.saturation(sc.dt, out[0], var[3], (out[3] - var[3]) / 1, var[4], (out[4] - var[4]) / sc.dt, out[5], out[6],
                              var[7], (out[7] - var[7]) / sc.dt, direct=True)
        clamped += int(np.count_nonzero(dens != out[0]))
        out[0] = dens
        var = out
    assert clamped > (1000 if amplitude > 2 else 0), clamped
    ens.advance(sc.dt, steps, saturate=True)
    got = ens.to_var()
    assert max_rel(got[0], var[0]) <= tol
    close(got, var, sc.var(), ray_tol=tol, grid_tol=10 * tol)


def test_full_driver_run_1440_steps_on_the_device():
    """The driver's whole run (nt_max = 1440, R:157-188) on the device, against the unmodified driver's history.
    Tolerances follow the reference's own noise floor (the reference against itself with the
    rays permuted differs by 5.9e-14 in m at step 720 and 4.3e-8 at step 1440)."""
    from conftest import load_golden
    from helpers import field_rel, max_rel
    from msgwam_b200.ensemble import History, RayEnsemble
    d = load_golden("driver_history.npz")
    sc = scenarios.default_column()
    ens = RayEnsemble.from_scenario(sc)
    hist = History(ens, nsnap=5, every=360)
    ens.advance(sc.dt, 1440, saturate=True, history=hist)
    h = hist.to_host()
    steps = list(d["steps"])
    for nt, tol in ((360, 1e-10), (720, 1e-10), (1440, 5e-6)):
        k, j = steps.index(nt), list(h["steps"]).index(nt)
        for nm in ("dens", "rr", "mm", "drr", "dmm"):
            assert max_rel(h[nm][j], d[nm][k]) <= tol, (nt, nm, max_rel(h[nm][j], d[nm][k]))
        assert field_rel(h["uu"][j], d["uu"][k]) <= tol, nt
    # the conservation diagnostic of the driver (R:198-240) on the final state: projected wave action
    import msgwam_b200.libprop as lprop
    sc.install(lprop)
    v = ens.to_var()
    wa = lprop.wave_projection(v[0], v[1], v[2], v[3] - .5 * v[4], v[3] + .5 * v[4], v[5], v[6], v[7] - .5 * v[8],
                               v[7] + .5 * v[8], sc.dkk, sc.dll, v[8], sc.grid, var=2)
    assert np.all(np.isfinite(wa)) and wa.max() > 0


class _LoopbackExchange:
    """world = 1 peer exchange: this GPU's own buffer is the only inbox, so the fused multi-GPU step
    (msgwam_column_step_p2p / msgwam_column_step_nz with peers: 16-byte self-validating cells, pass A pushes, the CTAs
    of pass B poll the cells their chain slices read) runs on one GPU and must reproduce the single-GPU step bit for bit."""

    def __init__(self, G):
        import torch
        from msgwam_b200 import _cabi
        self.inbox = torch.zeros(int(_cabi.lib.msgwam_p2p_inbox_doubles(G, 1)), dtype=torch.float64, device="cuda")
        self.epoch = 0

    def next(self, count=1):
        from msgwam_b200 import _cabi
        pe = _cabi.Peers()
        pe.world, pe.rank, pe.epoch = 1, 0, self.epoch + 1
        self.epoch += count
        pe.inbox[0] = self.inbox.data_ptr()
        return pe


@pytest.mark.parametrize("profile", [False, True])
@pytest.mark.parametrize("ngrid", [5, 150, 1001])
def test_peer_exchange_loopback_matches_the_single_gpu_step(profile, ngrid):
    import torch
    from msgwam_b200.ensemble import RayEnsemble
    sc = scenarios.column_ensemble(40003, seed=5, ngrid=ngrid, sheared=True, shuffled=(ngrid == 150), amplitude=0.3)
    if profile:
        sc.model = dict(sc.model, bvf=np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3)))))
    plain, looped = RayEnsemble.from_scenario(sc), RayEnsemble.from_scenario(sc)
    looped.exchange = _LoopbackExchange(looped.G)
    for _ in range(5):                                        # both parities of the inbox, several epochs
        plain.step(sc.dt); looped.step(sc.dt)
    plain.check_errors(); looped.check_errors()
    # per-ray state: identical given identical mean flow; the mean flow differs only by the order of the fp64 atomics
    # inside each launch (the deposits of the two runs are separate sums of the same contributions)
    for nm in ("rr", "drr", "mm", "dmm"):
        a, b = plain.field(nm), looped.field(nm)
        assert torch.allclose(a, b, rtol=1e-12, atol=0.0), nm
    for a, b in ((plain.uu, looped.uu), (plain.vv, looped.vv)):
        scale = float(a.abs().max()) or 1.0
        assert float((a - b).abs().max()) <= 1e-12 * scale
