"""Developer tool (round 2): fixed-point CTA histogram for the outlier lanes of the deposit (deposit.cuh), switched by
the developer hook msgwam_debug_fx_scale -- step time of a dispersed ensemble with and without it, and parity against
the oracle with it.   usage: python tools/fx_probe.py <nz|const> <rays> [steps_disperse]"""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios, _cabi
from msgwam_b200.ensemble import RayEnsemble
import oracle

lib = _cabi.lib
lib.msgwam_debug_fx_scale.argtypes = [ctypes.c_double]
mode = sys.argv[1] if len(sys.argv) > 1 else "const"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
k_disp = int(sys.argv[3]) if len(sys.argv) > 3 else 30


def make(n, amplitude=None, seed=1234):
    if mode == "nz":
        return scenarios.nz_sheared_ensemble(n, seed=seed, amplitude=0.01 if amplitude is None else amplitude)
    return scenarios.column_ensemble(n, seed=seed, ngrid=1001, sheared=amplitude is not None, amplitude=amplitude)


def scale_for(sc):
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    kh2 = kk * kk + ll * ll
    n2 = np.max(np.asarray(sc.model["bvf"], dtype=float)) ** 2
    om = np.sqrt(n2 * kh2 / (kh2 + mm * mm))
    cg = np.abs(mm) * om / (kh2 + mm * mm)
    b = float(np.sum(np.abs(sc.dkk * sc.dll * dmm) * cg * (np.abs(kk) + np.abs(ll)) * np.abs(dens)))
    return 2.0 ** np.floor(np.log2(2.0 ** 58 / b)), b


flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def timed(ens, dt, k):
    ts = []
    for _ in range(k):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(dt); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ts)


import msgwam_b200.ensemble as ens_mod

# ---- parity with the fixed-point histogram on (2e5 rays, 4 steps, feeding back; the first step measures the bounds) ----
for shuffled in (False, True):
    scp = make(200_003, amplitude=0.3, seed=77)
    if shuffled:
        perm = np.random.default_rng(5).permutation(scp.n)
        scp.state = [a[perm] for a in scp.state]; scp.dkk, scp.dll, scp.rr_mm_area = scp.dkk[perm], scp.dll[perm], scp.rr_mm_area[perm]
    res = {}
    for fixed in (True, False):
        ens_mod._FIXED_POINT_HISTOGRAM = fixed
        ens = RayEnsemble.from_scenario(scp)
        ens.step(scp.dt, 4)
        res[fixed] = ens.to_var()
        if fixed:
            print("bounds after 4 steps:", ens._bounds.cpu().numpy()[:6])
    want = scp.var()
    orc = oracle.Oracle(scp.oracle_cfg(), nthreads=oracle.max_threads())
    for _ in range(4):
        want = orc.RK3(scp.dt, want)
    def errs(g):
        e_r = max(float(np.max(np.abs(g[i] - want[i]) / np.maximum(np.abs(want[i]), 1e-300))) for i in (3, 7))
        e_g = max(float(np.max(np.abs(g[i] - want[i])) / np.max(np.abs(want[i]))) for i in (9, 10))
        return e_r, e_g
    print("parity (%s, shuffled=%s): fixed-point ray %.2e grid %.2e | fp64 CAS ray %.2e grid %.2e" % (
        mode, shuffled, *errs(res[True]), *errs(res[False])), flush=True)

# ---- timing ----
sc = make(n)
for fixed, allh in ((False, 0), (True, 0), (False, 0), (True, 0)):
    ens_mod._FIXED_POINT_HISTOGRAM = fixed
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt)
    t_first = timed(ens, sc.dt, 3)
    ens.step(sc.dt, k_disp)
    t = timed(ens, sc.dt, 10)
    ens.check_errors()
    print("%s %d rays, %-22s: ordered (steps 2-4) %.3f ms, dispersed (after %d steps) %.3f ms per step" % (
        mode, n, ("fixed-point, ALL lanes" if allh else "fixed-point histogram") if fixed else "fp64 CAS histogram", t_first, k_disp, t), flush=True)
    del ens
