"""Developer tool: GPU time of consecutive IN-PLACE column steps (the ensemble evolves, its spatial order degrades)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200._cabi import check, lib
from msgwam_b200.ensemble import RayEnsemble
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for sheared, amp in ((False, None), (True, 0.3)):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=sheared, amplitude=amp)
    ens = RayEnsemble.from_scenario(sc)
    eng = ens.eng; P = eng.ptr
    p = ens.params(sc.dt); g = eng.grid_struct(ens.grid_devs); rays = ens._rays()
    rr, mm = ens.field("rr"), ens.field("mm")
    ts = []
    for k in range(nsteps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        check(lib.msgwam_column_step(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr), P(mm), P(ens._uu2), P(ens._vv2), eng.stream))
        b.record(); torch.cuda.synchronize()
        ens.uu, ens._uu2 = ens._uu2, ens.uu; ens.vv, ens._vv2 = ens._vv2, ens.vv
        ts.append(round(a.elapsed_time(b) * 1e3, 1))
    r = rr.cpu().numpy()
    cells = (r / 100.0).astype(int).reshape(-1, 32) if n % 32 == 0 else (r[: n // 32 * 32] / 100.0).astype(int).reshape(-1, 32)
    print("sheared", sheared, "amp", amp, "GPU us per in-place step:", ts[:6], "...", ts[-6:], " warp cell-span median/90%% after %d steps: %d / %d" % (
        nsteps, np.median(cells.max(1) - cells.min(1)), np.percentile(cells.max(1) - cells.min(1), 90)), flush=True)
