"""Developer tool: step time of the non-column modes (stage-by-stage general path) next to the fused column step."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble

def timed(ens, dt, reps=5):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(dt); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
for mode in ("column", "saturate_online", "hprop", "nz_profile"):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=True, amplitude=0.3)
    if mode == "saturate_online": sc.model["saturate_online"] = True
    if mode == "hprop": sc.hprop = True
    if mode == "nz_profile":
        sc.model["bvf"] = np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3))))
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt, 2)
    us = timed(ens, sc.dt)
    print("%-16s n=%d  %.1f us per RK3 step  %.3e ray-steps/s" % (mode, n, us, n / us * 1e6), flush=True)
