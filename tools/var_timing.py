"""Developer tool: in-place step time of the dispersed ensembles (the product's steady state) for the library named by
MSGWAM_B200_LIB: constant N at 1e6 (L2 flushed) and 1e7 rays, N(z) at 1.25e7 rays.  usage: python tools/var_timing.py [which]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble
which = sys.argv[1] if len(sys.argv) > 1 else "all"
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
def timed(ens, dt, k, fl):
    ts = []
    for _ in range(k):
        if fl: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(dt); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ts)
out = []
cases = [("const1e6", lambda: scenarios.column_ensemble(1_000_000, seed=1234, ngrid=1001), True),
         ("const1e7", lambda: scenarios.column_ensemble(10_000_000, seed=1234, ngrid=1001), False),
         ("nz1.25e7", lambda: scenarios.nz_sheared_ensemble(12_500_000, seed=1234), False),
         ("pileup2e7", lambda: scenarios.critical_level_ensemble(20_000_000, ngrid=1001, stress=True), False)]
for name, mk, fl in cases:
    if which != "all" and which not in name: continue
    sc = mk(); ens = RayEnsemble.from_scenario(sc); del sc.state
    t0 = timed(ens, 120.0, 3, fl)
    ens.step(120.0, 30)
    t = timed(ens, 120.0, 10, fl)
    ens.check_errors()
    out.append("%s first %.3f dispersed %.3f ms" % (name, t0, t))
    del ens
print(os.path.basename(os.environ.get("MSGWAM_B200_LIB", "default")), " | ".join(out), flush=True)
