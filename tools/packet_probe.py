"""Developer tool: a LOCALISED wave packet (nearly all of the flux in a few per cent of the ray index range) -- step time
and parity; the overflow guard of the fixed-point deposit must not push such ensembles onto the slow fp64 path."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble
import oracle
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
for width in (1.0, 0.05, 0.01, 0.002):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=True, amplitude=0.05)
    if width < 1.0:
        x = (np.arange(n) / n - 0.4) / width
        sc.state[0] = sc.state[0] * (np.exp(-0.5 * x * x) + 1e-12) * (1.0 / width)
    ens = RayEnsemble.from_scenario(sc)
    ts = []
    for _ in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(sc.dt); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    t = [round(a.elapsed_time(b), 3) for a, b in ts]
    print("packet width %.3f of the index range: ms per step %s" % (width, t), flush=True)
# parity on a small localized packet
sc = scenarios.column_ensemble(200_003, seed=5, ngrid=1001, sheared=True, amplitude=0.05)
x = (np.arange(sc.n) / sc.n - 0.4) / 0.01
sc.state[0] = sc.state[0] * (np.exp(-0.5 * x * x) + 1e-12) * 100.0
ens = RayEnsemble.from_scenario(sc)
ens.step(sc.dt, 3)
got = ens.to_var()
orc = oracle.Oracle(sc.oracle_cfg(), nthreads=oracle.max_threads())
want = sc.var()
for _ in range(3): want = orc.RK3(sc.dt, want)
print("parity, packet 0.01: rr %.2e mm %.2e uu %.2e" % (np.max(np.abs(got[3] - want[3]) / np.abs(want[3])), np.max(np.abs(got[7] - want[7]) / np.abs(want[7])),
      np.max(np.abs(got[9] - want[9])) / np.max(np.abs(want[9]))))
