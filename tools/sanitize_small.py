"""Developer tool: every column kernel on small ensembles, for compute-sanitizer (memcheck / racecheck, one tool per run).
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble
import msgwam_b200.libprop as lprop
for mk in (lambda: scenarios.column_ensemble(5003, seed=1, ngrid=201, sheared=True, amplitude=1.0),
           lambda: scenarios.nz_sheared_ensemble(5003, seed=2, ngrid=201, amplitude=1.0),
           lambda: scenarios.column_ensemble(4001, seed=3, ngrid=201, sheared=True, amplitude=0.3, shuffled=True)):
    sc = mk()
    ens = RayEnsemble.from_scenario(sc)
    ens.step(sc.dt, 3)
    ens.advance(sc.dt, 2)
    if np.ndim(sc.model["bvf"]) == 0:
        ens.step_frozen(sc.dt, 2)
    ens.compact(sc.dt, float(np.quantile(np.abs(sc.state[7]), 0.7)))
    ens.step(sc.dt, 2)
    out = ens.to_var()
    sc.install(lprop)
    got = lprop.RK3(sc.dt, sc.var())
    assert np.isfinite(out[9]).all() and np.isfinite(np.asarray(got[9])).all()
    print(sc.name, "ok", ens.n, flush=True)
torch.cuda.synchronize()
print("done")
