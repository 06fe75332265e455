"""Developer tool: the N(z) column step on a dispersed ensemble, for ncu captures (the last step's two launches).
usage: python tools/nz_capture.py <rays> <steps> [const]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 34
const = len(sys.argv) > 3 and sys.argv[3] == "const"
sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=True, amplitude=0.01)
if not const:
    sc.model = dict(sc.model, bvf=np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3)))))
ens = RayEnsemble.from_scenario(sc)
ens.step(sc.dt, steps)
torch.cuda.synchronize()
ens.check_errors()
print("ok", bool(torch.isfinite(ens.uu).all()))
