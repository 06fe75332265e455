"""Developer tool: GPU time per step of RayEnsemble.step vs RayEnsemble.advance (RK3 + the driver's post-step clamp,
raytracer.py:182-188) at 1e6 rays, in place."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
for amp in (0.3, 3.0):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=True, amplitude=amp)
    for name in ("step", "advance"):
        ens = RayEnsemble.from_scenario(sc)
        fn = (lambda k: ens.step(sc.dt, k)) if name == "step" else (lambda k: ens.advance(sc.dt, k, saturate=True))
        fn(30)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(50); b.record(); torch.cuda.synchronize()
        print("amplitude %.1f  %-8s %.1f us per step (dispersed ensemble, no L2 flush)" % (amp, name, a.elapsed_time(b) * 1e3 / 50), flush=True)
