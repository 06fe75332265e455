"""Developer tool: BASELINE configs[4] -- the critical-level / turning-point stress case with ray deletion by stream
compaction -- at full size on one GPU: step time before and after deletions, compaction time and throughput."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
t0 = time.perf_counter()
sc = scenarios.critical_level_ensemble(n, ngrid=1001, stress=True)      # the variant in which deletion happens every cycle
ens = RayEnsemble.from_scenario(sc)
del sc.state
print("ensemble of %d rays built and uploaded in %.1f s" % (n, time.perf_counter() - t0), flush=True)
dt, m_crit = 120.0, scenarios.M_CRIT_STRESS

def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), r

for cycle in range(4):
    ms_steps, _ = timed(lambda: ens.step(dt, 10))
    before = ens.n
    ms_c, after = timed(lambda: ens.compact(dt, m_crit))
    nbytes = int(ens._slab.shape[0]) * 8 * (before + after)  # every field of the store read, survivors written
    print("cycle %d: 10 steps %.1f ms (%.3e ray-steps/s) | compaction %d -> %d rays in %.2f ms (%.0f GB/s of field traffic)" % (
        cycle, ms_steps, 10 * before / (ms_steps * 1e-3), before, after, ms_c, nbytes / (ms_c * 1e-3) / 1e9), flush=True)
ens.check_errors()
print("mean flow finite:", bool(torch.isfinite(ens.uu).all()), " rays finite:", bool(torch.isfinite(ens.field("rr")).all() and torch.isfinite(ens.field("mm")).all()))
