"""Developer tool: host time per RayEnsemble.step / advance call against the GPU time of the step (small ensembles are
launch-bound).  usage: python tools/host_overhead.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble
for n in (60, 10_000, 1_000_000):
    sc = scenarios.column_ensemble(n, seed=1, ngrid=1001, sheared=True, amplitude=0.1)
    ens = RayEnsemble.from_scenario(sc)
    for name, fn in (("step", lambda: ens.step(sc.dt)), ("advance", lambda: ens.advance(sc.dt, 1)), ("step x100 in one call", None)):
        if fn is None:
            ens.step(sc.dt, 5); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); a.record(); ens.step(sc.dt, 100); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
            print("n=%d %-22s host %.1f us per step, GPU %.1f us per step" % (n, name, (t1 - t0) * 1e4, a.elapsed_time(b) * 10), flush=True)
            continue
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(100): fn()
        b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        print("n=%d %-22s host %.1f us per call, GPU %.1f us per call" % (n, name, (t1 - t0) * 1e4, a.elapsed_time(b) * 10), flush=True)
