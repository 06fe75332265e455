"""Per-kernel CUDA-event timings of the column step at several ensemble sizes (developer tool)."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200._engine import Engine
from msgwam_b200._cabi import check, lib
from msgwam_b200.ensemble import RayEnsemble

def time_it(fn, reps, flush, pre=None):
    ts = []
    for _ in range(reps):
        if pre is not None: pre()
        if flush is not None: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)

def main():
    sizes = [int(float(x)) for x in (sys.argv[1:] or ["0", "1e5", "1e6", "1e7"])]
    eng = Engine.get()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=eng.device)
    for shuffled in (False, True):
        for n in sizes:
            sc = scenarios.column_ensemble(max(n, 1), seed=1234, ngrid=1001, shuffled=shuffled)
            ens = RayEnsemble.from_scenario(sc)
            if n == 0: ens.n = 0
            p = ens.params(sc.dt); g = eng.grid_struct(ens.grid_devs); rays = ens._rays(); P = eng.ptr
            rr_out, mm_out, uo, vo = eng.empty(max(n,1)), eng.empty(max(n,1)), eng.empty(ens.G), eng.empty(ens.G)
            fa = lambda: check(lib.msgwam_column_pass_a(p, rays, ens.n, g, P(ens.uu), P(ens.vv), P(ens.work), eng.stream))
            fb = lambda: check(lib.msgwam_column_pass_b(p, rays, ens.n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out), eng.stream))
            ff = lambda: check(lib.msgwam_column_finish(p, g, P(ens.uu), P(ens.vv), P(ens.work), P(uo), P(vo), eng.stream))
            def step(): fa(); fb(); ff()
            def fused(): check(lib.msgwam_column_step(p, rays, ens.n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out), P(uo), P(vo), eng.stream))
            for _ in range(3): step()
            res = dict(n=n, shuffled=shuffled, pass_a_us=time_it(fa, 10, flush), pass_b_us=time_it(fb, 10, flush, pre=fa),   # pass A arms the chain counter
                       finish_us=time_it(ff, 10, flush), step_us=time_it(step, 10, flush), fused_us=time_it(fused, 10, flush), fused_us_noflush=time_it(fused, 10, None))
            if n: res["ray_steps_per_s"] = n / (res["fused_us"] * 1e-6)
            print(json.dumps(res), flush=True)
            if n == 0 and shuffled: pass
main()
