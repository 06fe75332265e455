"""Developer tool: where the end-to-end (host buffers) step time goes -- raw PCIe copies vs the RK3 call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch

def bw(nbytes, direction, reps=10, pinned=True):
    h = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=pinned)
    d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    for _ in range(2):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

for nb in (8 << 20, 64 << 20, 256 << 20):
    print("pinned   %4d MiB  H2D %.1f GB/s   D2H %.1f GB/s" % (nb >> 20, bw(nb, "h2d"), bw(nb, "d2h")))
print("pageable   64 MiB  H2D %.1f GB/s   D2H %.1f GB/s" % (bw(64 << 20, "h2d", pinned=False), bw(64 << 20, "d2h", pinned=False)))

import msgwam_b200.libprop as lprop
from msgwam_b200 import scenarios
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t
for n in (1000, 100000):          # the fixed cost of a host call
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001)
    sc.install(lprop)
    keep = [pinned(np.ascontiguousarray(a)) for a in list(sc.state) + [sc.uu, sc.vv, sc.dkk, sc.dll, sc.rr_mm_area]]
    var = np.empty(11, dtype=object)
    for i in range(11): var[i] = keep[i].numpy()
    lprop.set_statics(dkk=keep[11].numpy(), dll=keep[12].numpy(), rr_mm_area=keep[13].numpy())
    for _ in range(3): lprop.RK3(sc.dt, var)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); out = lprop.RK3(sc.dt, var); ts.append(time.perf_counter() - t0)
    print("RK3 host call at n = %d: median %.3f ms  min %.3f ms" % (n, np.median(ts) * 1e3, min(ts) * 1e3))
n = 1000000
sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001)
sc.install(lprop)
keep = [pinned(np.ascontiguousarray(a)) for a in list(sc.state) + [sc.uu, sc.vv, sc.dkk, sc.dll, sc.rr_mm_area]]
var = np.empty(11, dtype=object)
for i in range(11): var[i] = keep[i].numpy()
lprop.set_statics(dkk=keep[11].numpy(), dll=keep[12].numpy(), rr_mm_area=keep[13].numpy())
for _ in range(3): lprop.RK3(sc.dt, var)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); out = lprop.RK3(sc.dt, var); ts.append(time.perf_counter() - t0)
print("RK3 host call: median %.3f ms  min %.3f ms" % (np.median(ts) * 1e3, min(ts) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): lprop.RK3(sc.dt, var)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
