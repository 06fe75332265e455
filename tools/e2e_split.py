"""Developer tool: host time of lprop.RK3 (numpy in / numpy out) split into Python before the C call, the C call, Python after."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
import msgwam_b200.libprop as lprop
from msgwam_b200 import scenarios, _cabi
marks = {}
class Wrap:
    def __init__(self, f): self.f = f
    def __call__(self, *a):
        marks["in"] = time.perf_counter(); r = self.f(*a); marks["out"] = time.perf_counter(); return r
class LibProxy:
    def __init__(self, lib): self._lib = lib; self._w = Wrap(lib.msgwam_rk3_column_host)
    def __getattr__(self, k): return self._w if k == "msgwam_rk3_column_host" else getattr(self._lib, k)
lprop.lib = LibProxy(lprop.lib)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t
for n in (1000, 1000000):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001)
    sc.install(lprop)
    keep = [pinned(np.ascontiguousarray(a)) for a in list(sc.state) + [sc.uu, sc.vv, sc.dkk, sc.dll, sc.rr_mm_area]]
    var = np.empty(11, dtype=object)
    for i in range(11): var[i] = keep[i].numpy()
    lprop.set_statics(dkk=keep[11].numpy(), dll=keep[12].numpy(), rr_mm_area=keep[13].numpy())
    for _ in range(3): lprop.RK3(sc.dt, var)
    pre, call, post = [], [], []
    for _ in range(20):
        t0 = time.perf_counter(); out = lprop.RK3(sc.dt, var); t1 = time.perf_counter()
        pre.append(marks["in"] - t0); call.append(marks["out"] - marks["in"]); post.append(t1 - marks["out"])
    print("n = %8d: Python before the C call %.0f us, C call %.0f us, Python after %.0f us (medians)" % (
        n, np.median(pre) * 1e6, np.median(call) * 1e6, np.median(post) * 1e6), flush=True)
