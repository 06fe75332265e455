import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import torch
from msgwam_b200 import scenarios
from msgwam_b200._engine import Engine
from msgwam_b200._cabi import check, lib
from msgwam_b200.ensemble import RayEnsemble
eng = Engine.get()
for n in (1000000,):
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001)
    ens = RayEnsemble.from_scenario(sc)
    p = ens.params(sc.dt); g = eng.grid_struct(ens.grid_devs); rays = ens._rays(); P = eng.ptr
    rr_out, mm_out, uo, vo = eng.empty(n), eng.empty(n), eng.empty(ens.G), eng.empty(ens.G)
    for it in range(3):
        print("---- fused step", it, flush=True)
        check(lib.msgwam_column_step(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out), P(uo), P(vo), eng.stream))
        torch.cuda.synchronize()
    import numpy as np, statistics
    base = int(lib.msgwam_column_work_doubles(ens.G)) - 2 * 160 * 16 - 16
    tr = ens.work[base:base + 2 * 160 * 16].cpu().numpy().reshape(2, 160, 16)[:, :148]
    gm = ens.work[base + 2 * 160 * 16:].cpu().numpy()
    print('chain phases (cycles): load %d tables0 %d chain0 %d tables1 %d chain1 %d tables2 %d' % tuple(gm[k + 1] - gm[k] for k in range(6)))
    names = ["prologue", "sweep", "winflush", "histflush", "ticket", "tail", "end"]
    for ps in (0, 1):
        t = tr[ps]
        end = t[:, 1]; e0 = end.min()
        print("pass", "AB"[ps], "end spread us %.1f" % ((end.max() - e0) / 1e3))
        for k, nm in enumerate(names):
            col = t[:, 3 + k]
            print("   %-10s cycles: min %8.0f med %8.0f max %8.0f  (max = %.1f us)" % (nm, col.min(), np.median(col), col.max(), col.max() / 1965.))
        worst = np.argsort(-end)[:4]
        print("   last CTAs:", [(int(c), int(t[c, 0]), round((end[c] - e0) / 1e3, 1), [int(x) for x in t[c, 3:10]]) for c in worst])
