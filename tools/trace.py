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
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=eng.device)
    for it in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.msgwam_column_step(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out), P(uo), P(vo), eng.stream))
        e1.record()
        torch.cuda.synchronize()
        print("---- fused step", it, "event time %.1f us" % (e0.elapsed_time(e1) * 1e3), flush=True)
    import numpy as np, statistics
    base = int(lib.msgwam_column_work_doubles(ens.G)) - 2 * 160 * 16 - 16
    tr = ens.work[base:base + 2 * 160 * 16].cpu().numpy().reshape(2, 160, 16)[:, :148]
    gm = ens.work[base + 2 * 160 * 16:].cpu().numpy()
    pass
    names = ["prologue", "sweep", "winflush", "histflush", "ticket", "tail", "end"]
    for ps in (0, 1):
        t = tr[ps]
        end = t[:, 1]; e0 = end.min(); start = t[:, 15]
        print("   kernel span: first CTA start -> last CTA end = %.1f us; CTA start spread %.1f us; passA end -> passB start gap: see below" % ((end.max() - start.min()) / 1e3, (start.max() - start.min()) / 1e3))
        if ps == 1: print("   gap between pass A last end and pass B first start: %.1f us" % ((start.min() - tr[0][:, 1].max()) / 1e3))
        print("pass", "AB"[ps], "end spread us %.1f" % ((end.max() - e0) / 1e3))
        for k, nm in enumerate(names):
            col = t[:, 3 + k]
            print("   %-10s cycles: min %8.0f med %8.0f max %8.0f  (max = %.1f us)" % (nm, col.min(), np.median(col), col.max(), col.max() / 1965.))
        sw = t[:, 4]
        print("   sweep cycles by CTA (every 8th):", [int(x) for x in sw[::8]])
        print("   sweep cycles by SM id (sorted by sm):", [int(x) for x in sw[np.argsort(t[:, 0])][::8]])
        worst = np.argsort(-end)[:4]
        print("   last CTAs:", [(int(c), int(t[c, 0]), round((end[c] - e0) / 1e3, 1), [int(x) for x in t[c, 3:10]]) for c in worst])
