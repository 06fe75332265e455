"""Developer tool (round 2): how much does the loss of spatial order cost the deposit, and how long does a global
re-ordering by (group-velocity bucket, cell) hold?  The re-ordering is prototyped with torch.sort / index_select here
(measurement only; the product's re-ordering kernels live in csrc/reorder.cu).

usage: python tools/disperse_probe.py <mode: nz|const> <rays> [steps_disperse] [steps_after]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble

mode = sys.argv[1] if len(sys.argv) > 1 else "nz"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 3_000_000
k_disp = int(sys.argv[3]) if len(sys.argv) > 3 else 40
k_after = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dt = 120.0

if mode == "nz":
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001, sheared=True, amplitude=0.01)
    prof = np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3))))
    sc.model = dict(sc.model, bvf=prof)
else:
    sc = scenarios.column_ensemble(n, seed=1234, ngrid=1001)
    prof = None
ens = RayEnsemble.from_scenario(sc)
grids0, dz = float(sc.grids[0]), float(sc.grids[1] - sc.grids[0])
G = len(sc.grids)
bvf_dev = torch.as_tensor(prof if prof is not None else np.full(G, 0.01), device="cuda")
del sc.state
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def timed_steps(k):
    ts = []
    for _ in range(k):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(dt); b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    return [round(a.elapsed_time(b), 3) for a, b in ts]


def cg_cells_per_step():
    rr, mm, kk, ll, ff = (ens.field(x) for x in ("rr", "mm", "kk", "ll", "ff"))
    j = ((rr - grids0) / dz).floor().clamp_(0, G - 1).long()
    n2 = bvf_dev[j] ** 2
    kh2 = kk * kk + ll * ll
    vk = kh2 + mm * mm
    om2 = (n2 * kh2 + ff * ff * mm * mm) / vk
    cg = -mm * (om2 - ff * ff) / om2.sqrt() / vk
    return cg * dt / dz, j


def reorder(width):
    """sort by (cg bucket of `width` cells per step, cell); width None: by cell only.  Returns ms spent."""
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, j = cg_cells_per_step()
    hd = ens.field("drr") * .5
    cell = ((ens.field("rr") - hd) / dz).floor().clamp_(0, G - 1).long()
    if width is None:
        key = cell
    else:
        b = (c / width).floor().clamp_(-2000, 2000).long() + 2000
        key = b * 1024 + cell
    perm = torch.argsort(key)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    nn = ens.n
    ens._slab[:, :nn] = ens._slab[:, :nn].index_select(1, perm)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, int(key.unique().numel())


print("mode %s, %d rays" % (mode, n), flush=True)
t = timed_steps(k_disp)
print("from the ordered initial ensemble, in place: ms per step", t, flush=True)
c, _ = cg_cells_per_step()
q = torch.quantile(c[:: max(1, n // 1_000_000)].abs(), torch.tensor([.01, .1, .5, .9, .99, 1.0], device="cuda", dtype=torch.float64))
print("|cg| dt/dz (cells per step) quantiles 1/10/50/90/99/100 %:", [round(float(x), 4) for x in q], flush=True)
for width in (None, 0.4, 0.1, 0.025):
    ks, kp, nk = reorder(width)
    t = timed_steps(k_after)
    print("re-ordered by %s: keys %d, torch sort %.1f ms, permute 14 fields %.1f ms; next steps ms: %s" % (
        "cell" if width is None else "(cg bucket %.3f cells/step, cell)" % width, nk, ks, kp, t), flush=True)
ens.check_errors()
print("finite:", bool(torch.isfinite(ens.uu).all() and torch.isfinite(ens.field("rr")).all()))
