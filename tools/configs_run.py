"""Developer tool: BASELINE configs[2] (1e7 rays, N^2(z) profile, sheared U, G = 1000, one GPU) and configs[3] (1e8
rays over 8 GPUs, the same physics) through RayEnsemble.step; run configs[3] under torchrun with 8 ranks.
usage: python tools/configs_run.py [rays_per_gpu]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))
import numpy as np, torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from msgwam_b200 import scenarios
from msgwam_b200.ensemble import RayEnsemble

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else (10_000_000 if world == 1 else 12_500_000)
sc = scenarios.column_ensemble(n, seed=1234 + rank, ngrid=1001, sheared=True, amplitude=0.01)
prof = np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3))))          # SURVEY.md 8(d), configs[2]
sc.model = dict(sc.model, bvf=prof)
ens = RayEnsemble.from_scenario(sc)
del sc.state
ts = []
for k in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    a.record(); ens.step(120.0); b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
ens.check_errors()
ok = bool(torch.isfinite(ens.uu).all() and torch.isfinite(ens.field("rr")).all())
if rank == 0:
    print("N(z) profile + sheared wind, %d GPU(s) x %d rays: step times (ms, max over ranks) %s -> first step %.3e, step 8 %.3e ray-steps/s; finite: %s" % (
        world, n, [round(x, 3) for x in ts], n * world / (ts[0] * 1e-3), n * world / (ts[-1] * 1e-3), ok), flush=True)
if world > 1:
    dist.destroy_process_group()
