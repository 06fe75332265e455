"""Multi-GPU parity check (run under torchrun, one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_check.py

Every rank owns a contiguous slice of one synthetic ensemble, steps it with RayEnsemble (all-reduce of the
deposited flux between the sweeps), and rank 0 compares the gathered result with the single-process oracle.
Also exercises the host-buffer sharded call (distributed.rk3_host_sharded).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-msgwam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import msgwam_b200.libprop as lprop
    from msgwam_b200 import scenarios
    from msgwam_b200.distributed import rk3_host_sharded, shard_range
    from msgwam_b200.ensemble import RayEnsemble
    import oracle
    from helpers import FIELDS, field_rel

    nsteps = 3
    for shuffled in (False, True):
        sc = scenarios.column_ensemble(200003, seed=77, ngrid=801, sheared=True, shuffled=shuffled, amplitude=0.3)
        b, e = shard_range(sc.n, rank, world)
        ens = RayEnsemble([a[b:e] for a in sc.state], sc.dkk[b:e], sc.dll[b:e], sc.rr_mm_area[b:e], sc.uu, sc.vv, sc.grid,
                          sc.grids, sc.rhobar, sc.pressure_gradient, bvf=sc.model["bvf"], phi0=sc.model["phi0"])
        ens.step(sc.dt, nsteps)
        mine = ens.to_var()
        # host-buffer sharded call, one step
        sc.install(lprop)
        lprop.set_statics(dkk=sc.dkk[b:e].copy(), dll=sc.dll[b:e].copy(), rr_mm_area=sc.rr_mm_area[b:e].copy())
        loc = np.empty(11, dtype=object)
        for i in range(9):
            loc[i] = np.ascontiguousarray(sc.state[i][b:e])
        loc[9], loc[10] = sc.uu, sc.vv
        host1 = rk3_host_sharded(lprop, sc.dt, loc)
        gathered = [None] * world
        dist.all_gather_object(gathered, ([mine[i] for i in range(11)], [host1[i] for i in range(11)]))
        if rank == 0:
            orc = oracle.Oracle(sc.oracle_cfg())
            want = sc.var()
            want1 = None
            for s in range(nsteps):
                want = orc.RK3(sc.dt, want)
                if s == 0:
                    want1 = want
            for tag, idx, ref in (("ensemble %d steps" % nsteps, 0, want), ("host sharded 1 step", 1, want1)):
                worst = 0.0
                for i, nm in enumerate(FIELDS):
                    if nm in ("uu", "vv"):
                        for r in range(world):
                            err = field_rel(gathered[r][idx][i], ref[i])
                            assert err <= 1e-12, (tag, nm, r, err)
                    else:
                        got = np.concatenate([gathered[r][idx][i] for r in range(world)])
                        scale = np.maximum(np.abs(ref[i]), np.abs(ref[i] - sc.var()[i]))
                        diff = np.abs(got - ref[i])
                        err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale))))
                        worst = max(worst, err)
                        assert err <= 1e-12, (tag, nm, err)
                # every rank must hold the identical mean flow
                for r in range(1, world):
                    assert np.array_equal(gathered[r][idx][9], gathered[0][idx][9]), (tag, "uu differs between ranks")
                print("multi-GPU parity ok: world=%d shuffled=%s %s worst per-ray rel err %.2e" % (world, shuffled, tag, worst), flush=True)
    # N(z) extension, sharded: the fused profile step with the all-reduces in its sweeps' tails (or, without peer
    # memory, the stage-by-stage path with NCCL)
    sc = scenarios.column_ensemble(120011, seed=78, ngrid=601, sheared=True, amplitude=0.3)
    prof = np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((sc.grids - 15e3) / 3e3))))
    sc.model = dict(sc.model, bvf=prof)
    b, e = shard_range(sc.n, rank, world)
    ens = RayEnsemble([a[b:e] for a in sc.state], sc.dkk[b:e], sc.dll[b:e], sc.rr_mm_area[b:e], sc.uu, sc.vv, sc.grid,
                      sc.grids, sc.rhobar, sc.pressure_gradient, bvf=prof, phi0=sc.model["phi0"])
    ens.step(sc.dt, nsteps)
    ens.check_errors()
    mine = ens.to_var()
    # the host-buffer sharded call with the profile (what bench.py's e2e leg runs at N > 1), one step
    sc.install(lprop)
    lprop.set_statics(dkk=sc.dkk[b:e].copy(), dll=sc.dll[b:e].copy(), rr_mm_area=sc.rr_mm_area[b:e].copy())
    loc = np.empty(11, dtype=object)
    for i in range(9):
        loc[i] = np.ascontiguousarray(sc.state[i][b:e])
    loc[9], loc[10] = sc.uu, sc.vv
    host1 = rk3_host_sharded(lprop, sc.dt, loc)
    gathered = [None] * world
    dist.all_gather_object(gathered, [mine[i] for i in range(11)])
    gh = [None] * world
    dist.all_gather_object(gh, [np.asarray(host1[i]) for i in range(11)])
    if rank == 0:
        orc = oracle.Oracle(sc.oracle_cfg())
        want = sc.var()
        for s in range(nsteps):
            want = orc.RK3(sc.dt, want)
            if s == 0:
                for i, nm in enumerate(FIELDS):
                    if nm in ("uu", "vv"):
                        for r in range(world):
                            assert field_rel(gh[r][i], want[i]) <= 1e-11, ("profile host sharded", nm, r)
                    else:
                        got1 = np.concatenate([gh[r][i] for r in range(world)])
                        sc1 = np.maximum(np.abs(want[i]), np.abs(want[i] - sc.var()[i]))
                        d1 = np.abs(got1 - want[i])
                        assert float(np.max(np.where(d1 == 0, 0.0, d1 / np.where(sc1 == 0, 1.0, sc1)))) <= 1e-12, ("profile host sharded", nm)
                print("multi-GPU parity ok: world=%d N(z) profile host sharded 1 step" % world, flush=True)
        worst = 0.0
        for i, nm in enumerate(FIELDS):
            if nm in ("uu", "vv"):
                for r in range(world):
                    assert field_rel(gathered[r][i], want[i]) <= 1e-11, ("profile", nm, r)
            else:
                got = np.concatenate([gathered[r][i] for r in range(world)])
                scale = np.maximum(np.abs(want[i]), np.abs(want[i] - sc.var()[i]))
                diff = np.abs(got - want[i])
                err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale))))
                worst = max(worst, err)
                assert err <= 1e-12, ("profile", nm, err)
        for r in range(1, world):
            assert np.array_equal(gathered[r][9], gathered[0][9]), "uu differs between ranks (profile)"
        print("multi-GPU parity ok: world=%d N(z) profile ensemble %d steps worst per-ray rel err %.2e" % (world, nsteps, worst), flush=True)
    # ---- skewed deletion, re-balancing, then the driver loop (fused clamp) and the frozen-background mode, sharded ----
    sc = scenarios.column_ensemble(160_009, seed=79, ngrid=801, sheared=True, amplitude=1.0)
    ids = np.arange(sc.n, dtype=np.float64)
    sc.state[1] = ids.copy()                      # lam is inert in column mode: it carries the ray's identity
    b, e = shard_range(sc.n, rank, world)
    ens = RayEnsemble([a[b:e] for a in sc.state], sc.dkk[b:e], sc.dll[b:e], sc.rr_mm_area[b:e], sc.uu, sc.vv, sc.grid,
                      sc.grids, sc.rhobar, sc.pressure_gradient, bvf=sc.model["bvf"], phi0=sc.model["phi0"])
    # delete by |m| >= m_crit with a threshold that removes most rays of the low ranks (mm grows with the ray index? no:
    # random) -- so skew it by hand: rank 0 deletes with a tight threshold, the others with a loose one
    m_crit = float(np.quantile(np.abs(sc.state[7]), 0.15 if rank == 0 else 0.9))
    kept = ens.compact(sc.dt, m_crit)
    counts = [None] * world
    dist.all_gather_object(counts, kept)
    new_n = ens.rebalance()
    after = [None] * world
    dist.all_gather_object(after, new_n)
    assert sum(after) == sum(counts) and max(after) - min(after) <= 1, (counts, after)
    ens.advance(sc.dt, 2, saturate=True)          # RK3 + fused post-step clamp, twice
    ens.step_frozen(sc.dt, 1)                     # and one frozen-background step on top
    mine = ens.to_var()
    gathered = [None] * world
    dist.all_gather_object(gathered, [mine[i] for i in range(11)])
    if rank == 0:
        got = [np.concatenate([gathered[r][i] for r in range(world)]) for i in range(9)]
        order = np.argsort(got[1])
        sel = got[1][order].astype(np.int64)      # surviving ray ids
        assert len(np.unique(sel)) == len(sel) == sum(counts)
        cfg = sc.oracle_cfg()
        cfg.update(dkk=sc.dkk[sel], dll=sc.dll[sel], rr_mm_area=sc.rr_mm_area[sel])
        orc = oracle.Oracle(cfg, nthreads=max(1, len(os.sched_getaffinity(0))))
        var = np.empty(11, dtype=object)
        for i in range(9):
            var[i] = sc.state[i][sel]
        var[9], var[10] = sc.uu, sc.vv
        start = [np.array(a) for a in var[:9]]
        for _ in range(2):
            out = orc.RK3(sc.dt, var)
            out[0] = orc.saturation(sc.dt, out[0], var[3], (out[3] - var[3]) / 1, var[4], (out[4] - var[4]) / sc.dt, out[5], out[6],
                                    var[7], (out[7] - var[7]) / sc.dt, direct=True)
            var = out
        var = orc.RK3_frozen(sc.dt, var)
        worst = 0.0
        for i, nm in enumerate(FIELDS):
            if nm in ("uu", "vv"):
                for r in range(world):
                    assert field_rel(gathered[r][i], var[i]) <= 1e-10, ("rebalance", nm, r, field_rel(gathered[r][i], var[i]))
            else:
                g = got[i][order]
                scale = np.maximum(np.abs(var[i]), np.abs(var[i] - start[i]))
                diff = np.abs(g - var[i])
                err = float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale))))
                worst = max(worst, err)
                assert err <= 1e-10, ("rebalance", nm, err)
        for r in range(1, world):
            assert np.array_equal(gathered[r][9], gathered[0][9]), "uu differs between ranks (rebalance)"
        print("multi-GPU parity ok: world=%d deletion %s -> re-balanced %s, 2 x advance (fused clamp) + 1 frozen step, worst per-ray rel err %.2e" % (
            world, counts, after, worst), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
