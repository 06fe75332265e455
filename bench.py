#!/usr/bin/env python
"""bench.py -- ray-steps/s of the RK3 + flux-deposition hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R]

A "step" is one lprop.RK3 step (3 RK stages, 3 flux depositions, mean-flow update) over the
synthetic column ensemble of BASELINE.json configs[1]: R = 1e6 ray volumes per GPU, constant N,
zero mean wind.  One JSON line is printed by rank 0.

  value     device-resident throughput: inputs already in HBM, per-step CUDA-event timing on the launch
            stream, L2 flushed (256 MiB write) between timed steps, max over ranks.
  e2e       the same step through the reference-facing call lprop.RK3(dt, var) with HOST (pinned) numpy
            buffers: H2D of the step's inputs + kernels + D2H of rr, mm, uu, vv inside the timed region.
  roofline  dominant kernel (pass B: 3 RK stages + 1 deposit + store) against the measured HBM peak.
  cpu_baseline  the oracle port (C restatement of the reference, 1 thread) on this box's host cores.

--impl reference times the reference's CPU algorithm (the oracle port; the reference itself is Python
and cannot travel to the GPU box) with all host threads on the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "python-msgwam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "ray-steps/sec (RK step incl. flux deposition)"
UNIT = "ray-steps/s"
A_STEP = 96.0      # algorithmic B/ray-step, SURVEY.md 8(d): 10 fp64 fields read once + rr, mm written once
A_PASS_A = 72.0    # pass A: 9 fields read (dens, ff, rr, drr, kk, ll, mm, dmm, dkk*dll); its 24 B/ray hand-over to pass B
                   # (stage-1 increments, cg_rr(r1)) is implementation traffic, not algorithmic
A_PASS_B = 88.0    # pass B: the same 9 fields read + rr, mm written (+ the 24 B/ray hand-over read back)
# dram__bytes_read.sum + dram__bytes_write.sum per ray from the committed `ncu --set full` capture at 1e6 rays
# (profiles/r01d_column_pass_ncu_full_summary.json): pass A 72.1 + 5.5 MB, pass B 96.2 + 5.5 MB
NCU_TRAFFIC_PER_RAY = {"A": 79.9, "B": 102.8}
NCU_TRAFFIC_SOURCE = "ncu --set full at 1e6 rays, profiles/r01d_column_pass_ncu_full_summary.json"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: an NVML thread (a query costs microseconds and
    does not stall PCIe the way a polling nvidia-smi process does); nvidia-smi -lms is the fallback."""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None, period_s=0.025):
        self.proc = None
        self.thread = None
        self.sm, self.mx, self.reasons = [], [], set()
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self._stop = threading.Event()

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = int(reasons_fn(h))
                        for nm, bit in self.BITS.items():
                            if mask & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    self._stop.wait(period_s)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def make_scenario(n, rank):
    from msgwam_b200 import scenarios
    return scenarios.column_ensemble(n, seed=1234 + rank, ngrid=1001)


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    n = args.rays if args.steps <= 30 else min(args.rays, 250_000)     # bounded sample per step
    sc = make_scenario(n, 0)
    best = None
    tmax = oracle.max_threads()
    for nthreads in sorted({1, tmax}, reverse=True):
        orc = oracle.Oracle(sc.oracle_cfg(), nthreads=nthreads)
        for _ in range(max(args.warmup, 1) if nthreads == tmax else 1):
            orc.RK3(sc.dt, sc.var())
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.RK3(sc.dt, sc.var())
        dt = time.perf_counter() - t0
        rate = n * args.steps / dt
        if best is None or rate > best[0]:
            best = (rate, nthreads, dt)
    rate, cores, dt = best
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("configs[1]: 1e6 ray volumes per GPU, 1-D column, constant N, zero mean wind, G=1000" if args.rays == 1_000_000 else
                                "configs[1] ensemble at %d ray volumes per GPU (--rays), 1-D column, constant N, zero mean wind, G=1000" % args.rays),
                   "rays_per_gpu": args.rays, "grid_levels": 1000, "dt_s": sc.dt,
                   "sample": "each timed step advances a bounded sample of %d rays of that ensemble on the host cores" % n},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "oracle/msgwam_oracle.c (C restatement of lib/libprop.py RK3+rhs_default+wave_projection; the "
                                   "Python reference cannot run on the GPU box and does ~2.3e4 ray-steps/s, BASELINE.md) on %d rays per step, "
                                   "%d steps, best of {1,%d} OpenMP threads" % (n, args.steps, tmax)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import msgwam_b200.libprop as lprop
    from msgwam_b200._engine import Engine
    from msgwam_b200._cabi import check, lib
    from msgwam_b200.ensemble import RayEnsemble

    n = args.rays
    sc = make_scenario(n, rank)
    sc.install(lprop)
    eng = Engine.get()
    dev = eng.device
    ens = RayEnsemble.from_scenario(sc)
    p = ens.params(sc.dt)
    g = eng.grid_struct(ens.grid_devs)
    rays = ens._rays()
    rr_out, mm_out = eng.empty(n), eng.empty(n)
    uu_out, vv_out = eng.empty(ens.G), eng.empty(ens.G)
    nc = ens.G - 1
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    P = eng.ptr

    from msgwam_b200.distributed import PeerExchange
    exchange = PeerExchange.get(ens.G) if world > 1 else None

    def reduce_(t):
        if world > 1 and exchange is None:
            dist.all_reduce(t)

    def pass_a():
        check(lib.msgwam_column_pass_a(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), eng.stream), "pass_a")

    def pass_b():
        if exchange is not None:      # chain kernel all-reduces D0|D1 over peer memory, then the sweep
            check(lib.msgwam_column_pass_b_p2p(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out),
                                               exchange.next(), eng.stream), "pass_b_p2p")
        else:
            check(lib.msgwam_column_pass_b(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out), eng.stream), "pass_b")

    def finish():
        if exchange is not None:
            check(lib.msgwam_column_finish_p2p(p, g, P(ens.uu), P(ens.vv), P(ens.work), P(uu_out), P(vv_out), exchange.next(),
                                               eng.stream), "finish_p2p")
        else:
            check(lib.msgwam_column_finish(p, g, P(ens.uu), P(ens.vv), P(ens.work), P(uu_out), P(vv_out), eng.stream), "finish")

    def step():          # out of place: every timed step does identical work on the same input state
        if world == 1:   # the call libprop.RK3 / RayEnsemble.step make on one GPU: two launches, finish fused as pass B's tail
            check(lib.msgwam_column_step(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out),
                                         P(uu_out), P(vv_out), eng.stream), "column_step")
        elif exchange is not None:   # several GPUs: still two launches; the all-reduces run in the sweeps' tails over NVLink
            check(lib.msgwam_column_step_p2p(p, rays, n, g, P(ens.uu), P(ens.vv), P(ens.work), P(rr_out), P(mm_out),
                                             P(uu_out), P(vv_out), exchange.next(2), eng.stream), "column_step_p2p")
        else:                        # NCCL fallback: all-reduces between the kernels
            pass_a(); reduce_(ens.work[:4 * nc]); pass_b(); reduce_(ens.work[4 * nc:6 * nc]); finish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_(); step()
    barrier()
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid) if rank == 0 else None

    # ---- value: device-resident steps, per-step events, L2 flushed between steps ------------------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record(); step(); b.record()
    barrier()
    wall = time.perf_counter() - wall0
    t_steps = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    tt = torch.tensor([t_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_steps = float(tt.item())
    launches = (2 if (world == 1 or exchange is not None) else 3) * args.steps

    # ---- per-kernel timing for the roofline (single rank's kernels; no collectives inside) --------
    ka, kb, kf = [], [], []
    flush.zero_(); pass_a(); reduce_(ens.work[:4 * nc]); pass_b(); reduce_(ens.work[4 * nc:6 * nc]); finish()   # first launch of the split-form kernels
    torch.cuda.synchronize()
    for _ in range(max(3, min(args.steps, 20))):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); pass_a(); e[1].record(); reduce_(ens.work[:4 * nc]); pass_b_start = torch.cuda.Event(enable_timing=True)
        pass_b_start.record(); pass_b(); e[2].record(); reduce_(ens.work[4 * nc:6 * nc]); fin0 = torch.cuda.Event(enable_timing=True)
        fin0.record(); finish(); e[3].record()
        torch.cuda.synchronize()
        ka.append(e[0].elapsed_time(e[1])); kb.append(pass_b_start.elapsed_time(e[2])); kf.append(fin0.elapsed_time(e[3]))
    t_a, t_b, t_f = (statistics.mean(x) * 1e-3 for x in (ka, kb, kf))     # average launch duration, as the contract asks

    # ---- e2e: the reference-facing call with host buffers ------------------------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep = [pinned(np.ascontiguousarray(a)) for a in list(sc.state) + [sc.uu, sc.vv, sc.dkk, sc.dll, sc.rr_mm_area]]
    var = np.empty(11, dtype=object)
    for i in range(11):
        var[i] = keep[i].numpy()
    lprop.set_statics(dkk=keep[11].numpy(), dll=keep[12].numpy(), rr_mm_area=keep[13].numpy())
    # per step: dens, phi, rr, drr, kk, ll, mm, dmm + grid fields (the per-run statics dkk, dll are uploaded by the
    # first call only, as long as the caller keeps passing the same arrays)
    h2d = 8 * n * 8 + (ens.G + 1 + 6 * ens.G) * 8
    d2h = 2 * n * 8 + 2 * ens.G * 8                          # rr, mm, uu, vv
    if world == 1:
        def e2e_step():
            return lprop.RK3(sc.dt, var)
    else:
        from msgwam_b200.distributed import rk3_host_sharded
        def e2e_step():
            return rk3_host_sharded(lprop, sc.dt, var)
    k_e2e = max(3, min(args.steps, 10))
    outs = [e2e_step() for _ in range(3)]                    # warm-up; results kept alive so that the pinned result
    del outs                                                 # buffers of two steps are in the host allocator's cache
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        out = e2e_step()
        _ = float(out[9][0])                                 # the step's result is read on the host
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    # ---- the same step in the other regimes of the deposit (N = 1; reported beside the headline, not part of it) ----
    regimes = None
    if world == 1 and not args.no_regimes:
        def time_steps(e2, k, in_place):
            pp, gg, rr_ = e2.params(sc.dt), eng.grid_struct(e2.grid_devs), e2._rays()
            r_o, m_o = (e2.field("rr"), e2.field("mm")) if in_place else (rr_out, mm_out)
            ts = []
            for _ in range(k):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                check(lib.msgwam_column_step(pp, rr_, n, gg, P(e2.uu), P(e2.vv), P(e2.work), P(r_o), P(m_o),
                                             P(e2._uu2), P(e2._vv2), eng.stream), "column_step")
                b.record()
                if in_place:
                    e2.uu, e2._uu2 = e2._uu2, e2.uu
                    e2.vv, e2._vv2 = e2._vv2, e2.vv
                ts.append((a, b))
            torch.cuda.synchronize()
            return statistics.median(a.elapsed_time(b) for a, b in ts) * 1e-3
        from msgwam_b200 import scenarios as _scn
        k_r = max(5, min(args.steps, 20))
        ens_d = RayEnsemble.from_scenario(sc)
        time_steps(ens_d, 30, True)                              # the ensemble disperses within ~10 in-place steps
        t_disp = time_steps(ens_d, k_r, True)
        sc_s = _scn.column_ensemble(n, seed=1234 + rank, ngrid=1001, shuffled=True)
        ens_s = RayEnsemble.from_scenario(sc_s)
        time_steps(ens_s, 3, False)
        t_shuf = time_steps(ens_s, k_r, False)
        regimes = {"note": "median GPU time per step, L2 flushed; the headline times the height-ordered ensemble of SURVEY 8(d)",
                   "dispersed_after_30_in_place_steps": {"ms_per_step": t_disp * 1e3, "value": n / t_disp, "unit": UNIT},
                   "shuffled_ray_order": {"ms_per_step": t_shuf * 1e3, "value": n / t_shuf, "unit": UNIT}}
        del ens_d, ens_s
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        orc = oracle.Oracle(sc.oracle_cfg(), nthreads=1)
        orc.RK3(sc.dt, sc.var())
        t0 = time.perf_counter(); reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 40):
            orc.RK3(sc.dt, sc.var()); reps += 1
        dtc = time.perf_counter() - t0
        cpu = {"value": n * reps / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle port (C, 1 thread, -O2, no FMA) on the full %d-ray ensemble, %d RK3 steps in %.1f s; "
                         "the unmodified Python reference does ~2.3e4 ray-steps/s on one core (BASELINE.md)" % (n, reps, dtc),
               "host_cpus": os.cpu_count()}

    peak, peak_src = measured_peaks()
    total_rays = n * world
    value = total_rays * args.steps / t_steps
    kernels = {
        "A": {"name": "column_pass<0> (pass A: deposit r0, stage 1, deposit r1, hand-over)", "ms": t_a * 1e3, "alg": A_PASS_A},
        "B": {"name": "column_pass<1> (pass B: mean-flow chain, stages 2-3, deposit r2, store)", "ms": t_b * 1e3, "alg": A_PASS_B},
    }
    dom = "A" if t_a >= t_b else "B"           # the roofline is reported for whichever sweep takes longer
    oth = "B" if dom == "A" else "A"
    ach = kernels[dom]["alg"] * n / (kernels[dom]["ms"] * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_steps / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("configs[1]: 1e6 ray volumes per GPU, 1-D column, constant N, zero mean wind, G=1000" if n == 1_000_000 else
                                "configs[1] ensemble at %d ray volumes per GPU (--rays), 1-D column, constant N, zero mean wind, G=1000" % n),
                   "rays_per_gpu": n, "grid_levels": ens.G, "dt_s": sc.dt, "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": "rays sharded, %d rank(s); all-reduce of the deposited flux twice per step (%s)" % (
                       world, "none needed" if world == 1 else ("one-shot pushes over NVLink peer memory, fused into the tails of the two sweeps" if exchange is not None else "NCCL")),
                   "mode": "M1 coupled (reference RK3 semantics: mean flow inside the RK state), 2 ray sweeps per step"},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": NCU_TRAFFIC_PER_RAY[dom] * n, "traffic_source": NCU_TRAFFIC_SOURCE,
                     "kernel": kernels[dom]["name"], "algorithmic_bytes_per_ray": kernels[dom]["alg"],
                     "kernel_ms": kernels[dom]["ms"], "peak_source": peak_src,
                     "note": "the sweeps are bound by instruction issue / fp64 latency, not by HBM (ncu: issue slots 59 % busy, "
                             "fp64 pipe 37 %, DRAM 23 %): see profiles/r01_summary.md",
                     "other_kernels": {kernels[oth]["name"]: {"ms": kernels[oth]["ms"],
                                                              "achieved_gbs": kernels[oth]["alg"] * n / (kernels[oth]["ms"] * 1e-3) / 1e9,
                                                              "algorithmic_bytes_per_ray": kernels[oth]["alg"],
                                                              "traffic": NCU_TRAFFIC_PER_RAY[oth] * n},
                                       "column_finish (a separate kernel only in the split form)": {"ms": t_f * 1e3}},
                     "step": {"algorithmic_bytes_per_ray_step": A_STEP,
                              "achieved_gbs": A_STEP * total_rays * args.steps / t_steps / 1e9 / world,
                              "frac": A_STEP * total_rays * args.steps / t_steps / 1e9 / world / peak}},
        "e2e": {"value": total_rays * k_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": k_e2e, "api": "msgwam_b200.libprop.RK3(dt, var) with pinned numpy buffers"},
        "gpu_launches": launches,
        "clocks": clocks,
        "wall_s_timed_region_incl_flush": wall,
    }
    if regimes is not None:
        line["regimes"] = regimes
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1_000_000, help="ray volumes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the dispersed / shuffled ensemble timings (N = 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
