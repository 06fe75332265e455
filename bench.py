#!/usr/bin/env python
"""bench.py -- ray-steps/s of the RK3 + flux-deposition hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--workload c2|c1]

A "step" is one lprop.RK3 step (3 RK stages, 3 flux depositions, mean-flow update) over a synthetic column
ensemble.  Default workload: BASELINE.json configs[2] physics -- N^2(z) profile, sheared
U(z), G = 1000 -- at 1.25e7 ray volumes per GPU (8 GPUs = the 1e8 rays of configs[3]), advanced IN PLACE; the timed
steps follow >= 30 in-place warm-up steps, so the ensemble is in the dispersed steady state every long run lives in.
One JSON line is printed by rank 0.

  value     device-resident throughput: state in HBM, K in-place steps timed with CUDA events on the launch stream
            between barriers, max over ranks (the per-GPU state, 1.4 GB, is far larger than L2).
  e2e       the same step through the reference-facing call lprop.RK3(dt, var) with HOST numpy buffers
            (page-locked, statics frozen): H2D of the step's inputs + kernels + D2H of the changed slots inside
            the timed region; `e2e.pageable` is the same call with ordinary pageable arrays and default statics
            semantics -- what an unmodified driver script gets.
  parity    computed in this process at the same N ranks before anything is timed: 2e5 rays (sharded), 3 steps, the
            CUDA path against the CPU oracle on rank 0; a failure exits non-zero.
  roofline  dominant kernel against the measured HBM peak (algorithmic bytes per launch / live CUDA-event duration),
            the whole step beside it, and the fp64-pipe fraction (the sweeps are fp64-issue bound, not HBM bound).
  configs   (N = 1) the other BASELINE configurations as extra keys: configs[1] (1e6 rays, constant N, zero wind;
            ordered out-of-place as in round 1, dispersed in place, shuffled, the driver loop `advance`), a
            configs[4] deletion cycle, and the HPROP / online-saturation regimes.
  cpu_baseline  the oracle port (C restatement of the reference) on this box's host cores, bounded sample.

--impl reference times the reference's CPU algorithm (the oracle port; the reference itself is Python and cannot
travel to the GPU box) with all host threads on a bounded sample of the same workload.
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
# algorithmic bytes per ray (every per-ray fp64 field that must be read once + every field that changes
# written once; hand-over traffic between the two sweeps is implementation traffic and is NOT counted)
ALG = {
    "c1": {"step": 96.0, "A": 72.0, "B": 88.0},      # constant N: 10 fields read, rr, mm written
    "c2": {"step": 112.0, "A": 72.0, "B": 104.0},    # N(z): 10 fields read, rr, drr, mm, dmm written
}
# per-ray dram__bytes_read.sum + dram__bytes_write.sum and fp64-pipe utilisation from the committed ncu --set full
# captures (profiles/): used for roofline.traffic and roofline.fp64 (a profiler number is never a bench value; these
# only scale the live CUDA-event durations)
NCU = {
    "c1": {"A": {"traffic_per_ray": 90.7, "fp64_frac": 0.350, "us_per_mray": 42.4},
           "B": {"traffic_per_ray": 107.6, "fp64_frac": 0.327, "us_per_mray": 32.7},
           "source": "profiles/r02am_const_dispersed_ncu_full_summary.json (3e6 rays, constant N, sheared, dispersed)"},
    "c2": {"A": {"traffic_per_ray": 125.1, "fp64_frac": 0.416, "us_per_mray": 62.7},
           "B": {"traffic_per_ray": 157.8, "fp64_frac": 0.432, "us_per_mray": 42.2},
           "source": "profiles/r02bw_nz_bench_ncu_full_summary.json (the bench workload itself: 1.25e7 rays, dispersed)"},
}
# one fp64 warp instruction per 2.1 cycles per SM sub-partition, 4 per SM, 148 SMs (profiles/r01_fp64_pipe_microbench.txt)
FP64_PEAK_SOURCE = "measured DFMA issue rate, tools/micro/fp64_lat.cu (1 warp instruction / 2.1 cycles / SMSP) x 592 SMSPs x SM clock"
WORKLOADS = {
    "c2": "configs[2] physics: N^2(z) profile, sheared U(z), flux deposition on a 1000-level grid, %s ray volumes per GPU (8 GPUs: the 1e8 rays of configs[3]), in-place steady state",
    "c1": "configs[1]: %s ray volumes per GPU, 1-D column, constant N, zero mean wind, G=1000, in-place steady state",
}


def fmt_rays(n):
    return ("%.3g" % n).replace("e+0", "e").replace("e+", "e")


def config_dict(args, world):
    """identical for both arms (the driver compares it)"""
    return {"workload": WORKLOADS[args.workload] % fmt_rays(args.rays), "rays_per_gpu": args.rays, "grid_levels": 1000,
            "dt_s": 120.0, "regime": "in place after >= 30 in-place warm-up steps (dispersed ensemble)",
            "l2": "per-GPU state (%.2f GB) larger than L2; no flush needed" % (args.rays * 14 * 8 / 1e9),
            "parallelism": "rays sharded over %d rank(s), contiguous index ranges; sum of the deposited flux over ranks twice per step" % world,
            "mode": "M1 coupled (reference RK3 semantics: mean flow inside the RK state), 2 ray sweeps per step"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """threads the CPU arm may use: the affinity mask, not OMP_NUM_THREADS (torchrun sets that to 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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
        self.active = False          # samples are kept only while a timed region is open
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
                        if self.active:
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


def make_scenario(workload, n, rank, **kw):
    from msgwam_b200 import scenarios
    if workload == "c2":
        return scenarios.nz_sheared_ensemble(n, seed=1234 + rank, **kw)
    return scenarios.column_ensemble(n, seed=1234 + rank, ngrid=1001, **kw)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------
def oracle_rate(workload, n_sample, steps, warmup, nthreads, budget_s=30.0, min_s=0.0):
    """ray-steps/s of the oracle port on a bounded sample of the workload (in-place stepping, like the GPU arm):
    `steps` steps, more until min_s seconds have passed, never beyond budget_s"""
    import oracle
    sc = make_scenario(workload, n_sample, 0)
    orc = oracle.Oracle(sc.oracle_cfg(), nthreads=nthreads)
    var = sc.var()
    for _ in range(max(warmup, 1)):
        var = orc.RK3(sc.dt, var)
    t0 = time.perf_counter()
    done = 0
    while True:
        var = orc.RK3(sc.dt, var)
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s or (done >= steps and el >= min_s):
            break
    dt = time.perf_counter() - t0
    return n_sample * done / dt, done, dt


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = host_threads()
    n = min(args.rays, 2_000_000)                              # bounded sample per step
    steps = min(args.steps, 20)
    rate, done, dt = oracle_rate(args.workload, n, steps, min(args.warmup, 2), threads, budget_s=60.0, min_s=min(15.0, 0.5 * args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "oracle/msgwam_oracle.c (C restatement of lib/libprop.py RK3 + rhs_default + wave_projection; the "
                                   "Python reference cannot run on the GPU box and does ~2.3e4 ray-steps/s, BASELINE.md) on %d rays of the "
                                   "workload per step, %d in-place steps in %.1f s, %d OpenMP threads (affinity mask)" % (n, done, dt, threads)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def bind_numa(local_rank):
    """keep this rank's host threads (and the pages they touch) on the CPUs nearest its GPU: the pinned staging buffers
    of the e2e path then sit on the right socket.  NVML knows the affinity; failures are ignored."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def run_ours(args, rank, local_rank, world):
    numa_cpus = bind_numa(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import msgwam_b200.libprop as lprop
    from msgwam_b200 import scenarios
    from msgwam_b200._engine import Engine
    from msgwam_b200._cabi import check, lib
    from msgwam_b200.distributed import PeerExchange, rk3_host_sharded, shard_range
    from msgwam_b200.ensemble import RayEnsemble

    eng = Engine.get()
    dev = eng.device
    P = eng.ptr
    wl = args.workload
    n = args.rays

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity at this number of ranks, before anything is timed ---------------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_check(args, rank, world, dist if world > 1 else None)
        if rank == 0 and not parity["ok"]:
            print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity": parity, "n_gpus": world}), flush=True)
        if not parity["ok"]:
            if world > 1:
                dist.destroy_process_group()
            sys.exit(1)

    # ---- the workload, resident in HBM --------------------------------------------------------------------------
    sc = make_scenario(wl, n, rank)
    ens = RayEnsemble.from_scenario(sc)
    exchange = ens.exchange
    dt_s = sc.dt
    host_state = sc.state           # kept for the e2e leg
    k_warm = max(args.warmup, 30)
    ens.step(dt_s, k_warm)
    barrier()
    ens.check_errors()
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid) if rank == 0 else None

    # ---- value: K in-place steps between barriers, device time, max over ranks --------------------------------
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if sampler:
        sampler.active = True
    wall0 = time.perf_counter()
    ev0.record()
    for a, b in per:
        a.record(); ens.step(dt_s); b.record()
    ev1.record()
    barrier()
    wall = time.perf_counter() - wall0
    if sampler:
        sampler.active = False
    t_steps = maxr(ev0.elapsed_time(ev1) * 1e-3)
    per_ms = [a.elapsed_time(b) for a, b in per]
    launches = eng.launches - launches0
    ens.check_errors()                                      # the peer exchange's error word must be clean
    finite = bool(torch.isfinite(ens.uu).all() and torch.isfinite(ens.field("rr")).all() and torch.isfinite(ens.field("mm")).all())

    # ---- per-kernel durations for the roofline: an event between the two launches of a step ---------------------
    # (separate loop: the event record between the sweeps defeats the programmatic overlap of pass B's prologue with
    # pass A's tail, so the two durations add up to a little more than a fused step)
    mid, e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for e in (mid, e0, e1):
        e.record()
    torch.cuda.synchronize()
    ka, kb = [], []
    for _ in range(max(3, min(args.steps, 10))):
        barrier()
        check(lib.msgwam_debug_mid_event(mid.cuda_event), "msgwam_debug_mid_event")
        e0.record(); ens.step(dt_s); e1.record()
        check(lib.msgwam_debug_mid_event(None), "msgwam_debug_mid_event")
        torch.cuda.synchronize()
        ka.append(e0.elapsed_time(mid)); kb.append(mid.elapsed_time(e1))
    t_a, t_b = maxr(statistics.mean(ka) * 1e-3), maxr(statistics.mean(kb) * 1e-3)
    ens.check_errors()

    # ---- e2e: the reference-facing call with host buffers --------------------------------------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    sc.install(lprop)
    G = ens.G
    nz = wl == "c2"
    n_up = 8                                                 # dens, phi, rr, drr, kk, ll, mm, dmm (lam stays on the host)
    n_down = 4 if nz else 2                                  # rr, mm (+ drr, dmm with N(z))
    grid_bytes = (G + 1 + 6 * G + (G if nz else 0)) * 8
    if world == 1:
        def e2e_step(v):
            return lprop.RK3(dt_s, v)
    else:
        def e2e_step(v):
            return rk3_host_sharded(lprop, dt_s, v)

    def e2e_leg(pin, frozen, k):
        arrs = list(host_state) + [sc.uu, sc.vv, sc.dkk, sc.dll, sc.rr_mm_area]
        keep = [pinned(np.ascontiguousarray(a)) for a in arrs] if pin else None
        host = [t.numpy() for t in keep] if pin else [np.array(a, dtype=np.float64) for a in arrs]
        var = np.empty(11, dtype=object)
        for i in range(11):
            var[i] = host[i]
        lprop.set_statics(dkk=host[11], dll=host[12], rr_mm_area=host[13])
        lprop.freeze_statics(frozen)
        outs = [e2e_step(var) for _ in range(2)]             # warm-up (result buffers enter the pinned allocator's cache)
        del outs
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            out = e2e_step(var)
            _ = float(out[9][0])                             # the step's result is read on the host
        torch.cuda.synchronize()
        t = maxr(time.perf_counter() - t0)
        lprop.freeze_statics(False)
        h2d = (n_up + (0 if frozen else 2)) * n * 8 + grid_bytes
        d2h = n_down * n * 8 + 2 * G * 8
        return {"value": n * world * k / t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k,
                "ms_per_step": t / k * 1e3}
    k_e2e = max(3, min(args.steps, 5 if n > 4_000_000 else 10))
    if sampler:
        sampler.active = True
    e2e = e2e_leg(True, True, k_e2e)
    e2e["api"] = ("msgwam_b200.libprop.RK3(dt, var)" if world == 1 else "msgwam_b200.distributed.rk3_host_sharded(lprop, dt, var) (lprop.RK3 on this rank's slice)") + \
        " with page-locked numpy buffers, statics frozen (lprop.freeze_statics())"
    if sampler:
        sampler.active = False
    e2e_page = None
    if not args.no_extras:
        e2e_page = e2e_leg(False, False, max(2, k_e2e // 2))
        e2e_page["api"] = "the same call with ordinary pageable numpy arrays and the default statics semantics (dkk, dll uploaded every call): what an unmodified driver script gets"
        e2e["pageable"] = e2e_page
    del host_state

    # ---- the other BASELINE configurations and regimes (N = 1; reported beside the headline) --------------------
    extras = None
    if world == 1 and not args.no_extras:
        del ens
        torch.cuda.empty_cache()
        extras = extra_configs(args, eng, torch)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        ns = min(n, 2_000_000)
        r1, d1, s1 = oracle_rate(wl, min(ns, 500_000), 3, 1, 1, budget_s=8.0)
        rt, dn, sn = oracle_rate(wl, ns, 10, 1, threads, budget_s=30.0, min_s=12.0)
        cpu = {"value": rt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle port (C, -O2, no FMA contraction) on %d rays of the workload, %d in-place RK3 steps in %.1f s with %d OpenMP "
                         "threads; one thread: %.3e ray-steps/s; the unmodified Python reference does ~2.3e4 on one core (BASELINE.md)" % (
                             ns, dn, sn, threads, r1),
               "one_thread": r1, "host_cpus": os.cpu_count()}

    peak, peak_src = measured_peaks()
    total_rays = n * world
    value = total_rays * args.steps / t_steps
    names = {"c2": ("column_pass_nz<0> (pass A: deposit r0, stage 1, deposit r1, hand-over)",
                    "column_pass_nz<1> (pass B: mean-flow chain, stages 2-3, deposit r2, store)"),
             "c1": ("column_pass<0> (pass A: deposit r0, stage 1, deposit r1, hand-over)",
                    "column_pass<1> (pass B: mean-flow chain, stages 2-3, deposit r2, store)")}[wl]
    kern = {"A": {"name": names[0], "ms": t_a * 1e3, "alg": ALG[wl]["A"]}, "B": {"name": names[1], "ms": t_b * 1e3, "alg": ALG[wl]["B"]}}
    dom = "A" if t_a >= t_b else "B"
    oth = "B" if dom == "A" else "A"
    ach = kern[dom]["alg"] * n / (kern[dom]["ms"] * 1e-3) / 1e9
    clk = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 592 / 2.1 * clk * 1e6                        # fp64 warp instructions per second
    ncu = NCU[wl]

    def fp64_of(k):
        # fp64-pipe utilisation of the captured launch, rescaled by (captured time per ray / live time per ray)
        live_us_per_mray = kern[k]["ms"] * 1e3 / (n / 1e6)
        frac = ncu[k]["fp64_frac"] * ncu[k]["us_per_mray"] / live_us_per_mray
        return {"achieved": frac * fp64_peak, "peak": fp64_peak, "frac": frac, "unit": "fp64 warp-instructions/s",
                "source": "%s; utilisation of the ncu capture (%s) rescaled to the live duration" % (FP64_PEAK_SOURCE, ncu["source"])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": k_warm,
        "ms_per_step": t_steps / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "exchange": "none (one rank)" if world == 1 else ("16-byte self-validating cells pushed over NVLink peer memory inside the two sweeps (no collective kernel)" if exchange is not None else "NCCL all-reduce between the kernels"),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": ncu[dom]["traffic_per_ray"] * n, "traffic_source": "ncu --set full, " + ncu["source"],
                     "kernel": kern[dom]["name"], "algorithmic_bytes_per_ray": kern[dom]["alg"],
                     "kernel_ms": kern[dom]["ms"], "peak_source": peak_src,
                     "fp64": fp64_of(dom),
                     "note": "the sweeps are bound by fp64 instruction issue and shared-memory latency, not by HBM (see fp64 and profiles/r02_summary.md)",
                     "other_kernels": {kern[oth]["name"]: {"ms": kern[oth]["ms"], "achieved_gbs": kern[oth]["alg"] * n / (kern[oth]["ms"] * 1e-3) / 1e9,
                                                           "algorithmic_bytes_per_ray": kern[oth]["alg"],
                                                           "traffic": ncu[oth]["traffic_per_ray"] * n, "fp64": fp64_of(oth)}},
                     "step": {"algorithmic_bytes_per_ray_step": ALG[wl]["step"],
                              "achieved_gbs": ALG[wl]["step"] * n * args.steps / t_steps / 1e9,
                              "frac": ALG[wl]["step"] * n * args.steps / t_steps / 1e9 / peak}},
        "e2e": e2e,
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clocks,
        "step_ms_stats": {"min": min(per_ms), "median": statistics.median(per_ms), "max": max(per_ms), "note": "this rank's per-step event times"},
        "wall_s_timed_region": wall,
        "state_finite": finite,
    }
    if numa_cpus is not None:
        line["numa"] = "each rank bound to the %d CPUs nearest its GPU (NVML affinity)" % numa_cpus
    if extras is not None:
        line["configs"] = extras
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_check(args, rank, world, dist):
    """CUDA path vs the CPU oracle at this number of ranks: ~2e5 rays of the workload's physics (with a feeding-back
    amplitude), sharded like the bench, 3 in-place steps; the oracle runs the whole ensemble on rank 0."""
    import torch
    from msgwam_b200 import scenarios
    from msgwam_b200.distributed import shard_range
    from msgwam_b200.ensemble import RayEnsemble
    n_tot, nsteps = 200_003, 3
    if args.workload == "c2":
        sc = scenarios.nz_sheared_ensemble(n_tot, seed=77, amplitude=0.3)
    else:
        sc = scenarios.column_ensemble(n_tot, seed=77, ngrid=1001, sheared=True, amplitude=0.3)
    b, e = shard_range(sc.n, rank, world)
    ens = RayEnsemble([a[b:e] for a in sc.state], sc.dkk[b:e], sc.dll[b:e], sc.rr_mm_area[b:e], sc.uu, sc.vv, sc.grid, sc.grids,
                      sc.rhobar, sc.pressure_gradient, bvf=sc.model["bvf"], phi0=sc.model["phi0"])
    ens.step(sc.dt, nsteps)
    mine = ens.to_var()                                       # reads the error word, too
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, [mine[i] for i in range(11)])
    else:
        gathered = [[mine[i] for i in range(11)]]
    res = {"ok": True}
    if rank == 0:
        import oracle
        t0 = time.perf_counter()
        orc = oracle.Oracle(sc.oracle_cfg(), nthreads=host_threads())
        want = start = sc.var()
        for _ in range(nsteps):
            want = orc.RK3(sc.dt, want)
        ray_err, grid_err = 0.0, 0.0
        for i in range(11):
            if i >= 9:
                scale = max(float(np.max(np.abs(want[i]))), 1e-300)
                for r in range(world):
                    grid_err = max(grid_err, float(np.max(np.abs(gathered[r][i] - want[i]))) / scale)
            else:
                got = np.concatenate([gathered[r][i] for r in range(world)])
                # relative to the value or to its change over the run, whichever is larger (slots that pass through zero)
                scale = np.maximum(np.abs(want[i]), np.abs(want[i] - start[i]))
                diff = np.abs(got - want[i])
                ray_err = max(ray_err, float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale)))))
        same = all(np.array_equal(gathered[r][9], gathered[0][9]) and np.array_equal(gathered[r][10], gathered[0][10]) for r in range(1, world))
        tol_ray, tol_grid = 1e-10, 1e-10
        res = {"ok": bool(ray_err <= tol_ray and grid_err <= tol_grid and same), "rays": n_tot, "steps": nsteps, "ranks": world,
               "max_ray_rel_err": ray_err, "max_grid_rel_err": grid_err, "tol_ray": tol_ray, "tol_grid": tol_grid,
               "mean_flow_identical_on_all_ranks": bool(same),
               "oracle": "oracle port (pinned bit-for-bit to the Python reference for constant N; the N(z) terms are an extension "
                         "pinned to an independent numpy restatement), %d threads, %.1f s" % (host_threads(), time.perf_counter() - t0)}
    del ens
    # ---- phase 2: skewed deletion -> re-balancing over the ranks -> the driver loop with the fused post-step clamp ----
    sc = scenarios.nz_sheared_ensemble(120_011, seed=78, amplitude=1.0) if args.workload == "c2" else \
        scenarios.column_ensemble(120_011, seed=78, ngrid=1001, sheared=True, amplitude=1.0)
    sc.state[1] = np.arange(sc.n, dtype=np.float64)          # lam is inert in column mode: it carries the ray's identity
    b, e = shard_range(sc.n, rank, world)
    ens = RayEnsemble([a[b:e] for a in sc.state], sc.dkk[b:e], sc.dll[b:e], sc.rr_mm_area[b:e], sc.uu, sc.vv, sc.grid, sc.grids,
                      sc.rhobar, sc.pressure_gradient, bvf=sc.model["bvf"], phi0=sc.model["phi0"])
    kept = ens.compact(sc.dt, float(np.quantile(np.abs(sc.state[7]), 0.2 if rank == 0 else 0.9)))    # rank 0 loses most of its rays
    after = ens.rebalance()
    ens.advance(sc.dt, 2, saturate=True)
    mine = ens.to_var()
    pack = ([mine[i] for i in range(11)], kept, after)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, pack)
    else:
        gathered = [pack]
    if rank == 0:
        import oracle
        got = [np.concatenate([gathered[r][0][i] for r in range(world)]) for i in range(9)]
        order = np.argsort(got[1])
        sel = got[1][order].astype(np.int64)
        counts, afters = [g[1] for g in gathered], [g[2] for g in gathered]
        cfg = sc.oracle_cfg()
        cfg.update(dkk=sc.dkk[sel], dll=sc.dll[sel], rr_mm_area=sc.rr_mm_area[sel])
        orc = oracle.Oracle(cfg, nthreads=host_threads())
        var = np.empty(11, dtype=object)
        for i in range(9):
            var[i] = sc.state[i][sel]
        var[9], var[10] = sc.uu, sc.vv
        start = [np.array(a) for a in var[:9]]
        clamped = 0
        for _ in range(2):
            out = orc.RK3(sc.dt, var)
            dens = orc.saturation(sc.dt, out[0], var[3], (out[3] - var[3]) / 1, var[4], (out[4] - var[4]) / sc.dt, out[5], out[6],
                                  var[7], (out[7] - var[7]) / sc.dt, direct=True)
            clamped += int(np.count_nonzero(dens != out[0]))
            out[0] = dens
            var = out
        ray2, grid2 = 0.0, 0.0
        for i in range(11):
            if i >= 9:
                scale = max(float(np.max(np.abs(var[i]))), 1e-300)
                for r in range(world):
                    grid2 = max(grid2, float(np.max(np.abs(gathered[r][0][i] - var[i]))) / scale)
            else:
                g = got[i][order]
                scale = np.maximum(np.abs(var[i]), np.abs(var[i] - start[i]))
                diff = np.abs(g - var[i])
                ray2 = max(ray2, float(np.max(np.where(diff == 0, 0.0, diff / np.where(scale == 0, 1.0, scale)))))
        ok2 = bool(len(np.unique(sel)) == len(sel) == sum(counts) == sum(afters) and max(afters) - min(afters) <= 1 and
                   ray2 <= 1e-10 and grid2 <= 1e-10 and clamped > 0)
        res["deletion_rebalance_advance"] = {"ok": ok2, "rays_before_deletion": sc.n, "survivors_per_rank": counts,
                                             "after_rebalance": afters, "steps": 2, "rays_clamped": clamped,
                                             "max_ray_rel_err": ray2, "max_grid_rel_err": grid2}
        res["ok"] = bool(res["ok"] and ok2)
    if world > 1:
        flag = torch.tensor([1 if res["ok"] else 0], dtype=torch.int32, device="cuda")
        dist.broadcast(flag, 0)
        if rank != 0:
            res = {"ok": bool(int(flag.item()))}
    del ens
    return res


def extra_configs(args, eng, torch):
    """N = 1: BASELINE configs[1] and [4] and the general-mode regimes, median GPU time per step."""
    import msgwam_b200.libprop as lprop
    from msgwam_b200 import scenarios
    from msgwam_b200._cabi import check, lib
    from msgwam_b200.ensemble import RayEnsemble
    P = eng.ptr
    out = {}
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=eng.device)

    def timed(fn, k, flush_l2=True):
        ts = []
        for _ in range(k):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        return statistics.median(a.elapsed_time(b) for a, b in ts) * 1e-3

    def entry(n, t, **kw):
        return dict({"ms_per_step": t * 1e3, "value": n / t, "unit": UNIT}, **kw)

    k = max(5, min(args.steps, 20))
    # ---- configs[1]: 1e6 rays, constant N, zero wind (the round-1 headline and its regimes) ----
    n1 = 1_000_000
    sc1 = scenarios.column_ensemble(n1, seed=1234, ngrid=1001)
    e1 = RayEnsemble.from_scenario(sc1)
    p1, g1, r1 = e1.params(sc1.dt), eng.grid_struct(e1.grid_devs), e1._rays()
    rr_o, mm_o = eng.empty(n1), eng.empty(n1)

    def oop():
        check(lib.msgwam_column_step(p1, r1, n1, g1, P(e1.uu), P(e1.vv), P(e1.work), P(rr_o), P(mm_o), P(e1._uu2), P(e1._vv2), eng.stream), "column_step")
    e1.measure_bounds(sc1.dt)
    timed(oop, 5)
    out["c1_ordered_out_of_place"] = entry(n1, timed(oop, k), note="round-1 headline regime: height-ordered ensemble, same input every step, L2 flushed")
    e1.step(sc1.dt, 30)
    out["c1_dispersed_in_place"] = entry(n1, timed(lambda: e1.step(sc1.dt), k), note="after 30 in-place steps, L2 flushed")
    out["c1_advance"] = entry(n1, timed(lambda: e1.advance(sc1.dt, 1), k), note="driver loop on the device: RK3 + post-step saturation clamp (raytracer.py:175-188), dispersed, L2 flushed")
    out["c1_frozen_background_M2"] = entry(n1, timed(lambda: e1.step_frozen(sc1.dt), k), note="EXTENSION, a different scheme (mean flow frozen over the step, one deposit per step; never answers for RK3): one sweep, one launch per step; dispersed, L2 flushed")
    scs = scenarios.column_ensemble(n1, seed=1234, ngrid=1001, shuffled=True)
    es = RayEnsemble.from_scenario(scs)
    ps, gs_, rs = es.params(scs.dt), eng.grid_struct(es.grid_devs), es._rays()

    def oop_s():
        check(lib.msgwam_column_step(ps, rs, n1, gs_, P(es.uu), P(es.vv), P(es.work), P(rr_o), P(mm_o), P(es._uu2), P(es._vv2), eng.stream), "column_step")
    es.measure_bounds(scs.dt)
    timed(oop_s, 3)
    out["c1_shuffled_out_of_place"] = entry(n1, timed(oop_s, k), note="random ray order, L2 flushed")
    del e1, es
    # ---- constant N at the size of configs[2] (1e7 rays): the coupled step and the frozen-background mode ----
    n7 = 10_000_000
    sc7 = scenarios.column_ensemble(n7, seed=1234, ngrid=1001, sheared=True, amplitude=0.01)
    e7 = RayEnsemble.from_scenario(sc7)
    del sc7.state
    e7.step(120.0, 30)
    out["constN_sheared_1e7_in_place"] = entry(n7, timed(lambda: e7.step(120.0), k, flush_l2=False), note="constant N, sheared wind, 1e7 rays, dispersed; A = 96 B/ray-step -> %.0f GB/s algorithmic" % 0.0)
    t7 = out["constN_sheared_1e7_in_place"]["ms_per_step"] * 1e-3
    out["constN_sheared_1e7_in_place"]["note"] = "constant N, sheared wind, 1e7 rays, dispersed; 96 B/ray-step algorithmic = %.0f GB/s" % (96.0 * n7 / t7 / 1e9)
    tf = timed(lambda: e7.step_frozen(120.0), k, flush_l2=False)
    out["constN_sheared_1e7_frozen_background_M2"] = entry(n7, tf, note="EXTENSION (see c1_frozen_background_M2); 96 B/ray-step algorithmic = %.0f GB/s, real traffic 88 B/ray" % (96.0 * n7 / tf / 1e9))
    del e7
    # ---- general-mode regimes: HPROP on, online saturation on, 1e6 rays ----
    for key, kw in (("hprop", dict(hprop=True)), ("saturate_online", dict(saturate_online=True))):
        scg = scenarios.column_ensemble(n1, seed=1234, ngrid=1001, sheared=True, amplitude=0.3, phi0=np.deg2rad(-40.0))
        if "hprop" in kw:
            scg.hprop = True
        else:
            scg.model = dict(scg.model, saturate_online=True)
        eg = RayEnsemble.from_scenario(scg)
        eg.step(scg.dt, 3)
        out["general_" + key] = entry(n1, timed(lambda: eg.step(scg.dt), max(3, k // 2)), note="stage-by-stage kernels (rhs + deposit + low-storage update per stage), in place, L2 flushed")
        del eg
    # ---- configs[4]: critical-level stress case with deletion by stream compaction ----
    n4 = args.c4_rays
    t0 = time.perf_counter()
    sc4 = scenarios.critical_level_ensemble(n4, ngrid=1001, stress=True)
    e4 = RayEnsemble.from_scenario(sc4)
    del sc4.state
    build_s = time.perf_counter() - t0
    cyc = []
    nf = int(e4._slab.shape[0])                             # fields of the store: 9 state + 3 statics + 2 derived
    e4.compact(120.0, float("inf"))                         # first call allocates the second slab (not timed)
    for c in range(3):
        a, b, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        before = e4.n
        torch.cuda.synchronize()
        a.record(); e4.step(120.0, 10); b.record()
        after = e4.compact(120.0, scenarios.M_CRIT_STRESS)
        c2.record(); torch.cuda.synchronize()
        ms_s, ms_c = a.elapsed_time(b), b.elapsed_time(c2)
        cyc.append({"rays_before": before, "survivors": after, "ms_10_steps": ms_s, "ray_steps_per_s": 10 * before / (ms_s * 1e-3),
                    "ms_compaction": ms_c, "compaction_gbs": nf * 8 * (before + after) / (ms_c * 1e-3) / 1e9})
    out["c4_critical_level"] = {"rays": n4, "cycles": cyc, "host_build_s": build_s,
                                "note": "10 in-place steps, then deletion of rays outside the deposit domain or with |m| >= m_crit (flag + stable compaction of the %d-field store + the host read of the survivor count: bytes = %d fields x 8 B x (rays read + survivors written))" % (nf, nf)}
    del e4
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c1"], help="c2: N^2(z) + sheared wind (default); c1: constant N, zero wind")
    ap.add_argument("--rays", type=int, default=None, help="ray volumes per GPU (default 1.25e7 for c2, 1e6 for c1)")
    ap.add_argument("--c4-rays", type=int, default=50_000_000, help="rays of the configs[4] deletion cycle (N = 1 extras)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs / regimes and the pageable e2e leg")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.rays is None:
        args.rays = 12_500_000 if args.workload == "c2" else 1_000_000
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
