"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host-derived
parameters match the oracle's, the setup helpers match the reference's outputs, there is no silent CPU
fallback, and the multi-rank plumbing (sharding + all-reduce of the deposit) works under gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

HEADER = os.path.join(ROOT, "include", "msgwam_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msgwam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from msgwam_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for nm in names:
        assert hasattr(lib, nm), nm
        assert nm in _cabi.SIGNATURES, "binding missing for " + nm
    assert lib.msgwam_abi_version() == _cabi.ABI_VERSION
    assert b"bad argument" in _cabi.lib.msgwam_error_string(-1)
    assert b"timed out" in _cabi.lib.msgwam_error_string(-4)
    assert _cabi.lib.msgwam_column_work_doubles(1000) >= 6 * 999
    assert _cabi.lib.msgwam_host_stage_doubles(1000, 100) > 14 * 1000


def test_library_is_sm100a_and_uses_no_legacy_paths():
    from msgwam_b200 import _cabi
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_params_snapshot_matches_oracle_derivation():
    import oracle
    from msgwam_b200 import _cabi, scenarios
    sc = scenarios.default_column()
    orc = oracle.Oracle(sc.oracle_cfg())
    p = _cabi.snapshot_params(sc.dt, bvf=sc.model["bvf"], phi0=np.deg2rad(-60), kappa=0.95, saturate_online=True,
                              hprop=True, grid=sc.grid, grids=sc.grids)
    cfg = sc.oracle_cfg(); cfg.update(phi0=np.deg2rad(-60), kappa=0.95, saturate_online=True, hprop=True)
    q = oracle.Oracle(cfg).p
    for k in ("n2", "two_rot", "rad_earth", "c8rot2", "f0", "f0sq", "k2half", "dz_grid", "dz_grids"):
        assert getattr(p, k) == getattr(q, k), k
    assert p.G == q.ngrid - 1 and p.hprop == 1 and p.saturate_online == 1
    assert p.inv_dz_grid == 1.0 / p.dz_grid and orc.p.f0 == 0.0


def test_setup_helpers_match_reference_outputs():
    """set_hydrostatics / set_pressure_gradient / velocities_sine_homogeneous against what the unmodified
    driver produced (tests/golden/driver_history.npz)."""
    import importlib
    import msgwam_b200.libprop as lp
    importlib.reload(lp)
    d = load_golden("driver_history.npz")
    assert lp.model_config["kappa"] == 0.95 and lp.model_config["saturate_online"] is True and lp.HPROP_GLOBAL is True
    assert lp.statics == dict(int_dll=1, int_dkk=1, rr_mm_area=0) and lp.model_config["rhs"] is lp.rhs_default
    lp.HPROP_GLOBAL = False
    lp.set_model_setup(bvf=0.01, rhs=lp.rhs_default, boussinesq=False, sig_rr=10000, u0=4, rr0=40000, rr1=40000,
                       phi0=np.deg2rad(0), kappa=1., saturate_online=False)
    lp.grid, lp.grids = d["grid"], d["grids"]
    uu = lp.velocities_sine_homogeneous(lp.grids)
    lp.set_hydrostatics()
    lp.set_pressure_gradient(uu, np.zeros(uu.shape))
    assert np.array_equal(uu, d["uu"][0])
    assert np.array_equal(lp.rhobar, d["rhobar"])
    assert np.array_equal(lp.pressure_gradient, d["pressure_gradient"])
    assert lp.get_model_setup() is lp.model_config


def test_velocity_generators_match_live_reference():
    from _reference import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("/root/reference not present")
    import importlib
    import msgwam_b200.libprop as lp
    importlib.reload(lp)
    z = np.linspace(0, 80e3, 333)
    lam, phi = np.zeros(7), np.linspace(-1.2, -0.9, 7)
    for nm in ("velocities_tanh_homogeneous", "velocities_gauss_homogeneous", "velocities_sine_homogeneous"):
        assert np.array_equal(getattr(lp, nm)(z.copy()), getattr(ref, nm)(z.copy())), nm
    assert np.array_equal(lp.velocities_tanh(lam, phi, z[:7]), ref.velocities_tanh(lam, phi, z[:7]))


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import importlib
    import msgwam_b200.libprop as lp
    from msgwam_b200 import _cabi, scenarios
    importlib.reload(lp)
    sc = scenarios.default_column()
    sc.install(lp)
    with pytest.raises(_cabi.MsgwamError):
        lp.RK3(sc.dt, sc.var())
    with pytest.raises(_cabi.MsgwamError):
        lp.wave_projection(*([np.ones(3)] * 12), sc.grid)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "python-msgwam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f


def test_shard_range_partitions_exactly():
    from msgwam_b200.distributed import shard_range
    for n in (0, 1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "python-msgwam_b200"))
import oracle
from msgwam_b200 import scenarios
from msgwam_b200.distributed import all_reduce_sum, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
sc = scenarios.column_ensemble(4001, seed=9, ngrid=101, sheared=True, amplitude=0.3)
b, e = shard_range(sc.n, rank, 2)
cfg = sc.oracle_cfg()
sub = dict(cfg, dkk=cfg["dkk"][b:e], dll=cfg["dll"][b:e], rr_mm_area=cfg["rr_mm_area"][b:e])
var = sc.var()
loc = np.empty(11, dtype=object)
for i in range(9): loc[i] = var[i][b:e]
loc[9], loc[10] = var[9], var[10]
_, proj = oracle.Oracle(sub).rhs_default(sc.dt, loc, return_projection=True)
t = torch.from_numpy(proj.copy())
all_reduce_sum(t)                                   # the data path's only collective
_, full = oracle.Oracle(cfg).rhs_default(sc.dt, var, return_projection=True)
err = np.max(np.abs(t.numpy() - full)) / np.max(np.abs(full))
assert err < 1e-13, err
dist.destroy_process_group()
print("rank", rank, "ok", err)
'''


def test_two_rank_gloo_deposit_allreduce(tmp_path):
    """world_size 2 on CPU (gloo): each rank deposits its contiguous shard, the all-reduce used by the GPU
    path sums them, and the result equals the single-rank deposit up to summation order."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_rebalance_plan_evens_out_any_counts():
    from msgwam_b200.distributed import rebalance_plan
    rng = np.random.default_rng(3)
    for counts in ([10, 0], [0, 0, 0], [5, 5, 5], [1000, 3, 0, 17, 250, 1, 999, 12], list(rng.integers(0, 10**6, 8))):
        plan, target = rebalance_plan(counts)
        after = list(counts)
        for src, dst, n in plan:
            assert n > 0 and src != dst
            after[src] -= n; after[dst] += n
        assert after == target and sum(after) == sum(counts) and max(after) - min(after) <= 1
        assert all(counts[src] > target[src] for src, _, _ in plan) and all(counts[dst] < target[dst] for _, dst, _ in plan)


_REBALANCE_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "python-msgwam_b200"))
from msgwam_b200.distributed import exchange_rows, rebalance_plan
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
counts = [1003, 10]                                 # deletion emptied rank 1
rows = torch.arange(counts[rank] * 3, dtype=torch.float64).reshape(counts[rank], 3) + 1e6 * rank
plan, target = rebalance_plan(counts)
kept, got = exchange_rows(rows, counts[rank], plan, rank, dist)
mine = torch.cat([kept, got])
assert mine.shape[0] == target[rank], (mine.shape, target)
allrows = [None, None]
dist.all_gather_object(allrows, mine.numpy())
u = np.concatenate(allrows)
want = np.concatenate([np.arange(c * 3, dtype=np.float64).reshape(c, 3) + 1e6 * r for r, c in enumerate(counts)])
assert u.shape == want.shape and np.array_equal(np.sort(u[:, 0]), np.sort(want[:, 0]))      # every ray exactly once
assert np.array_equal(u[np.argsort(u[:, 0])], want[np.argsort(want[:, 0])])                  # rows intact
dist.destroy_process_group()
print("rank", rank, "ok", mine.shape[0])
'''


def test_two_rank_gloo_rebalance_moves_whole_rays(tmp_path):
    """world_size 2 on CPU (gloo): the row exchange behind RayEnsemble.rebalance moves surplus rays, all fields of a ray
    together, and every ray ends up on exactly one rank."""
    script = tmp_path / "worker_rebalance.py"
    script.write_text(_REBALANCE_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    our arm's metric / unit / config, an e2e object without copies and the cpu_baseline that describes the run."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--rays", "3000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ray-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ray-steps/sec") and d["dtype"] == "f64" and d["value"] > 0
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]


def test_snapshot_params_memo_follows_every_input():
    """_cabi.snapshot_params memoises on its inputs (the host-buffer RK3 calls it once per step): every input that
    changes must change the snapshot, and a hit must equal a fresh computation byte for byte."""
    from msgwam_b200 import _cabi
    g = np.linspace(0.0, 100e3, 101)
    gs = .5 * (g[:-1] + g[1:])
    base = dict(bvf=0.01, phi0=0.3, kappa=1.0, saturate_online=False, hprop=False, grid=g, grids=gs)

    def fresh(dt, kw):
        return bytes(_cabi._snapshot_uncached(dt, kw["bvf"], kw["phi0"], kw["kappa"], kw["saturate_online"], kw["hprop"],
                                              kw["grid"], kw["grids"], _cabi.ROT_EARTH_DEFAULT, _cabi.RAD_EARTH_DEFAULT))

    assert bytes(_cabi.snapshot_params(120.0, **base)) == fresh(120.0, base)
    assert bytes(_cabi.snapshot_params(120.0, **base)) == fresh(120.0, base)          # the memo hit
    for name, value in (("bvf", 0.02), ("phi0", -0.3), ("phi0", -0.0), ("kappa", 2.0), ("saturate_online", True),
                        ("hprop", True), ("grid", g * 2), ("grids", gs[:50])):
        kw = dict(base, **{name: value})
        assert bytes(_cabi.snapshot_params(120.0, **kw)) == fresh(120.0, kw), name
        assert bytes(_cabi.snapshot_params(60.0, **kw)) == fresh(60.0, kw), name
    # +0.0 and -0.0 latitudes are different inputs (the sign of f0 follows)
    a = _cabi.snapshot_params(120.0, **dict(base, phi0=0.0))
    b = _cabi.snapshot_params(120.0, **dict(base, phi0=-0.0))
    assert np.signbit(b.f0) and not np.signbit(a.f0)
    # a profile-valued bvf (extension) is only looked at for its rank
    prof = np.full(gs.shape, 0.01)
    assert np.isnan(_cabi.snapshot_params(120.0, **dict(base, bvf=prof)).n2)
    assert _cabi.snapshot_params(120.0, **base).n2 == 0.01 ** 2
