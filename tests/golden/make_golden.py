"""Generate the golden fixtures in this directory from the UNMODIFIED Python reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Everything written here is an output of /root/reference/lib/libprop.py (and, for
`driver_history.npz`, of /root/reference/raytracer.py executed unmodified through runpy
with matplotlib stubbed out, because matplotlib is not installed).  Inputs are stored
next to outputs so the fixtures are self-contained on the GPU box, where the
reference does not exist.
"""
import contextlib
import io
import os
import runpy
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))

from _reference import REFERENCE_ROOT, load_reference  # noqa: E402
from msgwam_b200 import scenarios  # noqa: E402


def obj11(arrs):
    out = np.empty(11, dtype=object)
    for i, a in enumerate(arrs):
        out[i] = a
    return out


def random_case(rng, n, G, hprop, saturate_online, phi_spread):
    """A small sheared column with all fields non-trivial."""
    grid = np.linspace(0, 40e3, G + 1)
    grids = .5 * (grid[:-1] + grid[1:])
    phi0 = np.deg2rad(-35.0) if phi_spread else 0.0
    rr = np.sort(rng.uniform(-500., 41e3, n))           # some rays outside the domain
    drr = rng.uniform(100., 2500., n)
    mm = -2 * np.pi / rng.uniform(1e3, 10e3, n) * rng.choice([-1., 1.], n)
    kh = 2 * np.pi / rng.uniform(20e3, 200e3, n)
    th = rng.uniform(0, 2 * np.pi, n)
    kk, ll = kh * np.sin(th), kh * np.cos(th)
    dmm = rng.uniform(1e-5, 3e-4, n) * np.abs(mm)
    area = dmm * drr
    lam = rng.uniform(-0.1, 0.1, n)
    phi = phi0 + (rng.uniform(-0.05, 0.05, n) if phi_spread else 0.0) * np.ones(n)
    dens = rng.uniform(0.5, 2.0, n) * 1e9
    dkk = rng.uniform(0.5e-4, 2e-4, n)
    dll = rng.uniform(0.5e-4, 2e-4, n)
    uu = 20. * np.tanh((grids - 20e3) / 5e3) + 3. * np.sin(grids / 3e3)
    vv = 5. * np.cos(grids / 4e3)
    rhobar = 1.2 * np.exp(-grids / 8500.)
    ff = 2 * 7.2921e-5 * np.sin(phi0)
    pg = np.empty((2, G))
    pg[0] = rhobar * ff * vv
    pg[1] = -rhobar * ff * uu
    model = dict(bvf=0.012, phi0=phi0, kappa=0.9, saturate_online=saturate_online)
    if saturate_online:
        dens = saturation_scale(rng, model, rr, kk, ll, mm, dkk, dll, dmm, grids, rhobar)
    return scenarios.Scenario("random", 60., [dens, lam, phi, rr, drr, kk, ll, mm, dmm], uu, vv, dkk, dll, area,
                              grid, grids, rhobar, pg, model, hprop=hprop)


def saturation_scale(rng, model, rr, kk, ll, mm, dkk, dll, dmm, grids, rhobar):
    """Wave-action densities scattered around the static-instability limit so that roughly half of the
    rays trigger the clamp of saturation() (L:604)."""
    f0 = 2 * 7.2921e-5 * np.sin(model["phi0"])
    omh = np.sqrt((model["bvf"] ** 2 * (kk ** 2 + ll ** 2) + f0 ** 2 * mm ** 2) / (kk ** 2 + ll ** 2 + mm ** 2))
    rho = np.interp(rr, grids, rhobar)
    limit = model["kappa"] ** 2 * .5 * rho * omh * model["bvf"] ** 2 / mm ** 2 / (omh ** 2 - f0 ** 2)
    return limit / (dkk * dll * dmm) * rng.uniform(0.3, 3.0, rr.shape)


def pack_scenario(sc, prefix=""):
    d = {prefix + k: v for k, v in zip(scenarios.STATE_NAMES, sc.state)}
    d.update({prefix + "uu": sc.uu, prefix + "vv": sc.vv, prefix + "dkk": sc.dkk, prefix + "dll": sc.dll,
              prefix + "rr_mm_area": sc.rr_mm_area, prefix + "grid": sc.grid, prefix + "grids": sc.grids,
              prefix + "rhobar": sc.rhobar, prefix + "pressure_gradient": sc.pressure_gradient,
              prefix + "dt": np.float64(sc.dt), prefix + "hprop": np.bool_(sc.hprop),
              prefix + "bvf": np.float64(sc.model["bvf"]), prefix + "phi0": np.float64(sc.model["phi0"]),
              prefix + "kappa": np.float64(sc.model.get("kappa", 1.0)),
              prefix + "saturate_online": np.bool_(sc.model.get("saturate_online", False))})
    return d


def main():
    ref = load_reference()
    assert ref is not None, "reference not found under %s" % REFERENCE_ROOT
    rng = np.random.default_rng(20261018)

    # ---- 1. the driver, unmodified (matplotlib stubbed) -----------------------------
    class _Anything:
        def __getattr__(self, name): return _Anything()
        def __call__(self, *a, **k): return _Anything()
        def __iter__(self): return iter((_Anything(), _Anything()))
        def __getitem__(self, i): return _Anything()
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
    plt.style = _Anything(); plt.subplots = lambda *a, **k: (_Anything(), _Anything())
    plt.colorbar = _Anything(); plt.show = lambda: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd(); os.chdir(REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):      # the driver prints a progress line per step
        g = runpy.run_path(os.path.join(REFERENCE_ROOT, "raytracer.py"), run_name="__main__")
    os.chdir(cwd); sys.path.remove(REFERENCE_ROOT)
    for m in ("lib", "lib.libprop"):
        sys.modules.pop(m, None)
    steps = np.array([0, 1, 2, 10, 100, 360, 720, 1440])
    np.savez_compressed(os.path.join(HERE, "driver_history.npz"), steps=steps,
                        **{k: g["int_" + k][steps] for k in ("dens", "dens_prop", "lambda", "phi", "rr", "drr", "kk", "ll", "mm", "dmm", "uu", "vv")},
                        init_dkk=g["init_dkk"], init_dll=g["init_dll"], rr_mm_area=g["rr_mm_area"],
                        grid=g["grid"], grids=g["grids"], rhobar=g["lprop"].rhobar,
                        pressure_gradient=g["lprop"].pressure_gradient,
                        wa_max=np.float64(g["wa"].max()), flux_diag_absmax=np.float64(np.abs(g["flux_diag"]).max()),
                        wa_rows=g["wa"][[0, 1, 100, 720]], flux_diag_rows=g["flux_diag"][[0, 1, 100, 720]])
    print("driver: int_rr[1,:3] =", g["int_rr"][1, :3], " wa.max =", g["wa"].max())

    # ---- 2. pure RK3 trajectory of the driver's initial condition -------------------
    sc = scenarios.default_column()
    sc.install(ref)
    assert all(np.array_equal(a, g["int_" + k][0]) for a, k in
               zip(sc.state, ("dens", "lambda", "phi", "rr", "drr", "kk", "ll", "mm", "dmm"))), "IC restatement differs"
    assert np.array_equal(sc.uu, g["int_uu"][0]) and np.array_equal(sc.rhobar, g["lprop"].rhobar)
    var = sc.var()
    out = pack_scenario(sc)
    for step in range(1, 721):
        var = ref.RK3(sc.dt, var)
        if step in (1, 2, 10, 100, 360, 720):
            for i, nm in enumerate(scenarios.STATE_NAMES + ("uu", "vv")):
                out["step%d_%s" % (step, nm)] = np.asarray(var[i], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "rk3_default_column.npz"), **out)

    # ---- 3. rhs_default / RK3 on random sheared columns, all mode combinations -------
    for tag, hprop, sat, spread in (("col", False, False, False), ("col_sat", False, True, False),
                                    ("col_phi", False, False, True), ("hprop", True, False, True),
                                    ("hprop_sat", True, True, True)):
        ref = load_reference()
        sc = random_case(rng, 257, 40, hprop, sat, spread)
        sc.install(ref)
        out = pack_scenario(sc)
        rhs = ref.rhs_default(sc.dt, sc.var())
        for i, nm in enumerate(scenarios.STATE_NAMES + ("uu", "vv")):
            out["rhs_" + nm] = np.asarray(rhs[i], dtype=np.float64) * np.ones_like(np.asarray(sc.var()[i], dtype=np.float64))
        var = sc.var()
        for step in (1, 2, 3):
            var = ref.RK3(sc.dt, var)
            for i, nm in enumerate(scenarios.STATE_NAMES + ("uu", "vv")):
                out["step%d_%s" % (step, nm)] = np.asarray(var[i], dtype=np.float64)
        # point functions on the same inputs
        dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
        out["omega"] = ref.omega(kk, ll, mm, phi)
        out["omega_phi0"] = ref.omega(kk, ll, mm, sc.model["phi0"])
        out["cg_rr"] = ref.cg_rr(kk, ll, mm, lam, phi, rr)
        out["cg_lambda"] = ref.cg_lambda(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        out["cg_phi"] = ref.cg_phi(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        out["dk_dt"] = ref.dk_dt(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        out["dl_dt"] = ref.dl_dt(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        out["dm_dt"] = ref.dm_dt(kk, ll, mm, lam, phi, rr, sc.uu, sc.vv)
        out["gradients"] = ref.gradients(lam, phi, rr, sc.uu, sc.vv)
        flux_grad = rng.normal(size=sc.grids.shape) * 1e-6
        out["flux_grad"] = flux_grad
        out["du_dt"] = ref.du_dt(sc.vv, flux_grad)
        out["dv_dt"] = ref.dv_dt(sc.uu, flux_grad)
        # saturation, both modes, with independent tendencies
        rr_st = rng.normal(size=rr.shape); drr_st = rng.normal(size=rr.shape) * 0.01; mm_st = rng.normal(size=rr.shape) * 1e-7
        out["sat_rr_st"], out["sat_drr_st"], out["sat_mm_st"] = rr_st, drr_st, mm_st
        big = saturation_scale(rng, sc.model, rr, kk, ll, mm, sc.dkk, sc.dll, dmm, sc.grids, sc.rhobar)
        out["sat_dens"] = big
        out["sat_tend"] = ref.saturation(sc.dt, big, rr, rr_st, drr, drr_st, kk, ll, mm, mm_st)
        out["sat_direct"] = ref.saturation(sc.dt, big, rr, rr_st, drr, drr_st, kk, ll, mm, mm_st, direct=True)
        # projections var = 0..4 on the staggered and on the full grid
        rl, ru = rr - .5 * drr, rr + .5 * drr
        ml, mu = mm - .5 * dmm, mm + .5 * dmm
        for var_id in range(5):
            out["proj%d_grids" % var_id] = ref.wave_projection(dens, lam, phi, rl.copy(), ru.copy(), kk, ll, ml, mu, sc.dkk, sc.dll, dmm, sc.grids, var=var_id)
            out["proj%d_grid" % var_id] = ref.wave_projection(dens, lam, phi, rl.copy(), ru.copy(), kk, ll, ml, mu, sc.dkk, sc.dll, dmm, sc.grid, var=var_id)
        np.savez_compressed(os.path.join(HERE, "random_%s.npz" % tag), **out)

    # ---- 4. deposition corner cases -------------------------------
    ref = load_reference()
    ref.HPROP_GLOBAL = False
    ref.set_model_setup(bvf=0.01, phi0=0.0)
    grid = np.linspace(0, 100e3, 101); grids = .5 * (grid[:-1] + grid[1:])
    rl = np.array([1200., 100., 98600., 96000., -3000., 99000., 100500., 600., 0., 2000., 49999.999999999, 50000., -10., 97000., 98000.])
    ru = np.array([1700., 350., 99400., 105000., -100., 101000., 104000., 1900., 0., 3000., 50000.000000001, 51000., 10., 98000., 99000.])
    n = len(rl)
    one = np.ones(n)
    kk = one * 2 * np.pi / 50e3; ll = one * 1e-5; mm = -one * 2 * np.pi / 5e3
    dmm = one * 1e-7; dkk = one * 1e-4; dll = one * 1e-4; dens = one * 1e9
    out = dict(rr_low=rl, rr_up=ru, kk=kk, ll=ll, mm=mm, dmm=dmm, dkk=dkk, dll=dll, dens=dens, grid=grid, grids=grids,
               bvf=np.float64(0.01), phi0=np.float64(0.0))
    for i in range(n):
        s = slice(i, i + 1)
        for var_id in (0, 1, 2):
            out["ray%d_proj%d_grids" % (i, var_id)] = ref.wave_projection(dens[s], 0 * one[s], 0 * one[s], rl[s].copy(), ru[s].copy(), kk[s], ll[s], mm[s] - .5 * dmm[s], mm[s] + .5 * dmm[s], dkk[s], dll[s], dmm[s], grids, var=var_id)
            out["ray%d_proj%d_grid" % (i, var_id)] = ref.wave_projection(dens[s], 0 * one[s], 0 * one[s], rl[s].copy(), ru[s].copy(), kk[s], ll[s], mm[s] - .5 * dmm[s], mm[s] + .5 * dmm[s], dkk[s], dll[s], dmm[s], grid, var=var_id)
    np.savez_compressed(os.path.join(HERE, "projection_corner_cases.npz"), **out)

    # ---- 5. np.interp semantics -------------------------------------------------------
    xp = np.linspace(1000., 39000., 39); fp = rng.normal(size=39)
    x = np.concatenate([rng.uniform(-2000, 42000, 500), xp, [xp[0], xp[-1], np.nextafter(xp[3], 0), np.nextafter(xp[3], 1e9)]])
    np.savez_compressed(os.path.join(HERE, "interp.npz"), x=x, xp=xp, fp=fp, y=np.interp(x, xp, fp))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
