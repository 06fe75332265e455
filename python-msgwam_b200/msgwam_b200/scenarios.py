"""Synthetic ray ensembles and the reference driver's initial condition.

Pure numpy input generators: they build the caller-side arrays (the 9 per-ray state
arrays, the mean flow on the staggered grid, the per-ray statics and the background
profiles) that ``libprop.RK3`` consumes.  They do not touch the device.

* ``default_column()`` restates the initial condition of the reference driver
  (``/root/reference/raytracer.py:32-117``; BASELINE.json configs[0]).
* ``column_ensemble()`` is the synthetic 1-D column ensemble of SURVEY.md section 8(d)
  (configs[1]: constant N, zero wind; ``sheared=True`` adds the tanh jet + sine shear
  wind of configs[2] on the same rays).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

ROT_EARTH = 7.2921e-5

STATE_NAMES = ("dens", "lam", "phi", "rr", "drr", "kk", "ll", "mm", "dmm")


@dataclass
class Scenario:
    name: str
    dt: float
    state: list                     # 9 arrays, STATE_NAMES order
    uu: np.ndarray
    vv: np.ndarray
    dkk: np.ndarray
    dll: np.ndarray
    rr_mm_area: np.ndarray
    grid: np.ndarray
    grids: np.ndarray
    rhobar: np.ndarray
    pressure_gradient: np.ndarray
    model: dict = field(default_factory=dict)   # bvf, phi0, kappa, saturate_online (+ background keys)
    hprop: bool = False

    @property
    def n(self) -> int:
        return int(self.state[0].shape[0])

    def var(self) -> np.ndarray:
        """The 11-slot object array the reference driver builds (raytracer.py:160-172)."""
        out = np.empty(11, dtype=object)
        for i, a in enumerate(self.state):
            out[i] = a
        out[9], out[10] = self.uu, self.vv
        return out

    def oracle_cfg(self) -> dict:
        return dict(bvf=self.model["bvf"], phi0=self.model["phi0"], kappa=self.model.get("kappa", 1.0),
                    saturate_online=self.model.get("saturate_online", False), hprop=self.hprop,
                    grid=self.grid, grids=self.grids, rhobar=self.rhobar,
                    pressure_gradient=self.pressure_gradient,
                    dkk=self.dkk, dll=self.dll, rr_mm_area=self.rr_mm_area)

    def install(self, lprop) -> None:
        """Configure a libprop-like module (ours or the reference's) the way the driver does
        (raytracer.py:38, 53-64, 76-77, 98-99, 105)."""
        lprop.HPROP_GLOBAL = self.hprop
        lprop.set_model_setup(rhs=lprop.rhs_default, **self.model)
        lprop.grid = self.grid
        lprop.grids = self.grids
        lprop.rhobar = self.rhobar
        lprop.pressure_gradient = self.pressure_gradient
        lprop.set_statics(dkk=self.dkk, dll=self.dll, rr_mm_area=self.rr_mm_area)


def _hydrostatic_rhobar(grids, rhobar0=1.2, hh=8500., boussinesq=False):
    return rhobar0 * np.ones(grids.shape) if boussinesq else rhobar0 * np.exp(-grids / hh)


def _geostrophic_pg(rhobar, phi0, uu, vv):
    ff = 2 * ROT_EARTH * np.sin(phi0)
    pg = np.empty((2, len(rhobar)))
    pg[0] = rhobar * ff * vv
    pg[1] = -rhobar * ff * uu
    return pg


def default_column() -> Scenario:
    """The reference driver's wave packet: 60 ray volumes, 101-point grid, dt = 120 s."""
    NN = 0.01
    nray = 60
    phi0 = np.deg2rad(0)
    alpha = 0.01
    grid = np.linspace(0, 100e3, 101)
    grids = .5 * (grid[:-1] + grid[1:])
    model = dict(bvf=NN, boussinesq=False, sig_rr=10000, u0=4, rr0=40000, rr1=40000, phi0=phi0,
                 kappa=1., saturate_online=False, hh=8500, rhobar0=1.2)

    k_abs = 2 * np.pi / 50e3
    direction = 90
    kk = np.ones(nray) * k_abs * np.sin(np.deg2rad(direction))
    ll = np.ones(nray) * k_abs * np.cos(np.deg2rad(direction))
    mm = np.ones(nray) * -2 * np.pi / 5e3
    lam = np.zeros(nray)
    phi = np.ones(nray) * phi0
    edges = np.linspace(0, 15000, nray + 1)
    rr = .5 * (edges[:-1] + edges[1:])
    drr = np.ones(nray) * np.diff(rr)[0]
    area = 5e-5 * drr
    dmm = area / drr
    envelope = .5 * (np.tanh((grids - model["rr0"]) / model["sig_rr"]) + 1)
    uu = model["u0"] * envelope * np.sin(grids / model["sig_rr"] * 2 * np.pi)
    vv = np.zeros(uu.shape)
    rhobar = _hydrostatic_rhobar(grids, model["rhobar0"], model["hh"], model["boussinesq"])
    pg = _geostrophic_pg(rhobar, phi0, uu, vv)
    dll = np.ones(nray) * 1e-4
    dkk = np.ones(nray) * 1e-4

    f0 = 2 * ROT_EARTH * np.sin(phi0)
    rhobar_ray = np.interp(rr, grids, rhobar)
    omh = np.sqrt((NN ** 2 * (kk ** 2 + ll ** 2) + f0 ** 2 * mm ** 2) / (kk ** 2 + ll ** 2 + mm ** 2))
    amplitude = alpha ** 2 * rhobar_ray / 2 * omh / mm ** 2 / (omh ** 2 - f0 ** 2) * NN ** 2
    profile = np.exp(-(rr - rr.mean()) ** 2 / 2 / 2000 ** 2)
    dens = amplitude * profile / dkk / dll / dmm

    return Scenario("default_column", 120., [dens, lam, phi, rr, drr, kk, ll, mm, dmm], uu, vv,
                    dkk, dll, area, grid, grids, rhobar, pg, model, hprop=False)


def column_ensemble(n: int, seed: int = 1234, ngrid: int = 1001, sheared: bool = False,
                    shuffled: bool = False, amplitude: float | None = None, phi0: float = 0.0,
                    ztop_rays: float = 60e3) -> Scenario:
    """SURVEY.md section 8(d) synthetic ensemble (C2; ``sheared`` -> the C3 wind on constant N).

    rr are the midpoints of linspace(0, 60 km, n+1) (sorted), drr ~ U(50,300) m,
    mm = -2pi/U(1,10) km, |k_h| = 2pi/U(20,200) km with a uniform azimuth,
    dmm = 1e-4 |mm|, dkk = dll = 1e-4, dens = 1 (or scaled by ``amplitude``).
    """
    rng = np.random.default_rng(seed)
    NN = 0.01
    grid = np.linspace(0, 100e3, ngrid)
    grids = .5 * (grid[:-1] + grid[1:])
    edges = np.linspace(0, ztop_rays, n + 1)
    rr = .5 * (edges[:-1] + edges[1:])
    drr = rng.uniform(50., 300., n)
    mm = -2 * np.pi / rng.uniform(1e3, 10e3, n)
    kh = 2 * np.pi / rng.uniform(20e3, 200e3, n)
    theta = rng.uniform(0., 2 * np.pi, n)
    kk = kh * np.sin(theta)
    ll = kh * np.cos(theta)
    dmm = 1e-4 * np.abs(mm)
    dkk = np.full(n, 1e-4)
    dll = np.full(n, 1e-4)
    area = dmm * drr
    lam = np.zeros(n)
    phi = np.full(n, float(phi0))
    dens = np.ones(n)
    model = dict(bvf=NN, phi0=float(phi0), kappa=1., saturate_online=False)
    rhobar = _hydrostatic_rhobar(grids)
    if sheared:
        jet = 40. * .5 * (np.tanh((grids - 30e3) / 10e3) + 1)
        uu = jet + 5. * np.sin(2 * np.pi * grids / 10e3)
        vv = 2. * np.cos(2 * np.pi * grids / 15e3)
    else:
        uu = np.zeros(grids.shape)
        vv = np.zeros(grids.shape)
    if amplitude is not None:
        # wave-action density that makes the deposit feed back on the wind (raytracer.py:115-117 scaling)
        f0 = 2 * ROT_EARTH * np.sin(phi0)
        omh = np.sqrt((NN ** 2 * (kk ** 2 + ll ** 2) + f0 ** 2 * mm ** 2) / (kk ** 2 + ll ** 2 + mm ** 2))
        rho_ray = np.interp(rr, grids, rhobar)
        # ... shared among the ray volumes that overlap one grid cell, so that the ensemble as a whole (not each
        # of its n members) carries the amplitude `amplitude` relative to static instability
        per_cell = max(1.0, n * float(np.mean(drr)) / ztop_rays)
        dens = amplitude ** 2 * rho_ray / 2 * omh / mm ** 2 / (omh ** 2 - f0 ** 2) * NN ** 2 / dkk / dll / dmm / per_cell
    pg = _geostrophic_pg(rhobar, phi0, uu, vv)
    state = [dens, lam, phi, rr, drr, kk, ll, mm, dmm]
    stat = [dkk, dll, area]
    if shuffled:
        perm = rng.permutation(n)
        state = [a[perm] for a in state]
        stat = [a[perm] for a in stat]
    name = "column_%s%s_n%d" % ("sheared" if sheared else "constN", "_shuffled" if shuffled else "", n)
    return Scenario(name, 120., state, uu, vv, stat[0], stat[1], stat[2], grid, grids, rhobar, pg, model, hprop=False)


def n_profile(grids):
    """N(z) of BASELINE configs[2-3]: N^2 = 1e-4 (1 + 3 * (1 + tanh((z - 15 km) / 3 km)) / 2),
    troposphere -> stratosphere, as the array of N on `grids` that the extension takes in model_config['bvf']."""
    return np.sqrt(1e-4 * (1 + 3 * .5 * (1 + np.tanh((np.asarray(grids) - 15e3) / 3e3))))


def nz_sheared_ensemble(n: int, seed: int = 1234, ngrid: int = 1001, amplitude: float = 0.01, shuffled: bool = False) -> Scenario:
    """BASELINE.json configs[2] (and, sharded over 8 GPUs, configs[3]): the rays of column_ensemble in a sheared wind
    (tanh jet of 40 m/s centred at 30 km plus a 5 m/s sine of 10 km wavelength) under the N^2(z) profile above, with
    wave-action amplitudes that make the deposited flux feed back on the wind (raytracer.py:115-117 scaling)."""
    sc = column_ensemble(n, seed=seed, ngrid=ngrid, sheared=True, amplitude=amplitude, shuffled=shuffled)
    sc.model = dict(sc.model, bvf=n_profile(sc.grids))
    sc.name = "nz_sheared_n%d" % n
    return sc


M_CRIT_STRESS = 2 * np.pi / 150.0     # |m| cut-off of the configs[4] stress scenario: vertical wavelength 150 m


def critical_level_ensemble(n: int, seed: int = 4321, ngrid: int = 1001, stress: bool = False) -> Scenario:
    """BASELINE.json configs[4] (constant-N variant): a jet U(z) = 20 m/s * exp(-(z - 30 km)^2 / (2 (5 km)^2)).
    Rays whose horizontal wavenumber is parallel to the jet are refracted towards a critical level
    (|m| grows without bound, c_g -> 0: they pile up below the jet); antiparallel rays run out of the top.
    Half of the rays have either sign of k; l = 0.  Ray deletion (RayEnsemble.compact) removes both kinds.

    stress=True: the variant in which deletion HAPPENS within tens of steps -- a sharper, stronger jet (40 m/s,
    sigma 2 km), rays launched on its lower flank (24-29 km) and, for the exit through the top, a tenth of them at
    98-99.7 km; vertical wavelengths of 180-1000 m and horizontal ones of 5-30 km; use with M_CRIT_STRESS."""
    rng = np.random.default_rng(seed)
    NN = 0.01
    grid = np.linspace(0, 100e3, ngrid)
    grids = .5 * (grid[:-1] + grid[1:])
    if stress:
        n_top = n // 10
        lo = np.linspace(24e3, 29e3, n - n_top + 1)
        hi = np.linspace(98e3, 99.7e3, n_top + 1)
        rr = np.concatenate((.5 * (lo[:-1] + lo[1:]), .5 * (hi[:-1] + hi[1:])))
        drr = rng.uniform(50., 300., n)
        mm = -2 * np.pi / rng.uniform(180., 1e3, n)
        kk = 2 * np.pi / rng.uniform(5e3, 30e3, n) * rng.choice([-1., 1.], n)
        jet = (40., 30e3, 2e3)
    else:
        edges = np.linspace(0, 25e3, n + 1)
        rr = .5 * (edges[:-1] + edges[1:])
        drr = rng.uniform(50., 300., n)
        mm = -2 * np.pi / rng.uniform(1e3, 5e3, n)
        kk = 2 * np.pi / rng.uniform(20e3, 100e3, n) * rng.choice([-1., 1.], n)
        jet = (20., 30e3, 5e3)
    ll = np.zeros(n)
    dmm = 1e-4 * np.abs(mm)
    dkk = np.full(n, 1e-4)
    dll = np.full(n, 1e-4)
    area = dmm * drr
    model = dict(bvf=NN, phi0=0.0, kappa=1., saturate_online=False)
    rhobar = _hydrostatic_rhobar(grids)
    uu = jet[0] * np.exp(-(grids - jet[1]) ** 2 / 2 / jet[2] ** 2)
    vv = np.zeros(grids.shape)
    omh = np.sqrt(NN ** 2 * (kk ** 2 + ll ** 2) / (kk ** 2 + ll ** 2 + mm ** 2))
    rho_ray = np.interp(rr, grids, rhobar)
    per_cell = max(1.0, n * float(np.mean(drr)) / (5e3 if stress else 25e3))
    dens = 0.1 ** 2 * rho_ray / 2 * omh / mm ** 2 / omh ** 2 * NN ** 2 / dkk / dll / dmm / per_cell
    pg = _geostrophic_pg(rhobar, 0.0, uu, vv)
    return Scenario("critical_level%s_n%d" % ("_stress" if stress else "", n), 120., [dens, np.zeros(n), np.zeros(n), rr, drr, kk, ll, mm, dmm], uu, vv,
                    dkk, dll, area, grid, grids, rhobar, pg, model, hprop=False)
