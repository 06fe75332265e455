"""CPU oracle for the python-msgwam hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product
(``python-msgwam_b200/``) never does.  Parity status: pinned against the unmodified
Python reference (see ``oracle/msgwam_oracle.c`` header and ``tests/golden/``).

The arithmetic lives in ``msgwam_oracle.c``; this module is the ctypes binding plus
the derivation of the Python-float scalars exactly as the reference derives them
(``/root/reference/lib/libprop.py`` line numbers are given as L:nnn).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmsgwam_oracle.so")

RAD_EARTH = 6378e3       # L:3
ROT_EARTH = 7.2921e-5    # L:4

_dp = ctypes.POINTER(ctypes.c_double)


class _Params(ctypes.Structure):
    _fields_ = [
        ("n2", ctypes.c_double), ("two_rot", ctypes.c_double), ("rad_earth", ctypes.c_double),
        ("c8rot2", ctypes.c_double), ("f0", ctypes.c_double), ("f0sq", ctypes.c_double),
        ("k2half", ctypes.c_double), ("dz_grid", ctypes.c_double), ("dz_grids", ctypes.c_double),
        ("ngrid", ctypes.c_int32), ("hprop", ctypes.c_int32), ("saturate_online", ctypes.c_int32),
        ("nthreads", ctypes.c_int32),
        ("bvf_prof", ctypes.c_void_p), ("bvf_grids", ctypes.c_void_p),     # extension: N(z) profile on grids
    ]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -ffp-contract=off).  Returns the .so path."""
    src = os.path.join(_HERE, "msgwam_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _ptr_array(arrs):
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    return keep, (_dp * len(keep))(*[k.ctypes.data_as(_dp) for k in keep])


class Oracle:
    """Reference arithmetic for one configuration snapshot.

    cfg keys: bvf, phi0, kappa, saturate_online, hprop, grid, grids, rhobar,
    pressure_gradient (2,G), dkk, dll, rr_mm_area (per-ray statics).
    """

    def __init__(self, cfg: dict, nthreads: int = 1):
        self.lib = _load()
        self.cfg = cfg
        grid = np.ascontiguousarray(cfg["grid"], dtype=np.float64)
        grids = np.ascontiguousarray(cfg["grids"], dtype=np.float64)
        self.grid, self.grids = grid, grids
        G = len(grids)
        self.G = G
        self.rhobar = np.ascontiguousarray(np.broadcast_to(np.asarray(cfg.get("rhobar", 1.0), dtype=np.float64), (G,)))
        pg = cfg.get("pressure_gradient", None)
        self.pg = np.zeros((2, G)) if pg is None else np.ascontiguousarray(pg, dtype=np.float64)
        phi0 = cfg["phi0"]
        bvf = cfg["bvf"]
        kappa = cfg.get("kappa", 1.0)
        f0 = 2 * ROT_EARTH * np.sin(phi0)                  # L:535, 589 (numpy scalar)
        p = _Params()
        # extension (not in the reference): bvf may be an array of N on `grids`
        self.bvf_prof = None
        if np.ndim(bvf) > 0:
            self.bvf_prof = np.ascontiguousarray(bvf, dtype=np.float64)
            assert self.bvf_prof.shape == (G,), "bvf profile must live on grids"
            p.bvf_prof = self.bvf_prof.ctypes.data
            p.bvf_grids = grids.ctypes.data
            bvf = float("nan")                             # the scalar must never be used then
        p.n2 = bvf ** 2                                    # L:383
        p.two_rot = 2 * ROT_EARTH                          # L:382
        p.rad_earth = RAD_EARTH
        p.c8rot2 = 8 * ROT_EARTH ** 2                      # L:491
        p.f0 = float(f0)
        p.f0sq = float(f0 ** 2)                            # L:383 / L:601 with scalar phi0
        p.k2half = kappa ** 2 * .5                         # L:601
        p.dz_grid = float(np.diff(grid[:2])[0])            # L:349, 662
        p.dz_grids = float(np.diff(grids[:2])[0])          # L:123 with grid := grids
        p.ngrid = len(grid)
        p.hprop = int(bool(cfg.get("hprop", False)))
        p.saturate_online = int(bool(cfg.get("saturate_online", False)))
        p.nthreads = int(nthreads)
        self.p = p

    # -- statics ------------------------------------------------------------
    def _statics(self, n):
        c = self.cfg
        return [np.ascontiguousarray(np.broadcast_to(np.asarray(c[k], dtype=np.float64), (n,)))
                for k in ("dkk", "dll", "rr_mm_area")]

    # -- point functions ------------------------------------------------------
    def interp(self, x, xp, fp):
        x, xc = _c(x); xp, xpc = _c(xp); fp, fpc = _c(fp)
        out = np.empty_like(x)
        self.lib.orc_interp(xc, ctypes.c_long(x.size), xpc, fpc, ctypes.c_long(xp.size), out.ctypes.data_as(_dp))
        return out

    def _rr(self, rr, n):
        """Position argument: only the N(z) extension looks at it."""
        if self.bvf_prof is None or rr is None:
            assert self.bvf_prof is None, "a bvf profile needs the position argument"
            return None, None
        return _c(np.broadcast_to(np.asarray(rr, dtype=np.float64), (n,)))

    def omega(self, kk, ll, mm, phi, rr=None):
        kk, a = _c(kk); ll, b = _c(ll); mm, c = _c(mm)
        out = np.empty_like(kk)
        keep, r = self._rr(rr, kk.size)
        if np.ndim(phi) == 0:
            f = 2 * ROT_EARTH * np.sin(phi)
            self.lib.orc_omega_scalar_phi(ctypes.c_long(kk.size), a, b, c, ctypes.c_double(float(f ** 2)), r,
                                          ctypes.byref(self.p), out.ctypes.data_as(_dp))
        else:
            phi, d = _c(phi)
            self.lib.orc_omega(ctypes.c_long(kk.size), a, b, c, d, r, ctypes.byref(self.p), out.ctypes.data_as(_dp))
        return out

    def cg_rr(self, kk, ll, mm, lam, phi, rr):
        kk, a = _c(kk); ll, b = _c(ll); mm, c = _c(mm); phi, d = _c(phi)
        out = np.empty_like(kk)
        keep, r = self._rr(rr, kk.size)
        self.lib.orc_cg_rr(ctypes.c_long(kk.size), a, b, c, d, r, ctypes.byref(self.p), out.ctypes.data_as(_dp))
        return out

    def wave_projection(self, dens, lam, phi, rr_low, rr_up, kk, ll, mm_low, mm_up, dkk, dll, dmm, grid, var=0):
        n = np.size(dens)
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
                for a in (dens, phi, rr_low, rr_up, kk, ll, mm_low, mm_up, dkk, dll, dmm)]
        grid, gc = _c(grid)
        ng = grid.size
        shape = {0: (2, ng - 1), 1: (ng - 1,), 2: (ng - 1,), 3: (ng,), 4: (2, ng)}[var]
        out = np.zeros(shape)
        self.lib.orc_wave_projection(ctypes.c_int(var), ctypes.c_long(n),
                                     *[a.ctypes.data_as(_dp) for a in arrs],
                                     gc, ctypes.c_long(ng), ctypes.byref(self.p), out.ctypes.data_as(_dp))
        return out

    def saturation(self, dt, dens, rr_center, rr_center_st, drr, drr_st, kk, ll, mm_center, mm_center_st,
                   direct=False):
        n = np.size(dens)
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
                for a in (dens, rr_center, rr_center_st, drr, drr_st, kk, ll, mm_center, mm_center_st)]
        st = self._statics(n)
        out = np.empty(n)
        self.lib.orc_saturation(ctypes.c_double(dt), ctypes.c_long(n),
                                *[a.ctypes.data_as(_dp) for a in arrs],
                                *[a.ctypes.data_as(_dp) for a in st],
                                self.grids.ctypes.data_as(_dp), self.rhobar.ctypes.data_as(_dp),
                                ctypes.byref(self.p), ctypes.c_int(int(direct)), out.ctypes.data_as(_dp))
        return out

    # -- rhs / integrator -----------------------------------------------------
    def rhs_default(self, dt, var_in, return_projection=False):
        n = np.size(var_in[0])
        keep, sp = _ptr_array(var_in[:9])
        uu, uc = _c(var_in[9]); vv, vc = _c(var_in[10])
        st = self._statics(n)
        stp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in st])
        tend = [np.empty(n) for _ in range(9)]
        tp = (_dp * 9)(*[a.ctypes.data_as(_dp) for a in tend])
        du = np.empty(self.G); dv = np.empty(self.G)
        proj = np.zeros((2, self.G - 1))
        self.lib.orc_rhs_default(ctypes.c_double(dt), ctypes.c_long(n), sp, uc, vc, stp,
                                 self.grid.ctypes.data_as(_dp), self.grids.ctypes.data_as(_dp),
                                 self.rhobar.ctypes.data_as(_dp), self.pg.ctypes.data_as(_dp),
                                 ctypes.byref(self.p), tp, du.ctypes.data_as(_dp), dv.ctypes.data_as(_dp),
                                 proj.ctypes.data_as(_dp))
        out = np.empty(11, dtype=object)
        for i in range(9):
            out[i] = tend[i]
        out[9], out[10] = du, dv
        return (out, proj) if return_projection else out

    def RK3(self, dt, var):
        n = np.size(var[0])
        state = [np.array(a, dtype=np.float64, copy=True) for a in var[:9]]
        uu = np.array(var[9], dtype=np.float64, copy=True)
        vv = np.array(var[10], dtype=np.float64, copy=True)
        sp = (_dp * 9)(*[a.ctypes.data_as(_dp) for a in state])
        st = self._statics(n)
        stp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in st])
        self.lib.orc_rk3(ctypes.c_double(dt), ctypes.c_long(n), sp, uu.ctypes.data_as(_dp), vv.ctypes.data_as(_dp),
                         stp, self.grid.ctypes.data_as(_dp), self.grids.ctypes.data_as(_dp),
                         self.rhobar.ctypes.data_as(_dp), self.pg.ctypes.data_as(_dp), ctypes.byref(self.p))
        out = np.empty(11, dtype=object)
        for i in range(9):
            out[i] = state[i]
        out[9], out[10] = uu, vv
        return out


    def RK3_frozen(self, dt, var):
        """ORACLE OF AN EXTENSION (frozen-background mode "M2"; no such stepper in the reference): composed
        only of restated reference functions, the way SURVEY.md 7.3-1 prescribes --
          rays : the reference's RK3 (L:693-698) with model_config['rhs'] (L:691) = rhs_default whose du_st, dv_st are
                 replaced by zeros, so that uu, vv stay frozen over the step;
          flow : once per step  pm_flux[:, 1:-1] = wave_projection(new rays, var=0); edge copies; diff / dz (L:653-663);
                 uu += dt * du_dt(vv, grad[0]); vv += dt * dv_dt(uu, grad[1])  (L:523-558).
        Pinned against exactly that composition of the live Python reference (tests/test_oracle_vs_reference.py) and the
        golden fixture it produced (tests/golden/frozen_col.npz)."""
        def rhs(v):
            t = self.rhs_default(dt, v)
            t[9] = np.zeros(self.G); t[10] = np.zeros(self.G)
            return t
        var = np.array(list(var), dtype=object)
        qq = dt * rhs(var)                                   # L:693-698, on the object array like the reference
        var = var + qq / 3
        qq = dt * rhs(var) - 5 / 9 * qq
        var = var + 15 / 16 * qq
        qq = dt * rhs(var) - 153 / 128 * qq
        var = var + 8 / 15 * qq
        dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv = var
        n = np.size(rr)
        dkk, dll, _ = self._statics(n)
        pm_flux = np.zeros((2, self.G + 1))
        pm_flux[:, 1:-1] = self.wave_projection(dens, lam, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                                                dkk, dll, dmm, self.grids)
        pm_flux[:, 0] = pm_flux[:, 1]
        pm_flux[:, -1] = pm_flux[:, -2]
        dz = np.diff(self.grid[:2])[0]
        grad = (pm_flux[:, 1:] - pm_flux[:, :-1]) / dz
        ff = 2 * ROT_EARTH * np.sin(self.cfg["phi0"])
        du = ff * vv - self.rhobar ** -1 * (self.pg[0] + grad[0])          # du_dt, L:537
        dv = -ff * uu - self.rhobar ** -1 * (self.pg[1] + grad[1])         # dv_dt, L:556
        out = np.empty(11, dtype=object)
        for i in range(9):
            out[i] = var[i]
        out[9], out[10] = uu + dt * du, vv + dt * dv
        return out


def max_threads() -> int:
    return int(_load().orc_max_threads())
