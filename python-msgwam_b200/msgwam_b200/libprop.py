"""Drop-in for the reference's ``lib/libprop.py`` -- same module namespace, B200 kernels underneath.

The reference's plug-in boundary is this module's namespace (``import lib.libprop as lprop``,
/root/reference/raytracer.py:2): module-level mutable globals (``HPROP_GLOBAL``, ``grid``, ``grids``,
``rhobar``, ``pressure_gradient``, ``model_config``, ``statics``), setter functions, and the
propagation functions.  All of them keep their names, signatures, argument meaning and error
behaviour (no validation; a missing static raises ``KeyError`` like L:630-631).

What runs where
  * ``RK3`` / ``rhs_default`` / ``wave_projection`` / ``saturation`` and the point functions
    (``omega``, ``cg_*``, ``d?_dt``, ``gradients``, ``du_dt``, ``dv_dt``) run on the GPU through the C ABI
    (``include/msgwam_b200.h``).  Inputs may be numpy arrays (copied to the device and back, results
    are numpy) or torch CUDA float64 tensors (zero-copy, results are torch tensors).
  * the one-off setup helpers (``set_*``, ``velocities_*``) act on <= 1e3 grid points once per run and
    stay host numpy code (out of scope for acceleration).
There is no CPU fallback for the first group: without the CUDA library / a device they raise.

L:nnn = /root/reference/lib/libprop.py line.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, lib

RAD_EARTH = 6378e3          # L:3
ROT_EARTH = 7.2921e-5       # L:4
HPROP_GLOBAL = True         # L:5
pressure_gradient = 0       # L:6
grid = None                 # L:7
grids = None                # L:8
rhobar = 1                  # L:9
model_config = {}           # L:10
statics = {}                # L:11

_vp = ctypes.c_void_p


# ------------------------------------------------------------------------------------------------
# configuration (host; L:14-89)
# ------------------------------------------------------------------------------------------------
def set_statics(**kwargs):
    """Store per-run constants (ray-volume widths dkk, dll and the r-m area) by name.  L:14-27"""
    global _statics_on_device, _statics_frozen
    statics.update(kwargs)
    _statics_on_device = None           # device copies of dkk, dll are refreshed at the next RK3
    _statics_frozen = False


def freeze_statics(frozen=True):
    """Extension (not in the reference): promise that statics['dkk'] and statics['dll'] keep their values until the
    next set_statics() call.  The reference reads the dict at every rhs evaluation (L:630-632), so by default the
    host-buffer RK3 uploads both arrays with every call; after freeze_statics() their device copies are reused as
    long as the very same array objects stay in the dict (16 bytes per ray less over PCIe per step)."""
    global _statics_on_device, _statics_frozen
    _statics_frozen = bool(frozen)
    _statics_on_device = None


def set_model_setup(**kwargs):
    """Add/overwrite entries of the global model configuration.  L:30-44"""
    model_config.update(kwargs)


def get_model_setup():
    """The model configuration dictionary.  L:85-89"""
    return model_config


def set_hydrostatics():
    """Hydrostatic background density on the staggered grid.  L:47-62"""
    global rhobar
    scale = model_config['rhobar0']
    if model_config['boussinesq']:
        rhobar = scale * np.ones(grids.shape)
    else:
        rhobar = scale * np.exp(-grids / model_config['hh'])


def set_pressure_gradient(uu, vv):
    """Pressure gradient that balances (uu, vv) geostrophically.  L:65-82"""
    global pressure_gradient
    ff = 2 * ROT_EARTH * np.sin(model_config['phi0'])
    balanced = np.empty((2, len(grids)))
    balanced[0] = rhobar * ff * vv
    balanced[1] = - rhobar * ff * uu
    pressure_gradient = balanced


# ------------------------------------------------------------------------------------------------
# background wind generators (host; L:224-325)
# ------------------------------------------------------------------------------------------------
def _tanh_step(rr):
    return (np.tanh((rr - model_config['rr0']) / model_config['sig_rr']) + 1) * 0.5


def velocities_tanh(lam, phi, rr):
    """(4,3)+shape array whose [0,0] is a jet: Gaussian in latitude, tanh in height.  L:224-250"""
    lat = np.exp(-(phi - model_config['phi0'])**2 / 2 / model_config['sig_phi']**2)
    field = np.zeros((4, 3) + lam.shape)
    field[0] = model_config['u0'] * (lat * _tanh_step(rr))
    return field


def velocities_tanh_homogeneous(rr):
    """tanh jet in height.  L:253-273"""
    return model_config['u0'] * _tanh_step(rr)


def velocities_gauss_homogeneous(rr):
    """Gaussian jet in height.  L:276-303 (the reference's cut-off mask `<= & >=` is empty; kept)"""
    uu = model_config['u0'] * np.exp(-(rr - model_config['rr0'])**2 / 2 / model_config['sig_rr']**2)
    lo, hi = model_config['rr0'] - 3 * model_config['sig_rr'], model_config['rr0'] + 3 * model_config['sig_rr']
    uu[(rr <= lo) & (rr >= hi)] = 0.
    return uu


def velocities_sine_homogeneous(rr):
    """tanh envelope times a sine in height.  L:306-325"""
    envelope = .5 * (np.tanh((rr - model_config['rr0']) / model_config['sig_rr']) + 1)
    return model_config['u0'] * envelope * np.sin(rr / model_config['sig_rr'] * 2 * np.pi)


# ------------------------------------------------------------------------------------------------
# device plumbing
# ------------------------------------------------------------------------------------------------
def _engine():
    from ._engine import Engine
    return Engine.get()


def _params(dt=0.0):
    """POD snapshot of the globals the kernels read (they are mutable between calls: the driver pokes
    HPROP_GLOBAL, grid, grids directly, R:38, 76-77)."""
    return _cabi.snapshot_params(
        dt, bvf=model_config['bvf'], phi0=model_config['phi0'], kappa=model_config.get('kappa', 1.0),
        saturate_online=model_config.get('saturate_online', False), hprop=HPROP_GLOBAL,
        grid=grid, grids=grids, rot_earth=ROT_EARTH, rad_earth=RAD_EARTH)


def _params_nogrid(dt=0.0):
    """Snapshot for functions that do not need the grid (omega, cg_rr): grid entries are dummies."""
    g = grid if grid is not None else np.array([0., 1., 2., 3.])
    gs = grids if grids is not None else np.array([.5, 1.5, 2.5])
    return _cabi.snapshot_params(
        dt, bvf=model_config['bvf'], phi0=model_config.get('phi0', 0.0), kappa=model_config.get('kappa', 1.0),
        saturate_online=model_config.get('saturate_online', False), hprop=HPROP_GLOBAL,
        grid=g, grids=gs, rot_earth=ROT_EARTH, rad_earth=RAD_EARTH)


def _any_dev(eng, *xs):
    return any(eng.is_dev(x) for x in xs)


def _out(eng, t, like_dev):
    return t if like_dev else t.cpu().numpy()


def _size(*xs):
    """Number of rays: the size of the first array-like argument (scalars broadcast to it)."""
    for x in xs:
        if hasattr(x, "numel") and x.ndim > 0:
            return int(x.numel())
        if not hasattr(x, "numel") and np.ndim(x) > 0:
            return int(np.size(x))
    return 1


def _bvf_profile():
    """EXTENSION (not in the reference): model_config['bvf'] may be an array of N on
    `grids` instead of the reference's scalar.  Returns that array, or None for the reference's behaviour."""
    b = model_config['bvf']
    return b if np.ndim(b) > 0 else None


def _grid_devs(eng):
    if np.ndim(pressure_gradient) == 0:
        # the reference indexes pressure_gradient[0] (L:537): a scalar raises TypeError there as well
        raise TypeError("pressure_gradient is not set (call set_pressure_gradient first)")
    return eng.grid_on_device(grid, grids, rhobar, pressure_gradient, _bvf_profile())


def _pack11(slots):
    out = np.empty(11, dtype=object)
    for i, s in enumerate(slots):
        out[i] = s
    return out


# ------------------------------------------------------------------------------------------------
# point functions (L:328-520)
# ------------------------------------------------------------------------------------------------
def _pointwise(op, kk, ll, mm, phi, rr, uu, vv, f=0.0, f2=0.0, need_grid=False, nout=1):
    eng = _engine()
    like_dev = _any_dev(eng, kk, ll, mm, phi, rr)
    n = _size(*[x for x in (kk, ll, mm, phi, rr) if x is not None])
    d = lambda x: None if x is None else eng.dev(x, n)
    tk, tl, tm, tp, tr = d(kk), d(ll), d(mm), d(phi), d(rr)
    prof = _bvf_profile()
    if need_grid or prof is not None:
        p = _params()
        G = p.G
        gd = eng.grid_on_device(grid, grids, rhobar, pressure_gradient if np.ndim(pressure_gradient) else np.zeros((2, G)),
                                prof)
        g = eng.grid_struct(gd)
        tu, tv = (eng.dev(uu, G), eng.dev(vv, G)) if need_grid else (None, None)
    else:
        p = _params_nogrid()
        g, tu, tv = None, None, None
    out = eng.empty(nout, n) if nout > 1 else eng.empty(n)
    check(lib.msgwam_pointwise(op, p, n, eng.ptr(tk), eng.ptr(tl), eng.ptr(tm), eng.ptr(tp), eng.ptr(tr),
                               float(f), float(f2), g, eng.ptr(tu), eng.ptr(tv), eng.ptr(out), eng.stream),
          "msgwam_pointwise")
    eng.launches += 1
    return _out(eng, out, like_dev)


def _position(rr):
    """The height argument matters only to the N(z) extension; with the reference's scalar bvf it is dropped
    (L:434-448 ignores it)."""
    if _bvf_profile() is None:
        return None
    if rr is None:
        raise TypeError("model_config['bvf'] is a profile: omega() needs the height argument rr")
    return rr


def omega(kk, ll, mm, phi, rr=None):
    """Intrinsic frequency sqrt((N^2 (k^2+l^2) + f^2 m^2) / |k|^2).  L:369-383
    (`rr` is only read by the N(z) extension; the reference signature has no such argument.)"""
    if np.ndim(phi) == 0 and not hasattr(phi, "is_cuda"):
        f = 2 * ROT_EARTH * np.sin(phi)            # numpy scalar, squared with the scalar power (L:382-383)
        return _pointwise(_cabi.OP_OMEGA_F, kk, ll, mm, None, _position(rr), None, None, f=f, f2=f ** 2)
    return _pointwise(_cabi.OP_OMEGA, kk, ll, mm, phi, _position(rr), None, None)


def cg_rr(kk, ll, mm, lam, phi, rr):
    """Vertical group velocity -m (om^2 - f^2) / om / |k|^2 (lam and rr are ignored, as in L:434-448;
    the N(z) extension evaluates N at rr)."""
    return _pointwise(_cabi.OP_CG_RR, kk, ll, mm, phi, _position(rr), None, None)


def cg_lambda(kk, ll, mm, lam, phi, rr, uu, vv):
    """Zonal group velocity incl. the interpolated wind; zeros when HPROP_GLOBAL is off.  L:386-407"""
    return _pointwise(_cabi.OP_CG_LAMBDA, kk, ll, mm, phi, rr, uu, vv, need_grid=True)


def cg_phi(kk, ll, mm, lam, phi, rr, uu, vv):
    """Meridional group velocity; zeros when HPROP_GLOBAL is off.  L:410-431"""
    return _pointwise(_cabi.OP_CG_PHI, kk, ll, mm, phi, rr, uu, vv, need_grid=True)


def dk_dt(kk, ll, mm, lam, phi, rr, uu, vv):
    """Refraction of the zonal wavenumber (spherical metric terms); zeros when HPROP is off.  L:451-471"""
    return _pointwise(_cabi.OP_DK_DT, kk, ll, mm, phi, rr, uu, vv, need_grid=True)


def dl_dt(kk, ll, mm, lam, phi, rr, uu, vv):
    """Refraction of the meridional wavenumber; zeros when HPROP is off.  L:474-499"""
    return _pointwise(_cabi.OP_DL_DT, kk, ll, mm, phi, rr, uu, vv, need_grid=True)


def dm_dt(kk, ll, mm, lam, phi, rr, uu, vv):
    """Refraction of the vertical wavenumber: (k cg_lam + l cg_phi)/(R+r) - (k du/dz + l dv/dz).  L:502-520"""
    return _pointwise(_cabi.OP_DM_DT, kk, ll, mm, phi, rr, uu, vv, need_grid=True)


def gradients(lam_ray, phi_ray, rr_ray, uu, vv):
    """(4,3)+shape array: [0,0]=u, [0,1]=v, [1,2]=du/dz, [2,2]=dv/dz at the ray heights.  L:328-366"""
    eng = _engine()
    like_dev = _any_dev(eng, lam_ray, rr_ray)
    four = _pointwise(_cabi.OP_GRADIENTS, None, None, None, None, rr_ray, uu, vv, need_grid=True, nout=4)
    shape = tuple(lam_ray.shape)
    if like_dev:
        out = eng.zeros(4, 3, *shape)
    else:
        out = np.zeros((4, 3) + shape)
    out[0, 0] = four[0].reshape(shape)
    out[0, 1] = four[1].reshape(shape)
    out[1, 2] = four[2].reshape(shape)
    out[2, 2] = four[3].reshape(shape)
    return out


def _grid_tend(which, wind, pm_flux_gradient):
    """f v - (pg + dF/dz)/rhobar  /  -f u - (pg + dF/dz)/rhobar on the staggered grid (L:523-558)."""
    eng = _engine()
    like_dev = _any_dev(eng, wind, pm_flux_gradient)
    G = _size(wind)
    if np.ndim(pressure_gradient) == 0:
        raise TypeError("pressure_gradient is not set (call set_pressure_gradient first)")
    rho = eng.dev(np.broadcast_to(np.asarray(rhobar, dtype=np.float64), (G,)) if not eng.is_dev(rhobar) else rhobar, G)
    pg = eng.dev(np.asarray(pressure_gradient, dtype=np.float64)[which])
    out = eng.empty(G)
    f0 = float(2 * ROT_EARTH * np.sin(model_config['phi0']))
    tw, tg = eng.dev(wind, G), eng.dev(pm_flux_gradient, G)      # keep the temporaries alive until the launch is queued
    check(lib.msgwam_mean_flow_tendency(which, f0, G, eng.ptr(tw), eng.ptr(tg),
                                        eng.ptr(rho), eng.ptr(pg), eng.ptr(out), eng.stream), "msgwam_mean_flow_tendency")
    eng.launches += 1
    return _out(eng, out, like_dev)


def du_dt(vv, pm_flux_gradient):
    """Zonal mean-flow tendency.  L:523-539"""
    return _grid_tend(0, vv, pm_flux_gradient)


def dv_dt(uu, pm_flux_gradient):
    """Meridional mean-flow tendency.  L:542-558"""
    return _grid_tend(1, uu, pm_flux_gradient)


# ------------------------------------------------------------------------------------------------
# deposition (L:92-221)
# ------------------------------------------------------------------------------------------------
def wave_projection(dens, lam, phi, rr_low, rr_up, kk, ll, mm_low, mm_up, dkk, dll, dmm, grid, var=0):
    """Project ray-volume properties onto a uniform vertical grid.

    var = 0 pseudo-momentum fluxes at cell centres (2, len(grid)-1); 1 wave-action flux; 2 wave action
    (len(grid)-1,); 3 / 4 wave-action / pseudo-momentum fluxes at the interfaces.  Bug-for-bug with
    L:123-163: indices are rr/dz on a grid assumed to start at 0, weights use abs().
    """
    eng = _engine()
    like_dev = _any_dev(eng, dens, rr_low, rr_up, kk, mm_low)
    n = _size(dens, rr_low, rr_up)
    arrs = [eng.dev(x, n) for x in (dens, phi, rr_low, rr_up, kk, ll, mm_low, mm_up, dkk, dll, dmm)]
    if eng.is_dev(grid):
        g = grid.to(eng.torch.float64).contiguous()
        g01 = g[:2].cpu().numpy()
    else:
        gh = np.ascontiguousarray(grid, dtype=np.float64)
        g = eng.dev(gh)
        g01 = gh[:2]
    ng = int(g.numel())
    dz = float(np.diff(g01)[0])                                     # L:123
    shape = {0: (2, ng - 1), 1: (ng - 1,), 2: (ng - 1,), 3: (ng,), 4: (2, ng)}[var]
    out = eng.empty(*shape)
    prof = _bvf_profile()
    if prof is None:
        p, tb, tbg = _params_nogrid(), None, None
    else:                                                           # extension: N lives on the module's grids
        p = _params()
        tb, tbg = eng.dev(prof, p.G), eng.dev(grids, p.G)
    check(lib.msgwam_wave_projection(var, p, n, *[eng.ptr(a) for a in arrs], eng.ptr(g), ng, dz, 1.0 / dz,
                                     eng.ptr(tb), eng.ptr(tbg), eng.ptr(out), eng.stream), "msgwam_wave_projection")
    eng.launches += 2
    return _out(eng, out, like_dev)


# ------------------------------------------------------------------------------------------------
# saturation (L:561-615)
# ------------------------------------------------------------------------------------------------
def saturation(dt, dens, rr_center, rr_center_st, drr, drr_st, kk, ll, mm_center, mm_center_st, direct=False):
    """Wave-action change (or, with direct=True, the clamped wave action) from the static-instability
    saturation criterion."""
    eng = _engine()
    dkk, dll, area = statics['dkk'], statics['dll'], statics['rr_mm_area']
    like_dev = _any_dev(eng, dens, rr_center, mm_center)
    n = _size(dens, rr_center)
    arrs = [eng.dev(x, n) for x in (dens, rr_center, rr_center_st, drr, drr_st, kk, ll, mm_center, mm_center_st,
                                    dkk, dll, area)]
    p = _params(dt)
    G = p.G
    gs = eng.dev(grids, G)
    rho = eng.dev(np.broadcast_to(np.asarray(rhobar, dtype=np.float64), (G,)) if not eng.is_dev(rhobar) else rhobar, G)
    out = eng.empty(n)
    prof = _bvf_profile()
    tb = None if prof is None else eng.dev(prof, G)
    check(lib.msgwam_saturation(p, n, int(bool(direct)), *[eng.ptr(a) for a in arrs], eng.ptr(gs), eng.ptr(rho),
                                eng.ptr(tb), eng.ptr(out), eng.stream), "msgwam_saturation")
    eng.launches += 1
    return _out(eng, out, like_dev)


# ------------------------------------------------------------------------------------------------
# right-hand side and integrator (L:618-700)
# ------------------------------------------------------------------------------------------------
def _statics_dev(eng, n):
    return [eng.dev(statics[k], n) for k in ('dkk', 'dll', 'rr_mm_area')]


def rhs_default(dt, var_in):
    """All eleven tendencies of the state vector [dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv]:
    ray propagation/refraction, online saturation, flux deposition and the mean-flow forcing."""
    dkk, dll = statics['dkk'], statics['dll']          # KeyError if the caller never set them (L:630-631)
    area = statics['rr_mm_area']
    saturate_online = model_config['saturate_online']  # noqa: F841  (KeyError parity, L:633)
    eng = _engine()
    like_dev = _any_dev(eng, *var_in)
    n = _size(var_in[3])
    p = _params(dt)
    state = [eng.dev(x, n) for x in var_in[:9]]
    uu, vv = eng.dev(var_in[9], p.G), eng.dev(var_in[10], p.G)
    st = [eng.dev(x, n) for x in (dkk, dll, area)]
    tend, du, dv, _ = eng.rhs_general(p, state, st, uu, vv, _grid_devs(eng))
    return _pack11([_out(eng, t, like_dev) for t in tend] + [_out(eng, du, like_dev), _out(eng, dv, like_dev)])


def _host_f64(a, n):
    """The caller's array itself when it already is contiguous float64 of length n (no copy, pinned memory
    stays pinned), else a converted copy."""
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 1 and a.shape[0] == n and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))


def _pinned_empty(eng, n):
    """Page-locked host array for results (torch's caching host allocator recycles the blocks)."""
    return eng.torch.empty(n, dtype=eng.torch.float64, pin_memory=True).numpy()


def _unchanged(a):
    """Slot whose tendency is exactly zero in column mode: the reference returns `var + qq/3` = the same
    values in a new array.  A read-only view says the same thing without moving 8 bytes per ray through
    the host memory system (set MSGWAM_COPY_UNCHANGED=1 for writable copies)."""
    if _COPY_UNCHANGED:
        return a.copy()
    v = a.view()
    v.flags.writeable = False
    return v


_COPY_UNCHANGED = bool(int(__import__("os").environ.get("MSGWAM_COPY_UNCHANGED", "0")))
_statics_on_device = None          # (id/pointer key of dkk, dll, n, stage tensor id) of the last upload
_statics_frozen = False            # freeze_statics(): the caller promises not to edit dkk, dll in place


def _rk3_numpy_column(eng, p, var, prof=None):
    """Host-buffer path: one C-ABI call copies the step's inputs in, steps, copies rr, mm, uu, vv out
    (prof: the N(z) extension's profile on grids; then drr, dmm change as well and come back too)."""
    global _statics_on_device
    n = _size(var[3])
    G = p.G
    host = [_host_f64(a, n) for a in var[:9]]
    uu = np.ascontiguousarray(var[9], dtype=np.float64)
    vv = np.ascontiguousarray(var[10], dtype=np.float64)
    dkk = _host_f64(statics['dkk'], n)
    dll = _host_f64(statics['dll'], n)
    statics['rr_mm_area']                                           # KeyError parity with L:632
    if np.ndim(pressure_gradient) == 0:
        raise TypeError("pressure_gradient is not set (call set_pressure_gradient first)")
    g = np.ascontiguousarray(grid, dtype=np.float64)
    gs = np.ascontiguousarray(grids, dtype=np.float64)
    rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rhobar, dtype=np.float64), (G,)))
    pg = np.ascontiguousarray(pressure_gradient, dtype=np.float64)
    rr_new, mm_new = _pinned_empty(eng, n), _pinned_empty(eng, n)
    uu_new, vv_new = np.empty(G), np.empty(G)
    hp = (_vp * 9)(*[a.ctypes.data for a in host])
    stage = eng.host_stage(n, G, prof is not None)
    work = eng.column_work(G)
    cp = lambda a: _vp(a.ctypes.data)
    # per-run statics (L:722-726): uploaded with every call (the reference reads the dict at every evaluation, and an
    # in-place edit of statics['dkk'] must take effect) unless the caller froze them (freeze_statics) and keeps
    # passing the very same arrays (same object, same buffer); set_statics() or a new array re-uploads them
    key = (id(statics['dkk']), id(statics['dll']), dkk.ctypes.data, dll.ctypes.data, n, stage.data_ptr())
    reuse = _statics_frozen and _statics_on_device == key and not _COPY_UNCHANGED
    if prof is not None:
        bv = np.ascontiguousarray(prof, dtype=np.float64).reshape(G)
        drr_new, dmm_new = _pinned_empty(eng, n), _pinned_empty(eng, n)
        check(lib.msgwam_rk3_column_nz_host(p, n, hp, _vp(0) if reuse else cp(dkk), _vp(0) if reuse else cp(dll), cp(uu), cp(vv),
                                            cp(g), cp(gs), cp(rho), cp(pg), cp(bv),
                                            cp(rr_new), cp(drr_new), cp(mm_new), cp(dmm_new), cp(uu_new), cp(vv_new),
                                            eng.ptr(stage), eng.ptr(work), eng.stream), "msgwam_rk3_column_nz_host")
        _statics_on_device = key
        eng.launches += 3
        return _pack11([_unchanged(host[0]), _unchanged(host[1]), _unchanged(host[2]), rr_new, drr_new,
                        _unchanged(host[5]), _unchanged(host[6]), mm_new, dmm_new, uu_new, vv_new])
    check(lib.msgwam_rk3_column_host(p, n, hp, _vp(0) if reuse else cp(dkk), _vp(0) if reuse else cp(dll), cp(uu), cp(vv),
                                     cp(g), cp(gs), cp(rho), cp(pg),
                                     cp(rr_new), cp(mm_new), cp(uu_new), cp(vv_new), eng.ptr(stage), eng.ptr(work),
                                     eng.stream), "msgwam_rk3_column_host")
    _statics_on_device = key
    eng.launches += 3
    return _pack11([_unchanged(host[0]), _unchanged(host[1]), _unchanged(host[2]), rr_new, _unchanged(host[4]),
                    _unchanged(host[5]), _unchanged(host[6]), mm_new, _unchanged(host[8]), uu_new, vv_new])


def RK3(dt, var):
    """Advance the 11-slot state vector by dt with the 3-stage low-storage Runge-Kutta scheme of L:680-700,
    right-hand side = model_config['rhs'].

    With the stock ``rhs_default`` the whole step runs in fused CUDA kernels (two sweeps over the rays,
    see csrc/column_step.cu) when HPROP_GLOBAL and saturate_online are off, and stage by stage on the
    device otherwise.  Any other callable plugged into model_config['rhs'] is honoured with the generic
    low-storage update on whatever array type it returns.
    """
    rhs_ = model_config['rhs']
    if rhs_ is not rhs_default:
        qq = dt * rhs_(dt, var)
        var = var + qq / 3
        qq = dt * rhs_(dt, var) - 5 / 9 * qq
        var = var + 15 / 16 * qq
        qq = dt * rhs_(dt, var) - 153 / 128 * qq
        return var + 8 / 15 * qq

    statics['dkk'], statics['dll']                                  # KeyError parity with L:630-631
    eng = _engine()
    p = _params(dt)
    like_dev = _any_dev(eng, *var)
    prof = _bvf_profile()
    column = not p.hprop and not p.saturate_online and prof is None and p.G <= eng.column_max_levels()
    column_nz = not p.hprop and not p.saturate_online and prof is not None and p.G <= eng.column_nz_max_levels()
    if column and not like_dev:
        return _rk3_numpy_column(eng, p, var)
    if column_nz and not like_dev:
        return _rk3_numpy_column(eng, p, var, prof)
    n = _size(var[3])
    state = [eng.dev(x, n) for x in var[:9]]
    uu, vv = eng.dev(var[9], p.G), eng.dev(var[10], p.G)
    gd = _grid_devs(eng)
    if column_nz:
        # N(z) extension: fused column step with the profile (rr, drr, mm, dmm evolve)
        dkk, dll = eng.dev(statics['dkk'], n), eng.dev(statics['dll'], n)
        statics['rr_mm_area']
        rr_new, drr_new, mm_new, dmm_new, uu_new, vv_new = eng.column_step_nz(p, state, dkk, dll, uu, vv, gd)
        slots = [state[0].clone(), state[1].clone(), state[2].clone(), rr_new, drr_new, state[5].clone(),
                 state[6].clone(), mm_new, dmm_new, uu_new, vv_new]
    elif column:
        dkk, dll = eng.dev(statics['dkk'], n), eng.dev(statics['dll'], n)
        statics['rr_mm_area']
        rr_new, mm_new, uu_new, vv_new = eng.column_step(p, state, dkk, dll, uu, vv, gd)
        slots = [state[0].clone(), state[1].clone(), state[2].clone(), rr_new, state[4].clone(), state[5].clone(),
                 state[6].clone(), mm_new, state[8].clone(), uu_new, vv_new]
    else:
        x, uu_new, vv_new = eng.rk3_general(p, state, _statics_dev(eng, n), uu, vv, gd)
        slots = x + [uu_new, vv_new]
    return _pack11([_out(eng, t, like_dev) for t in slots])


def rhs_frozen(dt, var_in):
    """EXTENSION (not in the reference): rhs_default with the mean-flow tendencies du_st, dv_st replaced by zeros --
    plugged into model_config['rhs'] (L:691) it makes RK3 advance the rays against a mean flow frozen over the step."""
    t = rhs_default(dt, var_in)
    t[9] = t[9] * 0
    t[10] = t[10] * 0
    return t


def RK3_frozen(dt, var):
    """EXTENSION (not in the reference; never used by RK3): one step of the frozen-background mode "M2" -- the rays
    advance through the reference's RK3 with rhs = rhs_frozen (all three stages in registers of one fused CUDA sweep),
    their pseudo-momentum flux is deposited once, at the end of the step, and the mean flow takes one forward-Euler
    step with it: uu + dt * du_dt(vv, dF/dz), vv + dt * dv_dt(uu, dF/dz).  Same 11-slot state vector in and out as RK3.
    Constant N with HPROP_GLOBAL and saturate_online off."""
    statics['dkk'], statics['dll']
    eng = _engine()
    p = _params(dt)
    if p.hprop or p.saturate_online or _bvf_profile() is not None or p.G > eng.column_max_levels():
        raise NotImplementedError("RK3_frozen covers the constant-N column mode (HPROP_GLOBAL False, saturate_online False)")
    like_dev = _any_dev(eng, *var)
    n = _size(var[3])
    state = [eng.dev(x, n) for x in var[:9]]
    uu, vv = eng.dev(var[9], p.G), eng.dev(var[10], p.G)
    dkk, dll = eng.dev(statics['dkk'], n), eng.dev(statics['dll'], n)
    rr_new, mm_new, uu_new, vv_new = eng.column_step_frozen(p, state, dkk, dll, uu, vv, _grid_devs(eng))
    slots = [state[0].clone(), state[1].clone(), state[2].clone(), rr_new, state[4].clone(), state[5].clone(),
             state[6].clone(), mm_new, state[8].clone(), uu_new, vv_new]
    return _pack11([_out(eng, t, like_dev) for t in slots])


# default setup, as installed at import time by the reference (L:704-726)
set_model_setup(
    u0=80, phi0=np.deg2rad(-60), sig_phi=np.deg2rad(3), rr0=30000, rr1=40000, sig_rr=10000, drr=1,
    bvf=0.01, rhs=rhs_default, geostrophy=True, boussinesq=False, hh=8500, rhobar0=1.2, kappa=0.95,
    saturate_online=True)
set_statics(int_dll=1, int_dkk=1, rr_mm_area=0)
