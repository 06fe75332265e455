"""Device plumbing behind the drop-in functions of :mod:`msgwam_b200.libprop`.

torch is used for device memory, streams and (in :mod:`msgwam_b200.distributed`) NCCL; all
arithmetic is done by the CUDA kernels in ``csrc/`` through the C ABI.  There is no CPU path:
without a CUDA device every entry point raises.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import Grid, Params, Rays, check, lib

_vp = ctypes.c_void_p


def _torch():
    import torch
    return torch


class Engine:
    """Per-process device context: scratch buffers cached by size, current-stream launches."""

    _instance = None

    @classmethod
    def get(cls) -> "Engine":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _cabi.MsgwamError("msgwam_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._work = {}
        self._stage = {}
        self._grid_cache = {}
        self._derived = None
        self._scratch = None
        self._gmax = None
        self._gmax_nz = None
        self._bounds = None        # deposit bounds of the stateless column calls (measured by a pre-pass every call)
        self.launches = 0          # kernels launched through this engine (bench.py reports it)

    # ---- helpers -----------------------------------------------------------------------------
    @property
    def stream(self) -> _vp:
        return _vp(self.torch.cuda.current_stream(self.device).cuda_stream)

    def is_dev(self, x) -> bool:
        return isinstance(x, self.torch.Tensor) and x.is_cuda

    def dev(self, x, n=None):
        """float64 contiguous device tensor of length n (scalars are broadcast)."""
        torch = self.torch
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=torch.float64)
        else:
            a = np.asarray(x, dtype=np.float64)
            if n is not None and a.ndim == 0:
                return torch.full((n,), float(a), dtype=torch.float64, device=self.device)
            a = np.ascontiguousarray(a)
            if not a.flags.writeable:
                a = a.copy()
            t = torch.from_numpy(a).to(self.device)
        if n is not None and t.ndim == 0:
            t = t.expand(n)
        return t.contiguous()

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float64, device=self.device)

    def zeros(self, *shape):
        return self.torch.zeros(*shape, dtype=self.torch.float64, device=self.device)

    @staticmethod
    def ptr(t) -> _vp:
        return _vp(t.data_ptr()) if t is not None else _vp(0)

    def column_work(self, G: int):
        """Deposit buffers D0|D1|D2 for the fused column step; zero on creation, kept zero by the
        finish kernel."""
        w = self._work.get(G)
        if w is None:
            w = self.zeros(int(lib.msgwam_column_work_doubles(G)))
            self._work[G] = w
        return w

    def column_max_levels(self) -> int:
        """Largest G the fused column kernels take on this device (their shear tables live in shared memory);
        taller grids run stage by stage through the general kernels."""
        if self._gmax is None:
            self._gmax = int(lib.msgwam_column_max_levels())
        return self._gmax

    def ray_scratch(self, n: int, per_ray: int = 3):
        """per_ray * n doubles for the stage-1 hand-over between the two sweeps (msgwam_rays_t.stage1): 3 per ray,
        7 with an N(z) profile."""
        t = self._scratch
        if t is None or t.numel() < per_ray * n:
            t = self._scratch = self.empty(max(per_ray * n, 1))
        return t

    def fresh_bounds(self, p, rays, n, g):
        """Stateless column calls see a new state every time: msgwam_column_bounds measures the bound of its deposits, so
        that the step's CTA histograms accumulate in fixed point (include/msgwam_b200.h: msgwam_rays_t.bounds)."""
        if self._bounds is None:
            self._bounds = self.zeros(16)
        rays.bounds = self._bounds.data_ptr()
        check(lib.msgwam_column_bounds(p, rays, n, g, self.stream), "msgwam_column_bounds")
        self.launches += 1

    def column_nz_max_levels(self) -> int:
        if self._gmax_nz is None:
            self._gmax_nz = int(lib.msgwam_column_nz_max_levels())
        return self._gmax_nz

    def column_step_nz(self, p: Params, state, dkk, dll, uu, vv, grid_devs, exchange=None):
        """The fused column step with an N(z) profile (grid_devs carries it as its fifth entry); with the rays
        sharded over several GPUs `exchange` (distributed.PeerExchange) carries the sums of the deposit.
        Returns (rr, drr, mm, dmm, uu, vv) after the step."""
        dens, lam, phi, rr, drr, kk, ll, mm, dmm = state
        n = rr.numel()
        ff, pkl = self.derived_statics(phi, dkk, dll, p.two_rot)
        rays = Rays()
        for k, t in (("dens", dens), ("phi", phi), ("rr", rr), ("drr", drr), ("kk", kk), ("ll", ll), ("mm", mm),
                     ("dmm", dmm), ("dkk", dkk), ("dll", dll), ("ff", ff), ("pkl", pkl)):
            setattr(rays, k, t.data_ptr())
        rays.stage1 = self.ray_scratch(n, 7).data_ptr()
        g = self.grid_struct(grid_devs)
        self.fresh_bounds(p, rays, n, g)
        work = self.column_work(p.G)
        outs = [self.empty(n) for _ in range(4)]
        uu_out, vv_out = self.empty(p.G), self.empty(p.G)
        check(lib.msgwam_column_step_nz(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), *[self.ptr(t) for t in outs],
                                        self.ptr(uu_out), self.ptr(vv_out), exchange.next(2) if exchange is not None else None,
                                        self.stream), "msgwam_column_step_nz")
        self.launches += 2
        return outs[0], outs[1], outs[2], outs[3], uu_out, vv_out

    def host_stage(self, n: int, G: int, nz: bool = False):
        key = (n, G, bool(nz))
        s = self._stage.get(key)
        if s is None:
            self._stage.clear()                 # one staging buffer at a time
            s = self.empty(int((lib.msgwam_host_stage_doubles_nz if nz else lib.msgwam_host_stage_doubles)(n, G)))
            self._stage[key] = s
        return s

    def grid_on_device(self, grid, grids, rhobar, pg, bvf=None):
        """Device copies of the background profiles, re-uploaded only when their values change.  bvf: None for
        the reference's scalar N, or (extension) the N profile on grids."""
        G = len(grids)
        host = (np.ascontiguousarray(grid, dtype=np.float64), np.ascontiguousarray(grids, dtype=np.float64),
                np.ascontiguousarray(np.broadcast_to(np.asarray(rhobar, dtype=np.float64), (G,))),
                np.ascontiguousarray(pg, dtype=np.float64).reshape(2, G))
        if bvf is not None:
            host = host + (np.ascontiguousarray(bvf, dtype=np.float64).reshape(G),)
        c = self._grid_cache.get(G)
        if c is not None and len(c[0]) == len(host) and all(np.array_equal(a, b) for a, b in zip(c[0], host)):
            return c[1]
        devs = tuple(self.torch.from_numpy(a.copy()).to(self.device) for a in host)
        self._grid_cache[G] = (tuple(a.copy() for a in host), devs)
        return devs

    def grid_struct(self, devs) -> Grid:
        return Grid(self.ptr(devs[0]), self.ptr(devs[1]), self.ptr(devs[2]), self.ptr(devs[3]),
                    self.ptr(devs[4]) if len(devs) > 4 else _vp(0))

    def derived_statics(self, phi, dkk, dll, two_rot):
        """ff = 2*ROT*sin(phi), pkl = dkk*dll on the device; cached on (storage, version) of the inputs.  The cache
        entry keeps the three keyed tensors alive, so the caching allocator cannot hand their addresses to
        different data while the key is still in use (a freed-and-reused address would match with stale values)."""
        key = tuple((t.data_ptr(), t._version, t.numel()) for t in (phi, dkk, dll)) + (two_rot,)
        if self._derived is not None and self._derived[0] == key:
            return self._derived[1], self._derived[2]
        n = phi.numel()
        ff, pkl = self.empty(n), self.empty(n)
        check(lib.msgwam_derive_statics(self.ptr(phi), self.ptr(dkk), self.ptr(dll), self.ptr(ff), self.ptr(pkl),
                                        n, two_rot, self.stream), "msgwam_derive_statics")
        self.launches += 1
        self._derived = (key, ff, pkl, (phi, dkk, dll))
        return ff, pkl

    # ---- fused column step on device tensors ---------------------------------------------------
    def column_step(self, p: Params, state, dkk, dll, uu, vv, grid_devs, rr_out=None, mm_out=None,
                    reduce_fn=None, exchange=None):
        """state: 9 device tensors (reference order).  Returns (rr_new, mm_new, uu_new, vv_new).

        Multi-GPU: exchange (distributed.PeerExchange) fuses the all-reduce of the deposit into the chain /
        finish kernels over peer memory; otherwise reduce_fn(tensor) all-reduces the buffers in place (NCCL)."""
        dens, lam, phi, rr, drr, kk, ll, mm, dmm = state
        n = rr.numel()
        ff, pkl = self.derived_statics(phi, dkk, dll, p.two_rot)
        rays = Rays()
        for k, t in (("dens", dens), ("phi", phi), ("rr", rr), ("drr", drr), ("kk", kk), ("ll", ll), ("mm", mm),
                     ("dmm", dmm), ("dkk", dkk), ("dll", dll), ("ff", ff), ("pkl", pkl)):
            setattr(rays, k, t.data_ptr())
        rays.stage1 = self.ray_scratch(n).data_ptr()
        g = self.grid_struct(grid_devs)
        self.fresh_bounds(p, rays, n, g)
        work = self.column_work(p.G)
        rr_out = self.empty(n) if rr_out is None else rr_out
        mm_out = self.empty(n) if mm_out is None else mm_out
        uu_out, vv_out = self.empty(p.G), self.empty(p.G)
        s = self.stream
        if exchange is not None:        # two launches; the all-reduces run in the sweeps' tails over NVLink peer memory
            check(lib.msgwam_column_step_p2p(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), self.ptr(rr_out),
                                             self.ptr(mm_out), self.ptr(uu_out), self.ptr(vv_out), exchange.next(2), s),
                  "msgwam_column_step_p2p")
            self.launches += 2
            return rr_out, mm_out, uu_out, vv_out
        elif reduce_fn is None:
            check(lib.msgwam_column_step(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), self.ptr(rr_out),
                                         self.ptr(mm_out), self.ptr(uu_out), self.ptr(vv_out), s), "msgwam_column_step")
        else:
            nc = p.G - 1
            check(lib.msgwam_column_pass_a(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), s), "msgwam_column_pass_a")
            reduce_fn(work[:4 * nc])
            check(lib.msgwam_column_pass_b(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), self.ptr(rr_out),
                                           self.ptr(mm_out), s), "msgwam_column_pass_b")
            reduce_fn(work[4 * nc:6 * nc])
            check(lib.msgwam_column_finish(p, g, self.ptr(uu), self.ptr(vv), self.ptr(work), self.ptr(uu_out),
                                           self.ptr(vv_out), s), "msgwam_column_finish")
        self.launches += 3
        return rr_out, mm_out, uu_out, vv_out

    def column_step_frozen(self, p: Params, state, dkk, dll, uu, vv, grid_devs):
        """Frozen-background mode "M2" (extension; msgwam_column_step_frozen) on device tensors.
        Returns (rr_new, mm_new, uu_new, vv_new)."""
        dens, lam, phi, rr, drr, kk, ll, mm, dmm = state
        n = rr.numel()
        ff, pkl = self.derived_statics(phi, dkk, dll, p.two_rot)
        rays = Rays()
        for k, t in (("dens", dens), ("phi", phi), ("rr", rr), ("drr", drr), ("kk", kk), ("ll", ll), ("mm", mm),
                     ("dmm", dmm), ("dkk", dkk), ("dll", dll), ("ff", ff), ("pkl", pkl)):
            setattr(rays, k, t.data_ptr())
        rays.stage1 = self.ray_scratch(n).data_ptr()
        g = self.grid_struct(grid_devs)
        self.fresh_bounds(p, rays, n, g)
        work = self.column_work(p.G)
        rr_out, mm_out, uu_out, vv_out = self.empty(n), self.empty(n), self.empty(p.G), self.empty(p.G)
        check(lib.msgwam_column_step_frozen(p, rays, n, g, self.ptr(uu), self.ptr(vv), self.ptr(work), self.ptr(rr_out),
                                            self.ptr(mm_out), self.ptr(uu_out), self.ptr(vv_out), None, self.stream),
              "msgwam_column_step_frozen")
        self.launches += 1
        return rr_out, mm_out, uu_out, vv_out

    # ---- general right-hand side on device tensors ----------------------------------------------
    def rhs_general(self, p: Params, state, statics, uu, vv, grid_devs, reduce_fn=None):
        """All 11 tendencies (9 ray arrays + du, dv) and the deposit, every branch of rhs_default."""
        n = state[0].numel()
        rays = Rays()
        for k, t in zip(("dens", "lam", "phi", "rr", "drr", "kk", "ll", "mm", "dmm"), state):
            setattr(rays, k, t.data_ptr())
        for k, t in zip(("dkk", "dll", "rr_mm_area"), statics):
            setattr(rays, k, t.data_ptr())
        g = self.grid_struct(grid_devs)
        tend = [self.empty(n) for _ in range(9)]
        tp = (_vp * 9)(*[t.data_ptr() for t in tend])
        proj = self.zeros(2, p.G - 1)
        s = self.stream
        check(lib.msgwam_rhs_rays(p, rays, n, g, self.ptr(uu), self.ptr(vv), tp, self.ptr(proj), s), "msgwam_rhs_rays")
        if reduce_fn is not None:
            reduce_fn(proj)
        du, dv = self.empty(p.G), self.empty(p.G)
        check(lib.msgwam_grid_tendency(p, g, self.ptr(uu), self.ptr(vv), self.ptr(proj), self.ptr(du), self.ptr(dv), s),
              "msgwam_grid_tendency")
        self.launches += 3
        return tend, du, dv, proj

    def rk3_general(self, p: Params, state, statics, uu, vv, grid_devs, reduce_fn=None, in_place=False, ff=None):
        """RK3 with rhs_default for any mode: per stage one fused ray sweep (rhs + deposit + low-storage update,
        msgwam_rk_stage_rays), the all-reduce of the deposit when sharded, and the one-CTA mean-flow stage
        (msgwam_rk_stage_grid) -- six launches per step.  in_place: overwrite `state` instead of allocating."""
        n = state[0].numel()
        x_in = list(state)
        x_out = list(state) if in_place else [self.empty(n) for _ in range(9)]
        q = [self.empty(n) for _ in range(9)]
        qu, qv = self.empty(p.G), self.empty(p.G)
        proj = self.zeros(2, p.G - 1)
        g = self.grid_struct(grid_devs)
        s = self.stream
        qp = (_vp * 9)(*[t.data_ptr() for t in q])
        xop = (_vp * 9)(*[t.data_ptr() for t in x_out])
        for stage in range(3):
            rays = Rays()
            for k, t in zip(("dens", "lam", "phi", "rr", "drr", "kk", "ll", "mm", "dmm"), x_in):
                setattr(rays, k, t.data_ptr())
            for k, t in zip(("dkk", "dll", "rr_mm_area"), statics):
                setattr(rays, k, t.data_ptr())
            if ff is not None and not p.hprop:          # 2 Omega sin(phi) of the (fixed) latitudes, see ray_rhs
                rays.ff = ff.data_ptr()
            check(lib.msgwam_rk_stage_rays(stage, p, rays, n, g, self.ptr(uu), self.ptr(vv), qp, xop, self.ptr(proj), s),
                  "msgwam_rk_stage_rays")
            if reduce_fn is not None:
                reduce_fn(proj)
            un, vn = self.empty(p.G), self.empty(p.G)
            check(lib.msgwam_rk_stage_grid(stage, p, g, self.ptr(uu), self.ptr(vv), self.ptr(proj), self.ptr(qu), self.ptr(qv),
                                           self.ptr(un), self.ptr(vn), s), "msgwam_rk_stage_grid")
            self.launches += 2
            x_in, uu, vv = x_out, un, vn
        return x_out, uu, vv
