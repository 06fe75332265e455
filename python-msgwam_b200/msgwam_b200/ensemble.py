"""Device-resident, optionally sharded ray store.

The drop-in ``libprop.RK3`` takes and returns the reference's 11-slot state vector; for large
ensembles that means moving every field over PCIe each step.  ``RayEnsemble`` keeps the
structure-of-arrays store (9 state fields, 3 statics, 2 derived statics) in HBM and advances it in
place with the same kernels.  With ``torch.distributed`` initialised (one process per GPU, NCCL) each
rank owns a contiguous slice of the rays; the only exchange is the all-reduce of the deposited flux
(2 x (G-1) doubles per RK stage, batched into two calls per step), after which every rank advances its
replica of the mean flow identically.

Ray deletion (``compact``) has no counterpart in the reference, whose only related predicate is
``out_of_domain`` in wave_projection (L:129-130): rays outside the deposit domain, or with
|m| >= m_crit (critical-level pile-up), are removed between steps by a stable stream compaction.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, lib
from ._engine import Engine

STATE = ("dens", "lam", "phi", "rr", "drr", "kk", "ll", "mm", "dmm")
STATICS = ("dkk", "dll", "rr_mm_area")
_vp = ctypes.c_void_p
# MSGWAM_FIXED_HIST=0: the CTA histogram of the deposit always uses fp64 compare-and-swap atomics (developer switch)
_FIXED_POINT_HISTOGRAM = __import__("os").environ.get("MSGWAM_FIXED_HIST", "1") != "0"


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


class RayEnsemble:
    def __init__(self, state, dkk, dll, rr_mm_area, uu, vv, grid, grids, rhobar, pressure_gradient, *,
                 bvf, phi0, kappa=1.0, saturate_online=False, hprop=False, distributed=None,
                 rot_earth=_cabi.ROT_EARTH_DEFAULT, rad_earth=_cabi.RAD_EARTH_DEFAULT):
        """state: the 9 per-ray arrays of this rank's slice (numpy or torch); grid fields are replicated.
        rot_earth / rad_earth: the module constants ROT_EARTH / RAD_EARTH of the reference (L:3-4), should the caller
        have overridden them."""
        self.eng = Engine.get()
        eng = self.eng
        n = int(np.size(state[3])) if not hasattr(state[3], "numel") else int(state[3].numel())
        self.n = n
        self.cap = max(n, 1)
        # one slab per buffer so that compaction can ping-pong between two of them
        self._slab = eng.empty(len(STATE) + len(STATICS) + 2, self.cap)
        self._slab2 = None
        self._stage1 = None
        self._old = None            # rr, drr, mm before a step (the post-step clamp needs both ends)
        self._params_cache = None
        self._rays_cache = None
        self.steps_done = 0
        # deposit bounds of the previous / running step (msgwam_rays_t.bounds): what lets the CTA histogram of the
        # deposit accumulate in fixed point.  Zero = unknown (the next step takes the fp64 path and measures them).
        self._bounds = eng.zeros(16)
        self._slab_version = None
        names = STATE + STATICS
        for i, (nm, a) in enumerate(zip(names, list(state) + [dkk, dll, rr_mm_area])):
            self._slab[i, :n].copy_(eng.dev(a, n))
        self.cfg = dict(bvf=bvf, phi0=phi0, kappa=kappa, saturate_online=saturate_online, hprop=hprop,
                        rot_earth=rot_earth, rad_earth=rad_earth)
        self.grid_host, self.grids_host = np.asarray(grid, dtype=np.float64), np.asarray(grids, dtype=np.float64)
        self.G = len(self.grids_host)
        self.grid_devs = tuple(eng.dev(a) for a in (self.grid_host, self.grids_host,
                                                    np.broadcast_to(np.asarray(rhobar, dtype=np.float64), (self.G,)),
                                                    np.asarray(pressure_gradient, dtype=np.float64).reshape(2, self.G)))
        if np.ndim(bvf) > 0:                        # extension: N profile on grids (general path only)
            self.grid_devs = self.grid_devs + (eng.dev(np.asarray(bvf, dtype=np.float64).reshape(self.G)),)
        self.uu, self.vv = eng.dev(uu, self.G).clone(), eng.dev(vv, self.G).clone()
        self._uu2, self._vv2 = eng.empty(self.G), eng.empty(self.G)
        self.work = eng.zeros(int(lib.msgwam_column_work_doubles(self.G)))
        self.dist = _dist() if distributed is None else (distributed or None)
        self.exchange = None
        if self.dist is not None and self.dist.get_world_size() > 1:
            from .distributed import PeerExchange
            self.exchange = PeerExchange.get(self.G)
        self._derive()

    @classmethod
    def from_scenario(cls, sc, **kw):
        return cls(sc.state, sc.dkk, sc.dll, sc.rr_mm_area, sc.uu, sc.vv, sc.grid, sc.grids, sc.rhobar,
                   sc.pressure_gradient, bvf=sc.model["bvf"], phi0=sc.model["phi0"], kappa=sc.model.get("kappa", 1.0),
                   saturate_online=sc.model.get("saturate_online", False), hprop=sc.hprop, **kw)

    # ---- views ------------------------------------------------------------------------------------
    def field(self, name):
        names = STATE + STATICS + ("ff", "pkl")
        return self._slab[names.index(name), :self.n]

    def _derive(self):
        eng = self.eng
        check(lib.msgwam_derive_statics(eng.ptr(self.field("phi")), eng.ptr(self.field("dkk")), eng.ptr(self.field("dll")),
                                        eng.ptr(self.field("ff")), eng.ptr(self.field("pkl")), self.n,
                                        2 * self.cfg["rot_earth"], eng.stream), "msgwam_derive_statics")
        eng.launches += 1

    def params(self, dt) -> _cabi.Params:
        """POD snapshot for time step dt (cached: building it costs ~0.1 ms of host time, more than a step at 1e6 rays)."""
        c = self._params_cache
        if c is None or c[0] != dt:
            c = self._params_cache = (dt, _cabi.snapshot_params(dt, grid=self.grid_host, grids=self.grids_host, **self.cfg))
        return c[1]

    def _rays(self, hand: int = 3) -> _cabi.Rays:
        """The ray store as a msgwam_rays_t; `hand` doubles per ray of hand-over scratch (3; 7 with an N(z) profile)."""
        key = (self._slab.data_ptr(), self.n, None if self._stage1 is None else self._stage1.data_ptr())
        if self._rays_cache is not None and self._rays_cache[0] == key and self._stage1 is not None and self._stage1.numel() >= hand * self.n:
            return self._rays_cache[1]
        r = self._build_rays(hand)
        self._rays_cache = ((self._slab.data_ptr(), self.n, self._stage1.data_ptr()), r)
        return r

    def _build_rays(self, hand: int = 3) -> _cabi.Rays:
        r = _cabi.Rays()
        for nm in STATE + STATICS + ("ff", "pkl"):
            setattr(r, nm, self.field(nm).data_ptr())
        if self._stage1 is None or self._stage1.numel() < hand * self.n:
            self._stage1 = self.eng.empty(max(hand * self.cap, 1))  # stage-1 hand-over between the two sweeps
        r.stage1 = self._stage1.data_ptr()
        r.bounds = self._bounds.data_ptr() if _FIXED_POINT_HISTOGRAM else 0
        return r

    def _reduce(self, t):
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.all_reduce(t)

    # ---- stepping ---------------------------------------------------------------------------------
    def _is_column(self, p):
        """The fused constant-N column step applies (HPROP off, online saturation off, scalar N, grid fits)."""
        return not p.hprop and not p.saturate_online and len(self.grid_devs) == 4 and self.G <= self.eng.column_max_levels()

    def step(self, dt, nsteps=1, _outs=None):
        """Advance rays and mean flow in place by nsteps RK3 steps (reference RK3 + rhs_default semantics).
        _outs = (rr_out, mm_out): the constant-N column step writes the new rr, mm there instead (advance())."""
        eng = self.eng
        p = self.params(dt)
        g = eng.grid_struct(self.grid_devs)
        column = self._is_column(p)
        assert _outs is None or (column and nsteps == 1)
        self._check_bounds(dt)
        sharded = self.dist is not None and self.dist.get_world_size() > 1
        column_nz = (not p.hprop and not p.saturate_online and len(self.grid_devs) == 5 and
                     (not sharded or self.exchange is not None) and self.G <= eng.column_nz_max_levels())
        if column_nz:                               # N(z) extension, fused; in place
            rays = self._rays(hand=7)
            P = eng.ptr
            for _ in range(nsteps):
                check(lib.msgwam_column_step_nz(p, rays, self.n, g, P(self.uu), P(self.vv), P(self.work), P(self.field("rr")),
                                                P(self.field("drr")), P(self.field("mm")), P(self.field("dmm")), P(self._uu2),
                                                P(self._vv2), self.exchange.next(2) if self.exchange is not None else None,
                                                eng.stream), "msgwam_column_step_nz")
                eng.launches += 2
                self.uu, self._uu2 = self._uu2, self.uu
                self.vv, self._vv2 = self._vv2, self.vv
            return
        nc = self.G - 1
        if column:
            rays = self._rays()
            s = eng.stream
            rr, mm = self.field("rr"), self.field("mm")
            rr_o, mm_o = (rr, mm) if _outs is None else _outs
        for _ in range(nsteps):
            if column:
                if self.exchange is not None:       # all-reduces fused into the tails of the two sweeps (peer memory)
                    check(lib.msgwam_column_step_p2p(p, rays, self.n, g, eng.ptr(self.uu), eng.ptr(self.vv), eng.ptr(self.work),
                                                     eng.ptr(rr_o), eng.ptr(mm_o), eng.ptr(self._uu2), eng.ptr(self._vv2),
                                                     self.exchange.next(2), s), "msgwam_column_step_p2p")
                elif self.dist is None or self.dist.get_world_size() == 1:
                    check(lib.msgwam_column_step(p, rays, self.n, g, eng.ptr(self.uu), eng.ptr(self.vv), eng.ptr(self.work),
                                                 eng.ptr(rr_o), eng.ptr(mm_o), eng.ptr(self._uu2), eng.ptr(self._vv2), s),
                          "msgwam_column_step")
                else:                               # NCCL all-reduces between the kernels
                    check(lib.msgwam_column_pass_a(p, rays, self.n, g, eng.ptr(self.uu), eng.ptr(self.vv), eng.ptr(self.work), s),
                          "msgwam_column_pass_a")
                    self._reduce(self.work[:4 * nc])
                    check(lib.msgwam_column_pass_b(p, rays, self.n, g, eng.ptr(self.uu), eng.ptr(self.vv), eng.ptr(self.work),
                                                   eng.ptr(rr_o), eng.ptr(mm_o), s), "msgwam_column_pass_b")
                    self._reduce(self.work[4 * nc:6 * nc])
                    check(lib.msgwam_column_finish(p, g, eng.ptr(self.uu), eng.ptr(self.vv), eng.ptr(self.work),
                                                   eng.ptr(self._uu2), eng.ptr(self._vv2), s), "msgwam_column_finish")
                    eng.launches += 1
                eng.launches += 2
                self.uu, self._uu2 = self._uu2, self.uu
                self.vv, self._vv2 = self._vv2, self.vv
            else:
                state = [self.field(nm) for nm in STATE]
                st = [self.field(nm) for nm in STATICS]
                x, uu, vv = eng.rk3_general(p, state, st, self.uu, self.vv, self.grid_devs, self._reduce, in_place=True,
                                            ff=self.field("ff"))
                self.uu, self.vv = uu, vv
                if p.hprop:
                    self._derive()                  # phi moved: ff = 2 Omega sin(phi) for the column kernels

    def step_frozen(self, dt, nsteps=1):
        """EXTENSION, not the reference's scheme (never used by step / advance): the frozen-background mode "M2" --
        every ray runs its three RK stages in registers against the mean flow of the start of the step, the rays'
        flux is deposited ONCE (at the end of the step) and uu += dt * du_dt(...), vv += dt * dv_dt(...) once per step
        (msgwam_column_step_frozen: one sweep and one launch per step).  Constant N, HPROP and online saturation off."""
        eng = self.eng
        p = self.params(dt)
        if not self._is_column(p):
            raise _cabi.MsgwamError("step_frozen covers the constant-N column mode (HPROP off, saturate_online off)")
        sharded = self.dist is not None and self.dist.get_world_size() > 1
        if sharded and self.exchange is None:
            raise _cabi.MsgwamError("step_frozen on several GPUs needs the peer-memory exchange")
        self._check_bounds(dt)
        g = eng.grid_struct(self.grid_devs)
        rays = self._rays()
        P = eng.ptr
        rr, mm = self.field("rr"), self.field("mm")
        for _ in range(nsteps):
            check(lib.msgwam_column_step_frozen(p, rays, self.n, g, P(self.uu), P(self.vv), P(self.work), P(rr), P(mm),
                                                P(self._uu2), P(self._vv2), self.exchange.next(1) if sharded else None,
                                                eng.stream), "msgwam_column_step_frozen")
            eng.launches += 1
            self.uu, self._uu2 = self._uu2, self.uu
            self.vv, self._vv2 = self._vv2, self.vv

    def _check_bounds(self, dt):
        p = self.params(dt)
        if self._slab._version != self._slab_version and _FIXED_POINT_HISTOGRAM and not p.hprop and not p.saturate_online:
            # the store is new or was written through torch (an upload, a caller editing a field() view): the deposit
            # bounds of the last step no longer describe it -- one cheap sweep measures them at the current state
            self.measure_bounds(dt)

    def measure_bounds(self, dt):
        """Deposit bounds of the store as it is (msgwam_column_bounds): what lets the next column step accumulate its
        deposits in fixed point.  step() calls this whenever the store is new or was edited through torch."""
        eng = self.eng
        p = self.params(dt)
        check(lib.msgwam_column_bounds(p, self._build_rays(7 if len(self.grid_devs) == 5 else 3), self.n,
                                       eng.grid_struct(self.grid_devs), eng.stream), "msgwam_column_bounds")
        eng.launches += 1
        self._slab_version = self._slab._version

    # ---- the driver's loop on the device ------------------------------------------------------------
    def advance(self, dt, nsteps, saturate=True, history=None):
        """The reference driver's time loop (raytracer.py:157-188) without leaving the device: nsteps times
        RK3, then -- unless online saturation is on (R:182) -- the post-step clamp saturation(direct=True)
        on the propagated wave action.  In the column modes the clamp is fused into the end of the step's second
        sweep (msgwam_column_advance / _nz: two launches per step of the driver loop); the general modes run it as one
        kernel after the step (msgwam_saturation_step).  `history` (a History) receives strided snapshots in place of
        the driver's (nt_max + 1, n) host arrays (R:125-136, 178-180)."""
        eng = self.eng
        p = self.params(dt)
        clamp = saturate and not p.saturate_online
        if history is not None and history.count == 0:
            history.record(self, 0)
        if not clamp:
            for k in range(nsteps):
                self.step(dt)
                self.steps_done += 1
                if history is not None and self.steps_done % history.every == 0:
                    history.record(self, self.steps_done)
            if self.exchange is not None:
                self.check_errors()
            return
        sharded = self.dist is not None and self.dist.get_world_size() > 1
        profile = len(self.grid_devs) == 5
        fused = (not p.hprop and (not sharded or self.exchange is not None) and
                 self.G <= (eng.column_nz_max_levels() if profile else eng.column_max_levels()))
        P = eng.ptr
        for k in range(1, nsteps + 1):
            if fused:
                self._check_bounds(dt)
                g = eng.grid_struct(self.grid_devs)
                rays = self._rays(hand=7 if profile else 3)
                peers = self.exchange.next(2) if self.exchange is not None else None
                rr, mm, dens = self.field("rr"), self.field("mm"), self.field("dens")
                if profile:
                    check(lib.msgwam_column_advance_nz(p, rays, self.n, g, P(self.uu), P(self.vv), P(self.work), P(rr),
                                                       P(self.field("drr")), P(mm), P(self.field("dmm")), P(dens), P(self._uu2),
                                                       P(self._vv2), peers, eng.stream), "msgwam_column_advance_nz")
                else:
                    check(lib.msgwam_column_advance(p, rays, self.n, g, P(self.uu), P(self.vv), P(self.work), P(rr), P(mm), P(dens),
                                                    P(self._uu2), P(self._vv2), peers, eng.stream), "msgwam_column_advance")
                eng.launches += 2
                self.uu, self._uu2 = self._uu2, self.uu
                self.vv, self._vv2 = self._vv2, self.vv
            else:
                if self._old is None or self._old.shape[1] < self.cap:
                    self._old = eng.empty(3, self.cap)
                old = self._old[:, :self.n]
                old[0].copy_(self.field("rr")); old[1].copy_(self.field("drr")); old[2].copy_(self.field("mm"))
                self.step(dt)
                dens = self.field("dens")
                gd = self.grid_devs
                check(lib.msgwam_saturation_step(
                    p, self.n, P(dens), P(old[0]), P(self.field("rr")), P(old[1]),
                    P(self.field("drr")), P(self.field("kk")), P(self.field("ll")), P(old[2]),
                    P(self.field("mm")), P(self.field("dkk")), P(self.field("dll")),
                    P(self.field("rr_mm_area")), P(gd[1]), P(gd[2]),
                    P(gd[4]) if len(gd) > 4 else _vp(0), P(dens), eng.stream), "msgwam_saturation_step")
                eng.launches += 1
            self.steps_done += 1
            if history is not None and self.steps_done % history.every == 0:
                history.record(self, self.steps_done)
        if self.exchange is not None:
            self.check_errors()                     # a missing peer must not yield silently invalid results

    # ---- deletion ---------------------------------------------------------------------------------
    def compact(self, dt=0.0, m_crit=float("inf")) -> int:
        """Delete rays outside the deposit domain (L:129-130) or with |m| >= m_crit.  Returns survivors."""
        eng = self.eng
        torch = eng.torch
        p = self.params(dt)
        n = self.n
        if n == 0:
            return 0
        keep = torch.empty(n + 16, dtype=torch.uint8, device=eng.device)
        check(lib.msgwam_flag_rays(p, n, eng.ptr(self.field("rr")), eng.ptr(self.field("drr")), eng.ptr(self.field("mm")),
                                   float(m_crit), _vp(keep.data_ptr()), eng.stream), "msgwam_flag_rays")
        if self._slab2 is None:
            self._slab2 = eng.empty(*self._slab.shape)
        nf = self._slab.shape[0]
        ins = (_vp * nf)(*[self._slab[f].data_ptr() for f in range(nf)])
        outs = (_vp * nf)(*[self._slab2[f].data_ptr() for f in range(nf)])
        count = torch.zeros(1, dtype=torch.int64, device=eng.device)
        scratch = torch.empty(int(lib.msgwam_compact_scratch_bytes(n)), dtype=torch.uint8, device=eng.device)
        check(lib.msgwam_compact(n, _vp(keep.data_ptr()), nf, ins, outs, _vp(count.data_ptr()), _vp(scratch.data_ptr()),
                                 eng.stream), "msgwam_compact")
        eng.launches += 4
        self._slab, self._slab2 = self._slab2, self._slab
        self._slab_version = None                   # the rays are dealt out to the CTAs anew: deposit bounds unknown
        off = int(lib.msgwam_column_error_offset(self.G))
        both = torch.stack((count[0].to(torch.float64), self.work[off])).cpu()     # one synchronising read for both
        self.n = int(both[0].item())
        self._raise_on(float(both[1].item()))
        return self.n

    def rebalance(self) -> int:
        """Even out the ray counts of the ranks after deletion has skewed them (the reference has neither
        deletion nor ranks): surplus ranks hand the rays at the end of their store to deficit ranks, all fields of a ray
        together.  Rays are independent given the mean flow, so any assignment gives the same physics; the order of the
        rays in a rank's store changes.  Collective over the process group; returns this rank's new count."""
        if self.dist is None or self.dist.get_world_size() < 2:
            return self.n
        from .distributed import exchange_rows, rebalance_plan
        eng, torch, dist = self.eng, self.eng.torch, self.dist
        counts = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=eng.device)
        counts[dist.get_rank()] = self.n
        dist.all_reduce(counts)
        plan, target = rebalance_plan(counts.tolist())
        if not plan:
            return self.n
        rank = dist.get_rank()
        gives = sum(n for src, _, n in plan if src == rank)
        rows = self._slab[:, self.n - gives:self.n].t().contiguous() if gives else self._slab[:, :0].t().contiguous()
        # only the rays that leave are transposed into rows (exchange_rows sends the tail of what it is given)
        _, got = exchange_rows(rows, gives, plan, rank, dist)
        new_n = self.n - gives + got.shape[0]
        assert new_n == target[rank], (new_n, target[rank])
        if new_n > self.cap:
            grown = eng.empty(self._slab.shape[0], new_n)
            grown[:, :self.n - gives].copy_(self._slab[:, :self.n - gives])
            self._slab, self._slab2, self.cap = grown, None, new_n
            self._stage1 = self._old = None
        if got.shape[0]:
            self._slab[:, self.n - gives:new_n].copy_(got.t())
        self.n = new_n
        self._rays_cache = None
        self._slab_version = None
        return self.n

    def _raise_on(self, word):
        if word != 0.0:
            self.work[int(lib.msgwam_column_error_offset(self.G))] = 0.0
            self._bounds.zero_()
            self._slab_version = None               # the next step measures the deposit bounds anew
            what = {1.0: "a peer did not deliver its deposit (the bounded wait of the peer-memory all-reduce timed out)",
                    2.0: "the mean-flow slices of pass B did not all arrive in time (its CTAs were not co-resident)"}.get(word, "code %g" % word)
            raise _cabi.MsgwamError("device-side check failed: %s; the results since the last check are invalid" % what)

    def check_errors(self):
        """Raise if a bounded device-side wait of a step since the last check timed out (synchronises the stream).
        Called by every method that synchronises anyway: advance (sharded), compact, to_var, History.to_host."""
        off = int(lib.msgwam_column_error_offset(self.G))
        self._raise_on(float(self.work[off].item()))

    # ---- export -----------------------------------------------------------------------------------
    def to_var(self):
        """The reference's 11-slot state vector (numpy copies)."""
        self.check_errors()
        out = np.empty(11, dtype=object)
        for i, nm in enumerate(STATE):
            out[i] = self.field(nm).cpu().numpy()
        out[9], out[10] = self.uu.cpu().numpy(), self.vv.cpu().numpy()
        return out


class History:
    """Device-side history of an ensemble run: what the reference driver keeps in (nt_max + 1, n) host arrays
    (raytracer.py:125-136), as strided snapshots in HBM.  Fields: any of the ensemble's per-ray fields plus
    'uu', 'vv'.  `to_host()` returns numpy arrays of shape (snapshots, n) / (snapshots, G) and the step numbers."""

    def __init__(self, ens: RayEnsemble, nsnap: int, every: int = 1, fields=("dens", "rr", "drr", "mm", "dmm", "uu", "vv")):
        self.every = max(int(every), 1)
        self.fields = tuple(fields)
        self.count = 0
        self.steps = []
        eng = ens.eng
        self.buf = {f: eng.empty(nsnap, ens.G if f in ("uu", "vv") else ens.n) for f in self.fields}
        self.nsnap = nsnap
        self.ens = ens

    def record(self, ens: RayEnsemble, step: int) -> None:
        if self.count >= self.nsnap:
            raise IndexError("History is full (%d snapshots)" % self.nsnap)
        for f in self.fields:
            src = ens.uu if f == "uu" else ens.vv if f == "vv" else ens.field(f)
            self.buf[f][self.count, :src.numel()].copy_(src)
        self.steps.append(int(step))
        self.count += 1

    def to_host(self):
        self.ens.check_errors()
        out = {f: self.buf[f][:self.count].cpu().numpy() for f in self.fields}
        out["steps"] = np.asarray(self.steps)
        return out
