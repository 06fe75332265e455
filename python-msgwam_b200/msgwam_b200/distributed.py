"""Multi-GPU (one process per GPU, torch.distributed / NCCL) variants of the reference-facing step.

Rays are sharded by contiguous index range across ranks; grid fields are replicated.  The only
exchange is the sum of the deposited pseudo-momentum flux (L:654-658 sums over *all* rays): an
in-place all-reduce of 4*(G-1) doubles after pass A and 2*(G-1) after pass B, enqueued on the
compute stream with no host synchronisation.  Every rank then advances its replica of the mean
flow with bit-identical inputs.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._engine import Engine


class PeerExchange:
    """Symmetric-memory inboxes for the all-reduce that is fused into the tails of the two sweeps
    (csrc/column_step.cu: p2p_allreduce, msgwam_column_step_p2p).  One instance per (G, process group); `next()` hands out the
    msgwam_peers_t of the next reduction -- the epoch sequence is identical on all ranks because every rank
    runs the same sequence of steps."""

    _cache = {}

    @classmethod
    def get(cls, G):
        """The exchange for grids of G levels, or None (single rank, MSGWAM_NCCL_ALLREDUCE=1, or symmetric
        memory unavailable -- then the caller all-reduces with NCCL)."""
        import os
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            return None
        if os.environ.get("MSGWAM_NCCL_ALLREDUCE", "0") == "1" or dist.get_world_size() > _cabi.MAX_PEERS:
            return None
        if G not in cls._cache:
            ex, why = None, ""
            try:
                ex = cls(G)
            except Exception as exc:                      # no P2P / symmetric memory on this system
                why = str(exc)
            # the choice between peer memory and NCCL must be the same on every rank: one rank polling inboxes while
            # another waits in an NCCL all-reduce would hang both
            import torch
            ok = torch.tensor([1 if ex is not None else 0], dtype=torch.int32, device=Engine.get().device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if ex is not None or why:
                    import warnings
                    warnings.warn("msgwam_b200: peer-memory all-reduce unavailable on some rank (%s); all ranks use NCCL" % (why or "another rank",))
                ex = None
            cls._cache[G] = ex
        return cls._cache[G]

    def __init__(self, G):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        eng = Engine.get()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        n = int(_cabi.lib.msgwam_p2p_inbox_doubles(G, self.world))
        self.inbox = symm_mem.empty(n, dtype=torch.float64, device=eng.device)
        self.inbox.zero_()
        self.handle = symm_mem.rendezvous(self.inbox, dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        self.handle.barrier()                             # every inbox is zeroed before anyone pushes
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.epoch = 0
        import os
        if os.environ.get("MSGWAM_PEER_TIMEOUT_S"):
            _cabi.check(_cabi.lib.msgwam_set_peer_timeout(float(os.environ["MSGWAM_PEER_TIMEOUT_S"])), "msgwam_set_peer_timeout")

    def next(self, count=1):
        """msgwam_peers_t for the next `count` reductions (the struct carries the epoch of the first)."""
        pe = _cabi.Peers()
        pe.world, pe.rank, pe.epoch = self.world, self.rank, self.epoch + 1
        self.epoch += count
        for r, ptr in enumerate(self.ptrs):
            pe.inbox[r] = ptr
        return pe


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def all_reduce_sum(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)


_stage = {}          # (n, device) -> staging slab for the sharded host path


def rk3_host_sharded(lprop, dt, var):
    """lprop.RK3 for this rank's slice of the rays, given as HOST numpy arrays (column mode only).

    Same contract as lprop.RK3: returns an 11-slot object array of numpy arrays; uu, vv are the globally
    coupled mean flow (identical on every rank).  The step's inputs go host->device with asynchronous copies
    on the compute stream (page-locked arrays are copied without staging), the two all-reduces of the
    deposit sit between the sweeps, and rr, mm, uu, vv come back into page-locked arrays.
    """
    eng = Engine.get()
    torch = eng.torch
    p = lprop._params(dt)
    if p.hprop or p.saturate_online:
        raise NotImplementedError("sharded host path covers the column mode (HPROP off, saturate_online off)")
    prof = lprop._bvf_profile()                 # N(z) extension: rr, drr, mm, dmm change; needs the peer exchange
    exchange = PeerExchange.get(p.G)
    if prof is not None and exchange is None and _world() > 1:
        raise NotImplementedError("the sharded host path with an N(z) profile needs the peer-memory exchange")
    n = int(np.size(var[3]))
    host = [lprop._host_f64(a, n) for a in var[:9]]
    dkk, dll = lprop._host_f64(lprop.statics['dkk'], n), lprop._host_f64(lprop.statics['dll'], n)
    key = (n, eng.device)
    st = _stage.get(key)
    if st is None:
        _stage.clear()
        st = dict(slab=eng.empty(10, max(n, 1)), statics=None)
        _stage[key] = st
    slab = st["slab"]

    def up(row, a):
        t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
        slab[row, :n].copy_(t, non_blocking=True)
        return slab[row, :n]
    # dens, phi, rr, drr, kk, ll, mm, dmm (lam is not needed on the device)
    dev = {i: up(k, host[i]) for k, i in enumerate((0, 2, 3, 4, 5, 6, 7, 8))}
    skey = (id(lprop.statics['dkk']), id(lprop.statics['dll']), dkk.ctypes.data, dll.ctypes.data)
    if not lprop._statics_frozen or st["statics"] != skey:       # see libprop.freeze_statics
        up(8, dkk); up(9, dll)
        st["statics"] = skey
    state = [dev[0], dev[2], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], dev[8]]
    uu = torch.from_numpy(np.ascontiguousarray(var[9], dtype=np.float64)).to(eng.device, non_blocking=True)
    vv = torch.from_numpy(np.ascontiguousarray(var[10], dtype=np.float64)).to(eng.device, non_blocking=True)
    gd = lprop._grid_devs(eng)
    if prof is not None:
        rr_new, drr_new, mm_new, dmm_new, uu_new, vv_new = eng.column_step_nz(p, state, slab[8, :n], slab[9, :n], uu, vv, gd,
                                                                             exchange=exchange)
        res = (rr_new, mm_new, uu_new, vv_new, drr_new, dmm_new)
    else:
        res = eng.column_step(p, state, slab[8, :n], slab[9, :n], uu, vv, gd, reduce_fn=all_reduce_sum, exchange=exchange)
    outs = [lprop._pinned_empty(eng, t.numel()) for t in res]
    for o, t in zip(outs, res):
        torch.from_numpy(o).copy_(t, non_blocking=True)
    err = eng.column_work(p.G)[int(_cabi.lib.msgwam_column_error_offset(p.G))]
    err_host = err.to("cpu", non_blocking=False)          # synchronises the stream: the copies above have landed
    if float(err_host) != 0.0:
        err.zero_()
        raise _cabi.MsgwamError("a bounded device-side wait timed out (code %g): a peer did not deliver its deposit, or the "
                                "mean-flow slices did not arrive; the step's results are invalid" % float(err_host))
    u = lprop._unchanged
    return lprop._pack11([u(host[0]), u(host[1]), u(host[2]), outs[0], outs[4] if prof is not None else u(host[4]), u(host[5]),
                          u(host[6]), outs[1], outs[5] if prof is not None else u(host[8]), outs[2], outs[3]])


def shard_range(n: int, rank: int, world: int):
    """Contiguous ray-index range [begin, end) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


# ---- load re-balancing after deletion (no counterpart in the reference) -------------------------
def rebalance_plan(counts):
    """Transfers that even out per-rank ray counts: a list of (src, dst, n) with every rank ending within one ray of
    the mean.  Deterministic in `counts`, so every rank computes the same plan without communication.  Surplus ranks
    give away rays from the END of their store, in rank order, to the deficit ranks in rank order."""
    counts = [int(c) for c in counts]
    world, total = len(counts), sum(counts)
    base, rem = divmod(total, world)
    target = [base + (1 if r < rem else 0) for r in range(world)]
    surplus = [(r, counts[r] - target[r]) for r in range(world) if counts[r] > target[r]]
    deficit = [(r, target[r] - counts[r]) for r in range(world) if counts[r] < target[r]]
    plan, di = [], 0
    for src, s in surplus:
        while s > 0:
            dst, d = deficit[di]
            n = min(s, d)
            plan.append((src, dst, n))
            s -= n; d -= n
            deficit[di] = (dst, d)
            if d == 0:
                di += 1
    return plan, target


def exchange_rows(rows, count, plan, rank, dist=None):
    """Carry out `plan` on a (count, nfields) row-major tensor `rows` of this rank (CPU with gloo, CUDA with NCCL):
    returns (kept_rows_view, received_rows) -- the rays this rank keeps (a prefix of `rows`) and the ones it receives."""
    import torch
    if dist is None:
        import torch.distributed as dist
    nf = rows.shape[1]
    send = [(dst, n) for src, dst, n in plan if src == rank]
    recv = [(src, n) for src, dst, n in plan if dst == rank]
    give = sum(n for _, n in send)
    keep = count - give
    got = torch.empty((sum(n for _, n in recv), nf), dtype=rows.dtype, device=rows.device)
    ops, off = [], keep
    for dst, n in send:                                   # the tail of the store goes out, in plan order
        ops.append(dist.P2POp(dist.isend, rows[off:off + n].contiguous(), dst))
        off += n
    off = 0
    for src, n in recv:
        ops.append(dist.P2POp(dist.irecv, got[off:off + n], src))
        off += n
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return rows[:keep], got
