"""Multi-GPU (one process per GPU, torch.distributed / NCCL) variants of the reference-facing step.

Rays are sharded by contiguous index range across ranks; grid fields are replicated.  The only
exchange is the sum of the deposited pseudo-momentum flux (L:654-658 sums over *all* rays): an
in-place all-reduce of 4*(G-1) doubles after pass A and 2*(G-1) after pass B, enqueued on the
compute stream with no host synchronisation.  Every rank then advances its replica of the mean
flow with bit-identical inputs.
"""
from __future__ import annotations

import numpy as np

from ._engine import Engine


def all_reduce_sum(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)


def rk3_host_sharded(lprop, dt, var):
    """lprop.RK3 for this rank's slice of the rays, given as HOST numpy arrays (column mode only).

    Same contract as lprop.RK3: returns a fresh 11-slot object array of numpy arrays; uu, vv are the
    globally coupled mean flow (identical on every rank).
    """
    eng = Engine.get()
    torch = eng.torch
    p = lprop._params(dt)
    if p.hprop or p.saturate_online:
        raise NotImplementedError("sharded host path covers the column mode (HPROP off, saturate_online off)")
    n = int(np.size(var[3]))
    up = lambda a, m: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (m,)))).to(eng.device, non_blocking=True)
    state = [up(a, n) if i != 1 else None for i, a in enumerate(var[:9])]       # lam is not needed on the device
    state[1] = state[2]
    dkk, dll = up(lprop.statics['dkk'], n), up(lprop.statics['dll'], n)
    uu, vv = up(var[9], p.G), up(var[10], p.G)
    gd = lprop._grid_devs(eng)
    rr_new, mm_new, uu_new, vv_new = eng.column_step(p, state, dkk, dll, uu, vv, gd, reduce_fn=all_reduce_sum)
    host = lambda t: t.cpu().numpy()
    cp = lambda a: np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
    return lprop._pack11([cp(var[0]), cp(var[1]), cp(var[2]), host(rr_new), cp(var[4]), cp(var[5]), cp(var[6]),
                          host(mm_new), cp(var[8]), host(uu_new), host(vv_new)])


def shard_range(n: int, rank: int, world: int):
    """Contiguous ray-index range [begin, end) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
