// deposit.cuh -- binning of ray volumes onto the uniform vertical grid (wave_projection, L:92-163).
//
// Shared-memory fp64 atomicAdd is a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN.64), and in a
// spatially ordered ensemble all 32 lanes of a warp hit the same handful of cells, so per-ray atomics
// would serialise 32-fold.  Instead every warp owns a private window of WIN consecutive cells in
// shared memory with one column per lane ([WIN][32] double2: both flux components side by side).  A
// lane adds its ray's contributions to its own column with a plain 128-bit load/add/store -- no
// atomics, no bank conflicts, no cross-lane traffic.  Only when the cells touched by the warp leave
// the window (every ~100 iterations for an ordered ensemble) or at the end of the sweep is the window
// flushed: one shuffle reduction per cell, one lane adds to the CTA histogram.  Warps whose lanes are
// scattered over more than WIN cells (unordered rays) bypass the window and add to the histogram
// directly, where collisions are then rare.  The CTA histogram goes to HBM with one fp64 RED per
// non-zero cell when the CTA retires.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace mw {

constexpr int WIN = 8;                         // cells per warp window
constexpr int WIN_DOUBLES = WIN * 32 * 2;      // shared-memory doubles per warp and deposit target

struct Window {
    double2 *cell;   // this warp's [WIN][32] running sums
    int wb;          // grid cell of slot 0 (meaningful when live)
    int live;        // the window holds sums
};

__device__ __forceinline__ void window_init(Window &w, double *base)
{
    w.cell = reinterpret_cast<double2 *>(base);
    w.wb = 0; w.live = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < WIN; ++s) w.cell[s * 32 + lane] = make_double2(0.0, 0.0);
}

// warp-collective: column sums of the window go to the histogram h0/h1 (shared or global memory)
__device__ __forceinline__ void window_flush(Window &w, double *h0, double *h1)
{
    if (!w.live) return;                                        // warp-uniform
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < WIN; ++s) {
        const double2 a = w.cell[s * 32 + lane];
        w.cell[s * 32 + lane] = make_double2(0.0, 0.0);
        const double x0 = warp_sum(a.x), x1 = warp_sum(a.y);
        if (lane == 0 && (x0 != 0.0 || x1 != 0.0)) { atomicAdd(h0 + w.wb + s, x0); atomicAdd(h1 + w.wb + s, x1); }
    }
    w.live = 0;
}

// Overlap weights of one ray volume [rl, ru] with cells [nlow, nup) times (psv * v0, psv * v1),
// L:156-163:  w = |min(grid[c+1], ru) - max(grid[c], rl)| / dz;  out[c] += w * psv * v.
// `ok` is false for lanes without a ray or with an out-of-domain ray; all 32 lanes must call.
__device__ __forceinline__ void deposit_cells(bool ok, int nlow, int nup, double rl, double ru,
                                              double psv, double v0, double v1,
                                              double dz, double rdz, const double *__restrict__ g,
                                              Window &w, double *h0, double *h1)
{
    ok = ok && (nup > nlow);
    const int lo = __reduce_min_sync(FULL_MASK, ok ? nlow : INT_MAX);
    if (lo == INT_MAX) return;                                  // no lane has anything to deposit
    const int hi = __reduce_max_sync(FULL_MASK, ok ? nup : INT_MIN);
    const bool fits = (hi - lo) <= WIN;
    if (w.live && (!fits || lo < w.wb || hi > w.wb + WIN)) window_flush(w, h0, h1);
    if (fits && !w.live) { w.wb = lo; w.live = 1; }
    if (ok) {
        const int lane = threadIdx.x & 31;
        for (int c = nlow; c < nup; ++c) {
            const double zmin = dmax(g[c], rl), zmax = dmin(g[c + 1], ru);
            const double t = mul(div_inv(fabs(sub(zmax, zmin)), dz, rdz), psv);
            const double t0 = mul(t, v0), t1 = mul(t, v1);
            if (fits) {
                double2 *p = w.cell + (c - w.wb) * 32 + lane;
                double2 a = *p;
                a.x = add(a.x, t0); a.y = add(a.y, t1);
                *p = a;
            } else {
                atomicAdd(h0 + c, t0);
                atomicAdd(h1 + c, t1);
            }
        }
    }
}

}  // namespace mw
