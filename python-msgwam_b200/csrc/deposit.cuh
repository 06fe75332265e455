// deposit.cuh -- binning of ray volumes onto the uniform vertical grid (wave_projection, L:92-163).
//
// Register-resident running sums per lane + warp-collective flush + CTA shared-memory histogram.
// Shared-memory fp64 atomicAdd is a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN.64), so the
// design keeps same-address traffic out of it: lanes of a warp that hit the same cells are
// combined with shuffles first, and only one lane per cell touches shared memory.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace mw {

constexpr int K_SLOTS = 4;       // register cell sums per lane and deposit

struct Acc {
    int base;        // first cell of the window (meaningful when nonempty)
    int nonempty;
    double a0[K_SLOTS], a1[K_SLOTS];
    __device__ __forceinline__ void clear()
    {
        nonempty = 0; base = 0;
#pragma unroll
        for (int j = 0; j < K_SLOTS; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
    }
};

// warp-collective: move every lane's register sums into the histogram h0/h1
// (shared memory in the fused kernels; the same code works on global memory).
__device__ __forceinline__ void flush_acc(Acc &acc, double *h0, double *h1)
{
    const int lo = __reduce_min_sync(FULL_MASK, acc.nonempty ? acc.base : INT_MAX);
    if (lo == INT_MAX) return;                                  // warp-uniform
    const int hi = __reduce_max_sync(FULL_MASK, acc.nonempty ? acc.base + K_SLOTS : INT_MIN);
    const int lane = threadIdx.x & 31;
    if (hi - lo <= 3 * K_SLOTS) {
        // lanes' windows overlap: one shuffle reduction per cell, one lane adds
        for (int c = lo; c < hi; ++c) {
            const int j = c - acc.base;
            double x0 = 0.0, x1 = 0.0;
            if (acc.nonempty) {
#pragma unroll
                for (int jj = 0; jj < K_SLOTS; ++jj)
                    if (j == jj) { x0 = acc.a0[jj]; x1 = acc.a1[jj]; }
            }
            x0 = warp_sum(x0); x1 = warp_sum(x1);
            if (lane == 0 && (x0 != 0.0 || x1 != 0.0)) { atomicAdd(h0 + c, x0); atomicAdd(h1 + c, x1); }
        }
    } else if (acc.nonempty) {
        // scattered lanes (unordered rays): few collisions, add directly
#pragma unroll
        for (int jj = 0; jj < K_SLOTS; ++jj)
            if (acc.a0[jj] != 0.0 || acc.a1[jj] != 0.0) {
                atomicAdd(h0 + acc.base + jj, acc.a0[jj]);
                atomicAdd(h1 + acc.base + jj, acc.a1[jj]);
            }
    }
    acc.clear();
}

// Overlap weights of one ray volume [rl, ru] with cells [nlow, nup) times (psv * v0, psv * v1),
// L:156-163:  w = |min(grid[c+1], ru) - max(grid[c], rl)| / dz;  out[c] += w * psv * v.
// `ok` is false for lanes without a ray or with an out-of-domain ray; all 32 lanes must call.
__device__ __forceinline__ void deposit_cells(bool ok, int nlow, int nup, double rl, double ru,
                                              double psv, double v0, double v1,
                                              double dz, double rdz, const double *__restrict__ g,
                                              Acc &acc, double *h0, double *h1)
{
    ok = ok && (nup > nlow);
    const bool moved = ok && acc.nonempty && (nlow != acc.base);
    if (__any_sync(FULL_MASK, moved)) flush_acc(acc, h0, h1);
    if (ok) {
        acc.base = nlow; acc.nonempty = 1;
#pragma unroll
        for (int j = 0; j < K_SLOTS; ++j) {
            const int c = nlow + j;
            if (c < nup) {
                const double zmin = fmax(g[c], rl), zmax = fmin(g[c + 1], ru);
                const double t = mul(div_inv(fabs(sub(zmax, zmin)), dz, rdz), psv);
                acc.a0[j] = add(acc.a0[j], mul(t, v0));
                acc.a1[j] = add(acc.a1[j], mul(t, v1));
            }
        }
        for (int c = nlow + K_SLOTS; c < nup; ++c) {           // ray volumes taller than K_SLOTS cells
            const double zmin = fmax(g[c], rl), zmax = fmin(g[c + 1], ru);
            const double t = mul(div_inv(fabs(sub(zmax, zmin)), dz, rdz), psv);
            atomicAdd(h0 + c, mul(t, v0));
            atomicAdd(h1 + c, mul(t, v1));
        }
    }
}

}  // namespace mw
