// deposit.cuh -- binning of ray volumes onto the uniform vertical grid (wave_projection, L:92-163).
//
// Shared-memory fp64 atomicAdd is a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN.64), and in a
// spatially ordered ensemble all 32 lanes of a warp hit the same handful of cells, so per-ray atomics
// would serialise 32-fold.  Instead every warp owns a private window of WIN consecutive cells in
// shared memory with one column per lane ([WIN][32] double2: both flux components side by side).  A
// lane adds its ray's contributions to its own column with a plain 128-bit load/add/store -- no
// atomics, no bank conflicts, no cross-lane traffic.  Only when the cells touched by the warp leave
// the window (every ~100 iterations for an ordered ensemble) or at the end of the sweep is the window
// flushed: one shuffle reduction per cell, then WIN lanes add one cell each to the global deposit with
// fp64 RED operations (fire-and-forget L2 atomics).  Warps whose lanes are scattered over more than WIN
// cells (unordered rays) bypass the window and add to a CTA histogram in shared memory, where collisions
// are then rare; a CTA that used its histogram merges it into the global deposit when it retires.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace mw {

constexpr int WIN_DEFAULT = 8;                 // cells per warp window

template <int WIN>
struct WindowT {
    static constexpr int CELLS = WIN;
    static constexpr int DOUBLES = WIN * 32 * 2;   // shared-memory doubles per warp and deposit target
    double2 *cell;   // this warp's [WIN][32] running sums
    int wb;          // grid cell of slot 0 (meaningful when live)
    int live;        // the window holds sums
};
using Window = WindowT<WIN_DEFAULT>;
constexpr int WIN_DOUBLES = Window::DOUBLES;

template <int WIN>
__device__ __forceinline__ void window_init(WindowT<WIN> &w, double *base)
{
    w.cell = reinterpret_cast<double2 *>(base);
    w.wb = 0; w.live = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < WIN; ++s) w.cell[s * 32 + lane] = make_double2(0.0, 0.0);
}

// warp-collective: column sums of the window go to h0/h1 (global memory in the sweeps).
// Lane l owns slot l % WIN and sums every (32 / WIN)-th column of it straight from shared memory (the
// column index is rotated by the slot so that the slots fall into different banks); the 32 / WIN partial
// sums of a slot meet in a few shuffles and WIN lanes issue one atomic each, concurrently.
template <int WIN>
__device__ __noinline__ void window_flush_slow(double2 *cell, int wb, double *h0, double *h1)
{
    constexpr int GROUPS = 32 / WIN;
    const int lane = threadIdx.x & 31;
    const int s = lane % WIN, g = lane / WIN;
    __syncwarp();
    double ax = 0.0, ay = 0.0;
    if (g < GROUPS) {
        for (int col = g; col < 32; col += GROUPS) {
            const double2 a = cell[s * 32 + ((col + s) & 31)];
            ax += a.x; ay += a.y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 1; k < GROUPS; ++k) {
        const double bx = __shfl_sync(FULL_MASK, ax, (lane + k * WIN) & 31);
        const double by = __shfl_sync(FULL_MASK, ay, (lane + k * WIN) & 31);
        if (lane < WIN) { ax += bx; ay += by; }
    }
#pragma unroll
    for (int s2 = 0; s2 < WIN; ++s2) cell[s2 * 32 + lane] = make_double2(0.0, 0.0);
    __syncwarp();
    if (lane < WIN && (ax != 0.0 || ay != 0.0)) { atomicAdd(h0 + wb + lane, ax); atomicAdd(h1 + wb + lane, ay); }
}

template <int WIN>
__device__ __forceinline__ void window_flush(WindowT<WIN> &w, double *h0, double *h1)
{
    if (!w.live) return;                                        // warp-uniform
    window_flush_slow<WIN>(w.cell, w.wb, h0, h1);
    w.live = 0;
}

// overlap weight of [rl, ru] with cell c times psv (L:157-162): |min(g[c+1], ru) - max(g[c], rl)| / dz * psv
__device__ __forceinline__ double cell_weight(int c, double rl, double ru, double psv, double dz, double rdz,
                                              const double *__restrict__ g)
{
    const double zmin = dmax(g[c], rl), zmax = dmin(g[c + 1], ru);
    return mul(div_inv(fabs(sub(zmax, zmin)), dz, rdz), psv);
}

// Overlap weights of one ray volume [rl, ru] with cells [nlow, nup) times (psv * v0, psv * v1),
// L:156-163:  w = |min(grid[c+1], ru) - max(grid[c], rl)| / dz;  out[c] += w * psv * v.
// `ok` is false for lanes without a ray or with an out-of-domain ray; all 32 lanes must call.
// h0/h1: where window flushes go (the global deposit in the sweeps); sink: where outlier lanes add (see above).
// Outlier sinks: where a lane that does not fit its warp's window adds (cell, x, y): two arrays (a CTA histogram in
// shared memory, or the output itself), one fp64 atomicAdd each; `used`, if given, marks the histogram dirty for the
// merge at the end of the CTA.  (Measured and rejected: both components side by side with ONE 128-bit
// compare-and-swap per cell, ATOMS.CAS.128 -- 2-3x slower than two 64-bit CAS loops; fp64 RED straight to the global
// deposit -- 2.5x slower once the rays are dispersed, 6x for unordered rays; a 96-bit fixed-point histogram fed with
// native 32-bit ATOMS.ADD and carries -- exact and 2.5x the lane-add rate of the CAS loop in isolation
// (tools/micro/smem_atomics.cu), but its conversions and carry chains cost more issue slots than the CAS unit costs
// time: 198 instead of 145 us per step for a dispersed ensemble.)
// Fixed-point mode (scale != 0; the CTA histogram of the fused column sweeps): a contribution x is added as the 64-bit
// integer RN(x * scale), scale a power of two chosen from a bound on the sum of |x| over the CTA's rays so that no
// accumulator can overflow (column_step.cu: deposit bounds).  Shared memory has native 32-bit integer atomics only
// (fp64, f32 and 64-bit integer adds are all compare-and-swap loops, ATOMS.CAST.SPIN), so the 64-bit add is an
// ATOMS.ADD on the low word whose returned old value yields the carry, and one on the high word: 3.9 lane-adds per clock
// per SM against 0.8-1.5 for the fp64 CAS loop (tools/micro/smem_atomics.cu), independent of how the cells collide.
// Sums are exact in fixed point, hence independent of the order of the adds; the quantum 1 / scale is below 2^-58 of
// the bound.  Contributions that do not fit (non-finite, or larger than the bound allows) go to the global deposit as
// fp64 atomics, unscaled.
// high word of q plus the carry out of old + low word: one add with carry-out, one add with carry-in
__device__ __forceinline__ unsigned hi_plus_carry(unsigned old, unsigned lo, long long q)
{
    unsigned h;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(h) : "r"(old), "r"(lo), "r"((unsigned)(q >> 32)));
    return h;
}

struct SplitTargets {
    double *s0, *s1; int *used;
    double scale = 0.0, scale1 = 0.0;      // fixed-point scales of the two components (each has its own bound); 0: fp64 mode;
                                           // negative (-1): the component's bound is exactly zero -- its contributions were
                                           // all exact zeros, which add nothing (non-zero ones go to g0 / g1 in fp64)
    double *g0 = nullptr, *g1 = nullptr;   // fixed-point mode: the global deposit rows, for contributions that do not fit
    float lim = 1.0e15f;                   // fixed-point mode: a thread adds to the histogram while its running sums stay below
    int mode = -1;                         // fx_mode(scale, scale1), worked out once per CTA by the column sweeps (-1: look at the scales)
    __device__ __forceinline__ void mark() const { if (used != nullptr) *used = 1; }
    // 0: fp64 histogram; 1: both components in fixed point; 2: fixed point with a component whose bound is exactly zero
    static __device__ __forceinline__ int fx_mode(double sc, double sc1) { return sc == 0.0 ? 0 : (sc < 0.0 || sc1 < 0.0) ? 2 : 1; }
    __device__ __forceinline__ int get_mode() const { return mode >= 0 ? mode : fx_mode(scale, scale1); }
    // both components of up to two cells in fixed point: the four low-word adds first, then the carries and high words
    // one component of up to two cells (the other component is identically zero, see `scale`)
    __device__ __forceinline__ void add2_fixed_one(double *row, int c0, double x0, bool two, int c1, double x1) const
    {
        const long long q0 = __double2ll_rn(x0), q2 = __double2ll_rn(x1);
        unsigned *w0 = reinterpret_cast<unsigned *>(row + c0), *w2 = reinterpret_cast<unsigned *>(row + c1);
        const unsigned l0 = (unsigned)q0, l2 = (unsigned)q2;
        const unsigned o0 = atomicAdd(w0, l0);
        unsigned o2 = 0;
        if (two) o2 = atomicAdd(w2, l2);
        atomicAdd(w0 + 1, hi_plus_carry(o0, l0, q0));
        if (two) atomicAdd(w2 + 1, hi_plus_carry(o2, l2, q2));
    }
    __device__ __forceinline__ void add2_fixed(int c0, double x0, double y0, bool two, int c1, double x1, double y1) const
    {
        const long long q0 = __double2ll_rn(x0), q1 = __double2ll_rn(y0), q2 = __double2ll_rn(x1), q3 = __double2ll_rn(y1);
        unsigned *w0 = reinterpret_cast<unsigned *>(s0 + c0), *w1 = reinterpret_cast<unsigned *>(s1 + c0);
        unsigned *w2 = reinterpret_cast<unsigned *>(s0 + c1), *w3 = reinterpret_cast<unsigned *>(s1 + c1);
        const unsigned l0 = (unsigned)q0, l1 = (unsigned)q1, l2 = (unsigned)q2, l3 = (unsigned)q3;
        const unsigned o0 = atomicAdd(w0, l0), o1 = atomicAdd(w1, l1);
        unsigned o2 = 0, o3 = 0;
        if (two) { o2 = atomicAdd(w2, l2); o3 = atomicAdd(w3, l3); }
        const unsigned h0 = hi_plus_carry(o0, l0, q0), h1 = hi_plus_carry(o1, l1, q1);
        const unsigned h2 = hi_plus_carry(o2, l2, q2), h3 = hi_plus_carry(o3, l3, q3);
        // the high words are added unconditionally: with contributions of ~2^40 units a zero high word is rare, and
        // a branch around each add costs four instructions (ISETP, BSSY, BRA, BSYNC)
        atomicAdd(w0 + 1, h0);
        atomicAdd(w1 + 1, h1);
        if (two) { atomicAdd(w2 + 1, h2); atomicAdd(w3 + 1, h3); }
    }
    __device__ __forceinline__ void add(int c, double x, double y) const { atomicAdd(s0 + c, x); atomicAdd(s1 + c, y); }
    // Two cells at once, optimistically: the four read-add-CAS sequences are issued side by side so that their
    // shared-memory round trips overlap (atomicAdd's own CAS loop serialises them); a CAS that lost against another
    // lane falls back to atomicAdd.  c1 is ignored unless `two`.
    __device__ __forceinline__ void add2(int c0, double x0, double y0, bool two, int c1, double x1, double y1) const
    {
        typedef unsigned long long u64;
        u64 *p0 = reinterpret_cast<u64 *>(s0 + c0), *p1 = reinterpret_cast<u64 *>(s1 + c0);
        u64 *p2 = reinterpret_cast<u64 *>(s0 + c1), *p3 = reinterpret_cast<u64 *>(s1 + c1);
        const u64 o0 = *reinterpret_cast<volatile u64 *>(p0), o1 = *reinterpret_cast<volatile u64 *>(p1);
        u64 o2 = 0, o3 = 0;
        if (two) { o2 = *reinterpret_cast<volatile u64 *>(p2); o3 = *reinterpret_cast<volatile u64 *>(p3); }
        const u64 n0 = __double_as_longlong(mw::add(__longlong_as_double(o0), x0));
        const u64 n1 = __double_as_longlong(mw::add(__longlong_as_double(o1), y0));
        const u64 n2 = __double_as_longlong(mw::add(__longlong_as_double(o2), x1));
        const u64 n3 = __double_as_longlong(mw::add(__longlong_as_double(o3), y1));
        const u64 g0 = atomicCAS(p0, o0, n0), g1 = atomicCAS(p1, o1, n1);
        u64 g2 = o2, g3 = o3;
        if (two) { g2 = atomicCAS(p2, o2, n2); g3 = atomicCAS(p3, o3, n3); }
        if (g0 != o0) atomicAdd(s0 + c0, x0);
        if (g1 != o1) atomicAdd(s1 + c0, y0);
        if (g2 != o2) atomicAdd(s0 + c1, x1);
        if (g3 != o3) atomicAdd(s1 + c1, y1);
    }
};

// one ray volume's contributions to the fixed-point CTA histogram: w0, w1 = the two flux components times their scales
// (scaling by a power of two commutes with the rounding of t * v, so v is scaled once per ray); `fits`: the ray's
// contributions are representable, else they go to the global deposit as fp64 atomics, unscaled
template <class Sink>
__device__ __forceinline__ void deposit_fixed(int nlow, int nup, double rl, double ru, double psv, double v0, double v1,
                                              double w0, double w1, bool fits,
                                              double dz, double rdz, const double *__restrict__ g, const Sink &sink)
{
    if (fits) {
        if (sink.mode < 0) sink.mark();                         // the column sweeps mark the histogram once per CTA
        if (sink.get_mode() == 2) {                             // CTA-uniform: a component whose bound is exactly zero
            // Its contributions were all exact zeros in the previous step; exact zeros add nothing, and should a ray
            // bring a non-zero one now (it re-entered the deposit domain, say) it goes to the global deposit in fp64.
            const bool zx = sink.scale < 0.0, zy = sink.scale1 < 0.0;
            if ((zx && v0 != 0.0) || (zy && v1 != 0.0)) {
                for (int c = nlow; c < nup; ++c) {
                    const double t0 = cell_weight(c, rl, ru, psv, dz, rdz, g);
                    if (zx && v0 != 0.0) atomicAdd(sink.g0 + c, mul(t0, v0));
                    if (zy && v1 != 0.0) atomicAdd(sink.g1 + c, mul(t0, v1));
                }
            }
            if (zx && zy) return;
            double *row = zy ? sink.s0 : sink.s1;
            const double w = zy ? w0 : w1;
            for (int c = nlow; c < nup; c += 2) {                  // see the two-component loop below
                const bool two = c + 1 < nup;
                const double ga = g[c], gb = g[c + 1], gc = g[c + 2];
                const double t0 = mul(div_inv(fabs(sub(dmin(gb, ru), dmax(ga, rl))), dz, rdz), psv);
                const double t1 = mul(div_inv(fabs(sub(dmin(gc, ru), dmax(gb, rl))), dz, rdz), psv);
                sink.add2_fixed_one(row, c, mul(t0, w), two, c + 1, mul(t1, w));
            }
            return;
        }
        for (int c = nlow; c < nup; c += 2) {
            // cells c and c + 1 share the node between them; without a second cell (`two` false: its adds are
            // predicated off) the weight of [g[c + 1], g[c + 2]] is computed and dropped (c + 2 <= nup + 1 <= G - 1: cell_range
            // keeps nup <= len(grids) - 2, the top cell is never written, L:134)
            const bool two = c + 1 < nup;
            const double ga = g[c], gb = g[c + 1], gc = g[c + 2];
            const double t0 = mul(div_inv(fabs(sub(dmin(gb, ru), dmax(ga, rl))), dz, rdz), psv);
            const double t1 = mul(div_inv(fabs(sub(dmin(gc, ru), dmax(gb, rl))), dz, rdz), psv);
            sink.add2_fixed(c, mul(t0, w0), mul(t0, w1), two, c + 1, mul(t1, w0), mul(t1, w1));
        }
    } else {                                            // non-finite or outsized: fp64 atomics on the global deposit
        for (int c = nlow; c < nup; ++c) {
            const double t0 = cell_weight(c, rl, ru, psv, dz, rdz, g);
            atomicAdd(sink.g0 + c, mul(t0, v0)); atomicAdd(sink.g1 + c, mul(t0, v1));
        }
    }
}

// Window-free deposit of the fused column sweeps: every lane adds its ray volume's overlap weights (L:156-163) to the CTA
// histogram -- in fixed point when the deposit bound is known (the normal case), else with fp64 compare-and-swap atomics
// (correct for any input, slow when the lanes of a warp share cells).  In fixed point the result does not depend on
// the order of the rays, and the cost barely does: measured per step at 1e7 rays, 0.63 ms for a dispersed or shuffled
// ensemble against 0.79 ms when all 32 lanes of every warp hit the same cells (an exactly ordered ensemble).
// (Measured and removed: summing a coherent warp's contributions per cell with integer warp reductions -- REDUX.SUM on
// three 21-bit limbs, exact -- before two lanes add the totals: 6 REDUX per cell cost more than the 32 colliding atomics
// they replace; a 5e7-ray pile-up in 67 cells ran at 10.7 ms per step that way and runs at 3.2 ms without.)
// bx, by: this thread's running sums of the scaled |contributions| of the two components -- what the next step's deposit
// bounds are made of (column_step.cu: publish_bounds); single precision is plenty for a bound with a factor 32 of headroom.
template <class Sink>
__device__ __forceinline__ void deposit_direct(bool ok, int nlow, int nup, double rl, double ru, double psv, double v0,
                                               double v1, double dz, double rdz, const double *__restrict__ g,
                                               const Sink &sink, float &bx, float &by)
{
    ok = ok && (nup > nlow);
    const int mode = sink.get_mode();
    if (mode != 0) {                                            // CTA-uniform
        // A cell weight is at most psv (1 + 2^-52), so f0, f1 bound what this ray adds to any cell.  The thread's running
        // sums bx, by of them double as an overflow guard: a ray goes to the histogram only while both are below
        // sink.lim = 2^62 / (threads per CTA) -- then no cell of the CTA histogram can reach 2^63, whatever has become of
        // the bounds the scales were derived from; with accurate bounds the average thread ends a sweep a factor 64 below the limit.
        // Past it (stale bounds, a non-finite ray) contributions go to the global deposit in fp64.
        const double w0 = mul(v0, sink.scale), w1 = mul(v1, sink.scale1);
        const double f0 = mul(psv, fabs(w0)), f1 = mul(psv, fabs(w1));
        if (ok) {
            bx += __double2float_ru(f0); by += __double2float_ru(f1);
            const bool fits = mode == 1 ? (bx < sink.lim && by < sink.lim)
                                        : (sink.scale < 0.0 || bx < sink.lim) && (sink.scale1 < 0.0 || by < sink.lim);
            deposit_fixed(nlow, nup, rl, ru, psv, v0, v1, w0, w1, fits, dz, rdz, g, sink);
        }
        return;
    }
    if (!ok) return;
    sink.mark();
    for (int c = nlow; c < nup; c += 2) {
        const bool two = c + 1 < nup;
        const int c1 = two ? c + 1 : c;
        const double t0 = cell_weight(c, rl, ru, psv, dz, rdz, g), t1 = cell_weight(c1, rl, ru, psv, dz, rdz, g);
        sink.add2(c, mul(t0, v0), mul(t0, v1), two, c1, mul(t1, v0), mul(t1, v1));
    }
}

template <int WIN, class Sink>
__device__ __forceinline__ void deposit_cells(bool ok, int nlow, int nup, double rl, double ru,
                                              double psv, double v0, double v1,
                                              double dz, double rdz, const double *__restrict__ g,
                                              WindowT<WIN> &w, double *h0, double *h1, const Sink &sink)
{
    ok = ok && (nup > nlow);
    const int lo = __reduce_min_sync(FULL_MASK, ok ? nlow : INT_MAX);
    if (lo == INT_MAX) return;                                  // no lane has anything to deposit
    // (issuing this second reduction ahead of the branch costs the N(z) sweeps 17 %)
    const int hi = __reduce_max_sync(FULL_MASK, ok ? nup : INT_MIN);
    const bool fits = (hi - lo) <= WIN;
    bool inw = ok;
    if (fits) {
        if (w.live && (lo < w.wb || hi > w.wb + WIN)) window_flush(w, h0, h1);
        if (!w.live) { w.wb = lo; w.live = 1; }
    } else {
        // The warp as a whole does not fit: at r0 its rays sit within metres of each other, but by the later RK
        // states a few fast ones have run several cells ahead (0.7 % of the warp iterations of pass B in the
        // benchmark ensemble).  The window keeps serving the majority and only the lanes outside it add to the CTA
        // histogram.  (Sending the whole warp there cost ~1e4 cycles per event -- 32 lanes in a CAS loop on the same
        // few cells -- and those events decided when a CTA finished.)  Unordered rays: nearly every lane is outside.
        const unsigned m_ok = __ballot_sync(FULL_MASK, ok);
        if (w.live) {
            inw = ok && nlow >= w.wb && nup <= w.wb + WIN;
            if (2 * __popc(__ballot_sync(FULL_MASK, inw)) < __popc(m_ok)) window_flush(w, h0, h1);
        }
        if (!w.live) {
            // anchor at the low end or at the high end, whichever serves more lanes (none if that is under half)
            const int n_lo = __popc(__ballot_sync(FULL_MASK, ok && nup <= lo + WIN));
            const int n_hi = __popc(__ballot_sync(FULL_MASK, ok && nlow >= hi - WIN));
            const int wb = n_lo >= n_hi ? lo : hi - WIN;
            inw = false;
            if (2 * max(n_lo, n_hi) >= __popc(m_ok)) {
                w.wb = wb; w.live = 1;
                inw = ok && nlow >= wb && nup <= wb + WIN;
            }
        }
    }
    if (inw) {
        // private column of the warp window: plain load / add / store.  The first four cells are
        // unrolled so that their weight computations overlap (a ray volume rarely spans more).
        double2 *col = w.cell + (nlow - w.wb) * 32 + (threadIdx.x & 31);
        double t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = cell_weight(min(nlow + k, nup - 1), rl, ru, psv, dz, rdz, g);   // clamped: no selects
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (nlow + k < nup) {
                double2 a = col[k * 32];
                a.x = add(a.x, mul(t[k], v0)); a.y = add(a.y, mul(t[k], v1));
                col[k * 32] = a;
            }
        }
        for (int c = nlow + 4; c < nup; ++c) {
            const double tc = cell_weight(c, rl, ru, psv, dz, rdz, g);
            double2 a = col[(c - nlow) * 32];
            a.x = add(a.x, mul(tc, v0)); a.y = add(a.y, mul(tc, v1));
            col[(c - nlow) * 32] = a;
        }
    } else if (ok) {
        // outlier lane / unordered rays: add to the CTA histogram directly
        if (sink.scale != 0.0) {
            const double w0 = mul(v0, sink.scale), w1 = mul(v1, sink.scale1);
            deposit_fixed(nlow, nup, rl, ru, psv, v0, v1, w0, w1, mul(psv, add(fabs(w0), fabs(w1))) < sink.lim, dz, rdz, g, sink);
            return;
        }
        sink.mark();
        for (int c = nlow; c < nup; c += 2) {
            const bool two = c + 1 < nup;
            const int c1 = two ? c + 1 : c;
            const double t0 = cell_weight(c, rl, ru, psv, dz, rdz, g), t1 = cell_weight(c1, rl, ru, psv, dz, rdz, g);
            sink.add2(c, mul(t0, v0), mul(t0, v1), two, c1, mul(t1, v0), mul(t1, v1));
        }
    }
}

}  // namespace mw
