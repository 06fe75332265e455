// compact.cu -- ray deletion by stable stream compaction of the structure-of-arrays ray store.
//
// The reference never deletes rays; its only related predicate is `out_of_domain` in
// wave_projection (L:129-130).  Deletion is an explicit step between RK steps:
// msgwam_flag_rays marks survivors (inside the deposit domain and |m| below the critical-level
// cut-off m_crit), msgwam_compact packs every field of the store with the same permutation.
// Three kernels: per-tile survivor counts -> exclusive scan of the tile counts -> scatter.
#include "common.cuh"

namespace {

using namespace mw;

constexpr int CT = 256;            // threads per CTA
constexpr int ITEMS = 16;          // rays per thread; tile = 4096 rays
constexpr int TILE = CT * ITEMS;

__global__ void __launch_bounds__(CT) flag_kernel(msgwam_params_t p, int64_t n, const double *__restrict__ rr,
                                                  const double *__restrict__ drr, const double *__restrict__ mm,
                                                  double m_crit, uint8_t *__restrict__ keep)
{
    for (int64_t i = (int64_t)blockIdx.x * CT + threadIdx.x; i < n; i += (int64_t)gridDim.x * CT) {
        const double hd = mul(.5, drr[i]);
        int nlow, nup;
        const bool inside = cell_range(sub(rr[i], hd), add(rr[i], hd), p.dz_grids, p.inv_dz_grids, p.G - 2, nlow, nup);
        keep[i] = (inside && fabs(mm[i]) < m_crit) ? 1 : 0;
    }
}

// tile t covers rays [t*TILE, (t+1)*TILE); thread k of the tile owns the contiguous run
// [t*TILE + k*ITEMS, ... + ITEMS) so that the packing is stable.
__global__ void __launch_bounds__(CT) count_kernel(int64_t n, const uint8_t *__restrict__ keep, int64_t *__restrict__ tile_count)
{
    __shared__ int warp_tot[CT / 32];
    const int64_t start = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    int c = 0;
    if (start + ITEMS <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(keep + start);     // 16 flags, 16-byte aligned
        // a flag byte counts when it is non-zero, exactly as the scalar tail and scatter_kernel read it
        c = __popc(__vcmpne4(v.x, 0u) & 0x01010101u) + __popc(__vcmpne4(v.y, 0u) & 0x01010101u) +
            __popc(__vcmpne4(v.z, 0u) & 0x01010101u) + __popc(__vcmpne4(v.w, 0u) & 0x01010101u);
    } else {
        for (int k = 0; k < ITEMS; ++k) if (start + k < n && keep[start + k]) ++c;
    }
    c = __reduce_add_sync(FULL_MASK, c);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CT / 32; ++w) t += warp_tot[w];
        tile_count[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the tile counts in place, total to *count
__global__ void __launch_bounds__(1024) scan_kernel(int64_t ntiles, int64_t *__restrict__ tile_count, int64_t *__restrict__ count)
{
    __shared__ int64_t part[1024];
    const int64_t per = (ntiles + 1023) / 1024;
    const int64_t b = (int64_t)threadIdx.x * per, e = (b + per < ntiles) ? b + per : ntiles;
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i) s += tile_count[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int k = 0; k < 1024; ++k) { const int64_t v = part[k]; part[k] = run; run += v; }
        *count = run;
    }
    __syncthreads();
    int64_t run = part[threadIdx.x];
    for (int64_t i = b; i < e; ++i) { const int64_t v = tile_count[i]; tile_count[i] = run; run += v; }
}

struct ScatterArgs {
    int64_t n;
    const uint8_t *keep;
    const int64_t *tile_offset;
    int nfields;
    const double *in[16];
    double *out[16];
};

// Stable packing of one tile with coalesced accesses: in round k the CTA looks at the 256 consecutive rays
// [k * CT, (k + 1) * CT) of the tile, so a warp always reads 32 neighbouring rays of every field (256 contiguous bytes)
// and writes its survivors to neighbouring slots.  The tile's 128 warp-segments (16 rounds x 8 warps, in ray order)
// are counted with ballots first; one warp scans the 128 counts; then every lane knows its destination.
// (The first version gave each thread 16 consecutive rays: 128-byte strides between lanes, 534 GB/s at 5e7 rays.)
__global__ void __launch_bounds__(CT) scatter_kernel(const ScatterArgs a)
{
    constexpr int NWARP = CT / 32, NSEG = ITEMS * NWARP;          // 8 warps, 128 segments of 32 rays
    __shared__ int seg_off[NSEG];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    unsigned mine = 0;                                            // bit k: my ray of round k survives
    unsigned ballots[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = tile0 + k * CT + threadIdx.x;
        const bool keep = i < a.n && a.keep[i];
        ballots[k] = __ballot_sync(FULL_MASK, keep);
        if (keep) mine |= 1u << k;
        if (lane == 0) seg_off[k * NWARP + w] = __popc(ballots[k]);
    }
    __syncthreads();
    if (w == 0) {                                                 // exclusive scan of the 128 segment counts
        int v[NSEG / 32], sum = 0;
#pragma unroll
        for (int j = 0; j < NSEG / 32; ++j) { v[j] = seg_off[lane * (NSEG / 32) + j]; sum += v[j]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
        int run = incl - sum;
#pragma unroll
        for (int j = 0; j < NSEG / 32; ++j) { seg_off[lane * (NSEG / 32) + j] = run; run += v[j]; }
    }
    __syncthreads();
    const int64_t base = a.tile_offset[blockIdx.x];
    const unsigned below = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (mine & (1u << k)) {
            const int64_t src = tile0 + k * CT + threadIdx.x;
            const int64_t dst = base + seg_off[k * NWARP + w] + __popc(ballots[k] & below);
            for (int f = 0; f < a.nfields; ++f) a.out[f][dst] = __ldcs(a.in[f] + src);
        }
    }
}

}  // namespace

extern "C" {

int64_t msgwam_compact_scratch_bytes(int64_t n)
{
    const int64_t ntiles = (n + TILE - 1) / TILE;
    return (ntiles + 1) * (int64_t)sizeof(int64_t);
}

int msgwam_flag_rays(const msgwam_params_t *p, int64_t n, const double *d_rr, const double *d_drr, const double *d_mm,
                     double m_crit, uint8_t *d_keep, void *stream)
{
    if (!p || n < 0 || p->G < 3) return MSGWAM_E_BADARG;
    if (n == 0) return 0;
    if (!d_rr || !d_drr || !d_mm || !d_keep) return MSGWAM_E_BADARG;
    int sms = 0;
    int rc = msgwam_device_info(&sms, nullptr);
    if (rc) return rc;
    int64_t blocks = (n + CT - 1) / CT;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    flag_kernel<<<(int)blocks, CT, 0, (cudaStream_t)stream>>>(*p, n, d_rr, d_drr, d_mm, m_crit, d_keep);
    return (int)cudaGetLastError();
}

int msgwam_compact(int64_t n, const uint8_t *d_keep, int32_t nfields, const double *const d_in[], double *const d_out[],
                   int64_t *d_count, void *d_scratch, void *stream)
{
    if (n < 0 || nfields < 0 || nfields > 16 || !d_count) return MSGWAM_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(d_count, 0, sizeof(int64_t), s);
    if (!d_keep || !d_scratch || (nfields > 0 && (!d_in || !d_out))) return MSGWAM_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(d_keep) & 15) != 0) return MSGWAM_E_BADARG;   // flags are read 16 at a time
    const int64_t ntiles = (n + TILE - 1) / TILE;
    int64_t *tile = reinterpret_cast<int64_t *>(d_scratch);
    count_kernel<<<(unsigned)ntiles, CT, 0, s>>>(n, d_keep, tile);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    scan_kernel<<<1, 1024, 0, s>>>(ntiles, tile, d_count);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ScatterArgs a{};
    a.n = n; a.keep = d_keep; a.tile_offset = tile; a.nfields = nfields;
    for (int f = 0; f < nfields; ++f) {
        if (!d_in[f] || !d_out[f] || d_in[f] == d_out[f]) return MSGWAM_E_BADARG;
        a.in[f] = d_in[f]; a.out[f] = d_out[f];
    }
    scatter_kernel<<<(unsigned)ntiles, CT, 0, s>>>(a);
    return (int)cudaGetLastError();
}

}  // extern "C"
