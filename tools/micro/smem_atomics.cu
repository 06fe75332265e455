// Developer microbenchmark: throughput of shared-memory scatter-add flavours on one SM (not part of the product).
// 768 threads, a 2048-entry table, pseudo-random cells; cycles per warp-level operation (32 lane-adds).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ unsigned rnd(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(768, 1) k(double *out, long long *cyc, int iters, int span)
{
    __shared__ __align__(16) double hd[2048];
    unsigned *hu = reinterpret_cast<unsigned *>(hd);
    for (int j = threadIdx.x; j < 2048; j += blockDim.x) hd[j] = 0.0;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + 12345u;
    const int base = (threadIdx.x >> 5) * 7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const int c = (base + (int)(rnd(s) % (unsigned)span)) & 2047;
        const double x = 1.0 + (s & 255) * 1e-3;
        if (MODE == 0) atomicAdd(hd + c, x);                                  // fp64 CAS loop
        if (MODE == 1) atomicAdd(hu + c, (unsigned)(s & 1023));               // native u32 add, no return
        if (MODE == 2) { unsigned o = atomicAdd(hu + 2 * (c & 1023), (unsigned)s); if (o + (unsigned)s < o) atomicAdd(hu + 2 * (c & 1023) + 1, 1u); }  // 64-bit as lo + carry
        if (MODE == 3) { double o = hd[c]; hd[c] = o + x; }                   // plain (racy) load-add-store: the floor
        if (MODE == 4) {                                                      // 96-bit fixed point: lo, mid, hi words
            const long long v = __double2ll_rn(x * 1099511627776.0);
            unsigned *w = hu + 3 * (c % 1365);
            const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
            const unsigned o = atomicAdd(w, lo);
            const unsigned cy = (o + lo < o);
            const unsigned o2 = atomicAdd(w + 1, hi + cy);
            const unsigned add2 = (unsigned)(v >> 63) + (unsigned)((o2 + hi + cy < o2) | ((hi + cy) < hi));
            if (add2) atomicAdd(w + 2, add2);
        }
        if (MODE == 5) { u64 *p = reinterpret_cast<u64 *>(hd + c); u64 o = *(volatile u64 *)p; u64 n = __double_as_longlong(__longlong_as_double(o) + x);
                         if (atomicCAS(p, o, n) != o) atomicAdd(hd + c, x); }  // optimistic CAS
    }
    long long t1 = clock64();
    __syncthreads();
    double acc = 0; for (int j = threadIdx.x; j < 2048; j += blockDim.x) acc += hd[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int it = 4096;
    const char *names[] = {"fp64 atomicAdd (CAS loop)", "u32 atomicAdd no return", "u32 lo + carry (64-bit int)", "plain ld/add/st (racy floor)", "96-bit fixed point, 2-3 u32 atomics", "optimistic 64-bit CAS"};
    for (int span : {2000, 150, 24, 8}) {
#define RUN(M) k<M><<<1, 768>>>(out, cyc, it, span); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("span %4d  %-38s %7.1f cycles per warp-op per SM (24 warps: %.2f lane-adds/clk)\n", span, names[M], (double)h / it / 24, 32.0 * 24 * it / (double)h);
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
