// Developer microbenchmark: cost of cold straight-line code (instruction fetch) vs warm loops.
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct Unroll { static __device__ __forceinline__ void run(float &a, float &b, float &c, float &d) {
    a = a * 1.0001f + b; b = b * 0.9999f + c; c = c * 1.0002f + d; d = d * 0.9998f + a; Unroll<N - 1>::run(a, b, c, d); } };
template <> struct Unroll<0> { static __device__ __forceinline__ void run(float &, float &, float &, float &) {} };

template <int N>
__global__ void straight(float *out, long long *cyc, int reps)
{
    float a = threadIdx.x, b = 1.f, c = 2.f, d = 3.f;
    for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
        Unroll<N>::run(a, b, c, d);          // 4N independent-ish FFMAs, straight-line
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[r] = t1 - t0;
    }
    out[threadIdx.x] = a + b + c + d;
}
int main()
{
    float *out; long long *cyc, h[4];
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    char *flush; cudaMalloc(&flush, 256 << 20);
#define RUN(N, T) cudaMemset(flush, 1, 256 << 20); straight<N><<<1, T>>>(out, cyc, 3); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost); \
    printf("%5d FFMA straight-line, %4d threads: first pass %7lld cycles (%.1f/instr), 2nd %7lld, 3rd %7lld (%.1f/instr)\n", 4 * N, T, h[0], (double)h[0] / (4 * N), h[1], h[2], (double)h[2] / (4 * N));
    RUN(64, 32) RUN(256, 32) RUN(512, 32) RUN(1024, 32) RUN(256, 1024) RUN(512, 1024) RUN(1024, 1024) RUN(1024, 32)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
