// Developer microbenchmark: fp64 pipe latency / throughput on the device (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void divchain(double *out, long long *cyc, int iters, double b)
{
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __ddiv_rn(x, b) + 1.0;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void sqrtchain(double *out, long long *cyc, int iters)
{
    double x = 2.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __dsqrt_rn(x) + 1.5;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8 * 8); cudaMalloc(&cyc, 8);
    const int it = 4096;
#define RUN(K, blocks, threads, name) chain<K><<<blocks, threads>>>(out, cyc, it, 1.0000001, 1e-9); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-34s %8.2f cycles per dependent DFMA step (ILP %d, %d warps/SM)\n", name, (double)h / it, K, threads / 32 * (blocks >= 148 ? blocks / 148 : 1));
    RUN(1, 1, 32, "1 warp, ILP1") RUN(1, 1, 32, "1 warp, ILP1 (again)") RUN(2, 1, 32, "1 warp, ILP2") RUN(4, 1, 32, "1 warp, ILP4") RUN(8, 1, 32, "1 warp, ILP8")
    RUN(1, 1, 128, "4 warps (1/SMSP), ILP1") RUN(1, 1, 256, "8 warps (2/SMSP), ILP1") RUN(1, 1, 512, "16 warps (4/SMSP), ILP1") RUN(1, 1, 1024, "32 warps (8/SMSP), ILP1")
    RUN(2, 1, 512, "16 warps, ILP2") RUN(4, 1, 512, "16 warps, ILP4") RUN(1, 148, 512, "148 CTAs x 16 warps, ILP1") RUN(2, 148, 1024, "148 CTAs x 32 warps, ILP2")
    divchain<<<1, 32>>>(out, cyc, it, 1.7); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent __ddiv_rn + DADD, 1 warp:   %8.2f cycles per iteration\n", (double)h / it);
    divchain<<<1, 512>>>(out, cyc, it, 1.7); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent __ddiv_rn + DADD, 16 warps: %8.2f cycles per iteration\n", (double)h / it);
    sqrtchain<<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent __dsqrt_rn + DADD, 1 warp:  %8.2f cycles per iteration\n", (double)h / it);
    sqrtchain<<<1, 512>>>(out, cyc, it); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent __dsqrt_rn + DADD, 16 warps:%8.2f cycles per iteration\n", (double)h / it);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
