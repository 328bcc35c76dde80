// Dependent-issue latency of FP64 operations on one warp (clock64 around chains of N dependent ops), and of the
// IEEE sqrt + divide pair the Cholesky pivots need.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c);
    }
    long long t1 = clock64();
    double d = seed;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        d = __dadd_rn(d, c); d = __dadd_rn(d, c); d = __dadd_rn(d, c); d = __dadd_rn(d, c);
    }
    long long t2 = clock64();
    double e = seed + 2.0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) e = __ddiv_rn(1.0, __dsqrt_rn(e)) + 2.0;
    long long t3 = clock64();
    double f = seed + 2.0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) f = __dsqrt_rn(f) + 2.0;
    long long t4 = clock64();
    out[threadIdx.x] = a + d + e + f;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 4 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.5); cudaDeviceSynchronize(); }
    printf("dependent DFMA: %.1f cycles\ndependent DADD: %.1f cycles\n1/sqrt(x) (+1 DADD): %.1f cycles\nsqrt(x) (+1 DADD): %.1f cycles\n",
           cyc[0] / 1024.0, cyc[1] / 1024.0, cyc[2] / 256.0, cyc[3] / 256.0);
    return 0;
}
