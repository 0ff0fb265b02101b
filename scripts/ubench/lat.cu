// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/lat scripts/ubench/lat.cu   (run: scripts/ubench/lat on a B200)
// Dependent-chain latencies on sm_100a (cycles per instruction, one warp): DFMA, DADD, DMUL, F2F.F64.F16, 64-bit SHFL + DADD, LDS.64.
#include <cstdio>
#include <cuda_fp16.h>
__global__ void lat(double* out, long long* cyc, double seed, int n)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 32] = seed;
    __syncthreads();
    double a = seed, b = seed * 0.5, c = seed * 0.25;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = a + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = a * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // F2F chain: double -> half -> double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { unsigned short h; asm volatile("cvt.rn.f16.f64 %0, %1;" : "=h"(h) : "d"(a)); asm volatile("cvt.f64.f16 %0, %1;" : "=d"(a) : "h"(h)); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // SHFL(64-bit) + DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // LDS.64 pointer chase
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { double v = sm[idx]; idx = (__double2loint(v) + idx) & 63; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 2 independent DFMA chains (ILP 2), 4 chains (ILP 4)
    double x0 = a, x1 = b, x2 = c, x3 = seed;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = a + idx + x0 + x1 + x2 + x3;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64);
    const int n = 4096;
    for (int warps = 1; warps <= 4; warps *= 2) {
        lat<<<1, 32 * warps>>>(out, cyc, 1e-300, n);   // warps on ONE SM; 1 warp per SMSP up to 4
        lat<<<1, 32 * warps>>>(out, cyc, 1e-300, n);
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("warps %d: DFMA %.2f DADD %.2f DMUL %.2f F2Fx2 %.2f SHFL64+DADD %.2f LDS64chase %.2f | DFMA ilp2 %.2f/inst ilp4 %.2f/inst (cycles per iteration: %.2f %.2f)\n", warps,
               h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[5] / (double)n, h[6] / (2.0 * n), h[7] / (4.0 * n), h[6] / (double)n, h[7] / (double)n);
    }
    // 8 and 16 warps on one SM (2 and 4 per SMSP)
    for (int warps = 8; warps <= 16; warps *= 2) {
        lat<<<1, 32 * warps>>>(out, cyc, 1e-300, n);
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("warps %d (warp 0's view): DFMA %.2f DADD %.2f F2Fx2 %.2f SHFL64+DADD %.2f ilp4 %.2f/inst\n", warps, h[0] / (double)n, h[1] / (double)n, h[3] / (double)n, h[4] / (double)n, h[7] / (4.0 * n));
    }
    return 0;
}
