// Dependent-chain latencies of the FP64 instructions the pencil solve's recurrence is made of (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int sel)
{
    double acc = a + threadIdx.x, acc2 = a * 2 + threadIdx.x, x = b;
    long long t0, t1;
    // 1. DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) acc = fma(-x, acc, acc);
    t1 = clock64(); cyc[0] = t1 - t0;
    // 2. DFMA + select chain (as in solve_chunk_flat)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) { double t = fma(-x, b, acc); acc = ((sel >> (i & 15)) & 1) ? b : t; }
    t1 = clock64(); cyc[1] = t1 - t0;
    // 3. two interleaved DFMA chains
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) { acc = fma(-x, acc, acc); acc2 = fma(-x, acc2, acc2); }
    t1 = clock64(); cyc[2] = t1 - t0;
    // 4. DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) acc = acc + x;
    t1 = clock64(); cyc[3] = t1 - t0;
    // 5. FFMA chain for reference
    float f = (float)a;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) f = fmaf(-(float)x, f, f);
    t1 = clock64(); cyc[4] = t1 - t0;
    out[threadIdx.x] = acc + acc2 + f;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.0, 1e-9, 0x8000); cudaDeviceSynchronize(); }
    printf("DFMA chain %.1f cyc/op | DFMA+select %.1f | 2 interleaved DFMA chains %.1f cyc/pair | DADD %.1f | FFMA %.1f\n",
           cyc[0] / 512.0, cyc[1] / 512.0, cyc[2] / 256.0, cyc[3] / 512.0, cyc[4] / 512.0);
    // throughput: many warps
    return 0;
}
