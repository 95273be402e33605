// Dependent-issue latencies of the instructions the per-scenario kernels chain (one warp, clock64): NOT product code.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0, int n) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = x0 + threadIdx.x;
    __syncthreads();
    double x = x0 + threadIdx.x * 1e-3, y = 1.0000001;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL+DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // division
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; ++i) x = 1.0 / (x + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) * 4;
    // shuffle of a double (2 SHFL)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // REDUX max
    unsigned u = __double2uint_rn(x * 1000.0) + threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) u = __reduce_max_sync(0xffffffffu, u + threadIdx.x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // LDS dependent (pointer chase through doubles)
    int idx = threadIdx.x & 63;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) idx = ((int)sm[idx]) & 63;
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // DSETP + FSEL select chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) x = (x > y) ? y : x + 1e-300;
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // integer add chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) u = u * 3 + idx;
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 8 independent DFMA chains (throughput of one warp)
    double z[8];
    for (int j = 0; j < 8; ++j) z[j] = x + j;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fma(z[j], y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0) / 2;
    for (int j = 0; j < 8; ++j) x += z[j];
    out[threadIdx.x] = x + u + idx;
}
int main() {
    double* o; long long* c;
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 16 * 8);
    k<<<1, 32>>>(o, c, 1.25, 1024);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, c, 16 * 8, cudaMemcpyDeviceToHost);
    const char* nm[9] = {"DFMA dependent", "DMUL dependent", "1.0/x (+DADD)", "SHFL double (2 SHFL)", "REDUX.MAX (+IADD)", "LDS->I2F chase",
                         "DSETP+FSEL (+DADD)", "IMAD chain", "DFMA 8 independent chains: cycles per DFMA x4"};
    for (int i = 0; i < 9; ++i) printf("%-46s %.1f cycles per op\n", nm[i], h[i] / 1024.0);
    return 0;
}
