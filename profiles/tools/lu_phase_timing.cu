// Phase timing of the shared-memory panel LU (hpf_lu_panel.cuh) with clock64(): NOT product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o lu_phase_timing lu_phase_timing.cu
//   ./lu_phase_timing [N] [ctas_per_sm] [reps]
// Every CTA factors `reps` random N x N systems; thread 0 (in the panel warp) and the last thread (a
// warp that waits during the panel steps) accumulate the cycles between the LUP_TICK marks.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__shared__ long long lup_ticks[2][16];
__shared__ long long lup_last[2];
#define LUP_TICK(i)                                                              \
    do {                                                                         \
        if (threadIdx.x == 0 || threadIdx.x == blockDim.x - 1) {                 \
            const int w_ = threadIdx.x == 0 ? 0 : 1;                             \
            const long long t_ = clock64();                                      \
            lup_ticks[w_][i] += t_ - lup_last[w_];                               \
            lup_last[w_] = t_;                                                   \
        }                                                                        \
    } while (0)
#include "../../harmonic_power_flow_b200/csrc/hpf_lu_panel.cuh"

__global__ void __launch_bounds__(256, 2) k(const double* M, int N, int reps, long long* out, double* xout, int classic) {
    extern __shared__ __align__(16) double smem[];
    const int ld = N | 1;
    double* A = smem;
    double* X = A + (size_t)ld * (N + 1);
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0) X += 1;
    double* rinv = X + lup_extra_doubles(N);
    int* flag = reinterpret_cast<int*>(rinv + N + 2);
    if (threadIdx.x < 32) reinterpret_cast<long long*>(lup_ticks)[threadIdx.x] = 0;
    __syncthreads();
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        for (int t = threadIdx.x; t < N * (N + 1); t += blockDim.x) {
            const int col = t / N, row = t - col * N;
            A[row + (size_t)col * ld] = M[((size_t)(blockIdx.x % 7) * (N + 1) + col) * N + row];
        }
        __syncthreads();
        if (threadIdx.x == 0 || threadIdx.x == blockDim.x - 1) lup_last[threadIdx.x == 0 ? 0 : 1] = clock64();
        const long long t0 = clock64();
        if (classic) lu_solve_smem(A, N, ld, rinv, flag);
        else lu_solve_smem_panel(A, N, ld, rinv, flag, X);
        total += clock64() - t0;
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) out[threadIdx.x] = reinterpret_cast<long long*>(lup_ticks)[threadIdx.x];
        if (threadIdx.x == 0) out[32] = total;
        for (int t = threadIdx.x; t < N; t += blockDim.x) xout[t] = A[(size_t)N * ld + t];
    }
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 101;
    const int per_sm = argc > 2 ? atoi(argv[2]) : 2;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    std::vector<double> h((size_t)7 * N * (N + 1));
    srand(1);
    for (auto& v : h) v = rand() / (double)RAND_MAX - 0.5;
    double *dM, *dx;
    long long* dout;
    cudaMalloc(&dM, h.size() * 8); cudaMalloc(&dout, 64 * 8); cudaMalloc(&dx, N * 8);
    cudaMemcpy(dM, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const size_t smem = ((size_t)(N | 1) * (N + 1) + lup_extra_doubles(N) + N + 8) * 8 + 32;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // slot i = cycles between LUP_TICK(i-1) (or the previous tick reached) and LUP_TICK(i)
    const char* names[9] = {"-", "prologue: first panel (w0)", "barrier (+loop top)", "next panel's columns (all warps) + barrier",
                            "w0: P1 of next panel | w1-7: P2+P3 of the rest", "full barrier", "-", "loop exit", "back substitution"};
    for (int classic = 0; classic < 2; ++classic) {
        cudaMemset(dout, 0, 64 * 8);
        k<<<prop.multiProcessorCount * per_sm, 256, smem>>>(dM, N, reps, dout, dx, classic);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long o[64];
        std::vector<double> x(N);
        cudaMemcpy(o, dout, 64 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(x.data(), dx, N * 8, cudaMemcpyDeviceToHost);
        double cs = 0; for (double v : x) cs += v;
        printf("%s N=%d, %d CTAs/SM: %.0f cycles per matrix per CTA (checksum %.12e)\n", classic ? "rank-1 LU" : "panel LU",
               N, per_sm, (double)o[32] / reps, cs);
        if (!classic)
            for (int i = 0; i < 9; ++i)
                printf("   %-26s thread 0: %8.0f   last thread: %8.0f\n", names[i], (double)o[i] / reps, (double)o[16 + i] / reps);
    }
    return 0;
}
