// Complex FP64 GEMM on the FP64 tensor cores (mma.sync.m8n8k4 DMMA):  C [M x N] = A [M x K] B [K x N],
// all three row-major, complex128 interleaved (re, im).
//
// This is the Norton contraction of the structured step written as what it is for a whole batch,
//     w_N = W_NL I_N :   [nZ x qH] [qH x B]          (I = I_N - Y_N V, HG:313-323, through A_ZZ^-1),
// i.e. the north star's "tensor cores for the Norton contraction" candidate; the same kernel forms
// the operator products of the structured set-up.  wn_tile_kernel (hpf_structured.cuh) does the same
// product on the CUDA-core FP64 pipe: one broadcast LDG + one LDS.128 per complex MAC (4 DFMA).
//
// CTA tile 64 x 64 complex, 8 warps (4 x 2), warp tile 16 x 32 = 2 x 4 DMMA tiles with separate real
// and imaginary accumulators (64 registers); K in slices of 8 through a 3-stage cp.async pipeline
// (one 16-byte copy per complex element, zero-fill at the edges).  Both operands stay INTERLEAVED in
// shared memory: a fragment element (re, im) is ONE LDS.128 per lane; the row pitches (12 and 66
// complex) make every quarter-warp hit 8 distinct 16-byte bank groups.  Per 4-deep k step a warp
// issues 6 LDS.128 and 32 DMMA:  Cr += Ar Br + (-Ai) Bi,  Ci += Ar Bi + Ai Br.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define ZG_BM 64
#define ZG_BN 64
#define ZG_BK 8
#define ZG_LDA 12                    // complex per row of the A tile (8 + 4 padding)
#define ZG_LDB 66                    // complex per row of the B tile (64 + 2 padding)
#define ZG_STAGES 3
#define ZG_STAGE_ELEMS (ZG_BM * ZG_LDA + ZG_BK * ZG_LDB)
#define ZG_SMEM_BYTES (ZG_STAGES * ZG_STAGE_ELEMS * 16)

__device__ __forceinline__ void zg_dmma(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 16-byte asynchronous copy; `ok` false: the destination is zero-filled (src-size 0), the (clamped) source is not read
__device__ __forceinline__ void zg_cp16(double2* sdst, const double2* gsrc, const bool ok) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

// SUB = false: C = A B;  SUB = true: C -= A B (the deferred rank-64 update of the blocked Gauss-Jordan)
// n_dev (optional): the number of columns actually in use lives in device memory (the lock-step Newton
// rounds of hpf_lockstep.cuh launch without knowing how many scenarios are still active): column tiles
// beyond it exit at once.
template <bool SUB>
__global__ void __launch_bounds__(256, 2)
zgemm_dmma_kernel(const int M, const int N_, const int K, const double2* __restrict__ A, const size_t lda,
                  const double2* __restrict__ B, const size_t ldb, double2* __restrict__ C, const size_t ldc,
                  const int* __restrict__ n_dev) {
    extern __shared__ __align__(16) double2 zsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;                   // warp grid 4 (M) x 2 (N)
    const int m0 = blockIdx.y * ZG_BM, n0 = blockIdx.x * ZG_BN;
    int N = N_;
    if (n_dev) { const int nd = *n_dev; N = nd < N_ ? nd : N_; }
    if (n0 >= N) return;
    const int nk = (K + ZG_BK - 1) / ZG_BK;

    auto stage_load = [&](const int kt, const int st) {
        double2* sA = zsm + (size_t)st * ZG_STAGE_ELEMS;
        double2* sB = sA + ZG_BM * ZG_LDA;
        const int k0 = kt * ZG_BK;
#pragma unroll
        for (int t = 0; t < 2; ++t) {                          // A tile: 64 rows x 8 complex
            const int e = tid + 256 * t, r = e >> 3, kk = e & 7;
            const bool ok = (m0 + r < M) && (k0 + kk < K) && kt < nk;
            const size_t gr = ok ? (size_t)(m0 + r) : 0, gk = ok ? (size_t)(k0 + kk) : 0;
            zg_cp16(sA + r * ZG_LDA + kk, A + gr * lda + gk, ok);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {                          // B tile: 8 rows x 64 complex
            const int e = tid + 256 * t, kk = e >> 6, c = e & 63;
            const bool ok = (k0 + kk < K) && (n0 + c < N) && kt < nk;
            const size_t gk = ok ? (size_t)(k0 + kk) : 0, gc = ok ? (size_t)(n0 + c) : 0;
            zg_cp16(sB + kk * ZG_LDB + c, B + gk * ldb + gc, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    double cr[2][4][2], ci[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < ZG_STAGES - 1; ++s) stage_load(s, s);

    const int arow = lane >> 2, ak = lane & 3;                 // A fragment: row = lane / 4, k = lane % 4
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group %0;\n" :: "n"(ZG_STAGES - 2) : "memory");
        __syncthreads();
        stage_load(kt + ZG_STAGES - 1, (kt + ZG_STAGES - 1) % ZG_STAGES);     // (refills the stage read in iteration kt - 1)
        const double2* sA = zsm + (size_t)(kt % ZG_STAGES) * ZG_STAGE_ELEMS;
        const double2* sB = sA + ZG_BM * ZG_LDA;
#pragma unroll
        for (int ks = 0; ks < ZG_BK; ks += 4) {
            double2 af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = sA[(wm * 16 + i * 8 + arow) * ZG_LDA + ks + ak];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sB[(ks + ak) * ZG_LDB + wn * 32 + j * 8 + arow];   // B fragment: k = lane % 4, col = lane / 4
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double nai = -af[i].y;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    zg_dmma(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
                    zg_dmma(cr[i][j][0], cr[i][j][1], nai, bf[j].y);
                    zg_dmma(ci[i][j][0], ci[i][j][1], af[i].x, bf[j].y);
                    zg_dmma(ci[i][j][0], ci[i][j][1], af[i].y, bf[j].x);
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1}: two adjacent complex numbers per lane
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = m0 + wm * 16 + i * 8 + (lane >> 2);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
            double2* dst = C + (size_t)r * ldc + c;
            if (SUB) {
                if (c < N) { const double2 o = dst[0]; dst[0] = make_double2(o.x - cr[i][j][0], o.y - ci[i][j][0]); }
                if (c + 1 < N) { const double2 o = dst[1]; dst[1] = make_double2(o.x - cr[i][j][1], o.y - ci[i][j][1]); }
            } else {
                if (c < N) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
                if (c + 1 < N) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
            }
        }
    }
}
