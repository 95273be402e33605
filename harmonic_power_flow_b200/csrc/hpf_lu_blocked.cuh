// Blocked dense LU with partial pivoting + solve for systems that do not fit shared memory
// (N > 192): the linear algebra of HG:229 / HG:476-479 for the large networks (net1 dense
// path, the border system and the fundamental stage of the 200- / 1000-bus configurations).
//
// One CTA factorises one augmented matrix [A | b] that lives in GLOBAL memory (L2 resident or
// streamed), column-major, leading dimension ld (a multiple of 2 so that two consecutive rows
// of a column are one aligned 16-byte word).  Right-looking, panel width 32:
//   1. panel: staged in SHARED memory (panel width 32, 16 or 8, the widest whose remaining rows
//      fit the staging buffer), unblocked elimination with partial pivoting inside it (one
//      thread per row, the pivot search of column j+1 is fused into the update of column j);
//      worked directly on global memory every one of the N column steps would pay ~30
//      dependent L2 round trips;
//   2. the 32 row interchanges and the triangular solve with L11 for every column to the
//      right (one thread per column, 32 values in registers, L11 broadcast from smem);
//   3. trailing update A22 -= L21 U12 on the FP64 TENSOR CORES (mma.sync m8n8k4 DMMA): L21 and
//      U12 slices are staged in shared memory (conflict-free padded strides), every warp owns
//      32 x 32 sub-tiles = 4 x 4 DMMA tiles, accumulators in registers, the trailing matrix is
//      read-modified-written once per panel with 16-byte accesses.  The product is formed
//      TRANSPOSED (D[c][r] = sum_k U12[k][c] L21[r][k]) so that the row-major C fragment of the
//      instruction maps onto two consecutive rows of a column of the column-major matrix.
//      Why tensor cores here: a register-tiled FMA update needs 8 shared-memory operand loads
//      per 16 FMAs (4 x 4 tile), the DMMA path 8 fragment loads per 16 x 256 MACs, and the
//      DMMA pipe has twice the FMA pipe's FP64 rate on this part.
//      Large instance (BIG, matrix streamed from HBM): while the panels have more rows than
//      the CTA has threads they are taken in PAIRS - panel 1 is applied to panel 2's columns
//      only; after panel 2 is factored its interchanges are also applied to panel 1's L21, panel
//      1 is applied to panel 2's 32 pivot rows, and the rest of the trailing matrix takes both
//      panels in ONE rank-64 pass (lub_update_big) - and the U slice of the next column tile
//      is copied asynchronously (LDGSTS) while the current tile is multiplied.
//   4. blocked back substitution (32 x 32 diagonal blocks solved by one warp).
// L is not kept (the right-hand side is carried as column N), so columns left of the panel
// (pair) are never touched again.  Returns 0 or k+1 for a zero / non-finite pivot at step k, like
// lu_solve_smem.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#define LUB_NB 32
#define LUB_TR 128                 // matrix rows per staged slice
#define LUB_TC 128                 // matrix columns per staged slice
#define LUB_SL (LUB_TR + 4)        // smem stride of L21^T  [k][r]   (stride mod 16 == 4)
#define LUB_SU (LUB_NB + 4)        // smem stride of U12^T  [c][k]   (stride mod 16 == 4)
#define LUB_FIXED_DOUBLES (LUB_NB * (LUB_NB + 1) + 2 * LUB_NB + 80)        // L11 + reduction scratch + pivots
#define LUB_STAGE_DOUBLES (LUB_NB * LUB_SL + LUB_TC * LUB_SU)             // L21 / U12 slices (minimum)
#define LUB_SMEM_DOUBLES (LUB_FIXED_DOUBLES + LUB_STAGE_DOUBLES)          // minimum work area
// work area that lets an N-row system run with the widest possible panels, capped at `cap` bytes
__host__ __device__ inline size_t lub_smem_doubles_for(int N, size_t cap_bytes) {
    size_t want = (size_t)LUB_FIXED_DOUBLES + (size_t)(N + 1) * LUB_NB;
    const size_t mn = LUB_SMEM_DOUBLES, mx = cap_bytes / sizeof(double);
    if (want < mn) want = mn;
    if (want > mx) want = mx > mn ? mx : mn;
    return want;
}

__host__ __device__ inline int lub_ld(int N) { return (N + 7) & ~7; }

// 16-byte asynchronous global -> shared copy (LDGSTS); nbytes = 0 writes zeros without reading
__device__ __forceinline__ void lub_cp_async16(void* sdst, const void* gsrc, const int nbytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(sa), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void lub_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void lub_cp_async_wait() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// block-wide argmax of (value, index) with idamax tie-breaking (first maximal |a|)
__device__ __forceinline__ void lub_argmax(double& best, int& bi, double* redv, int* redi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        best = (lane < nw) ? redv[lane] : -1.0;
        bi = (lane < nw) ? redi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { redv[32] = best; redi[32] = bi; }
    }
    __syncthreads();
    best = redv[32];
    bi = redi[32];
}

// Trailing update of the large-matrix instance: C[r][c] -= sum_{k < K} L[r][kbase + k] U[kbase + k][c]
// for rows rb <= r < re and columns cb <= c <= ce, on the FP64 tensor cores.  K = 32 (one panel)
// or 64 (a PAIR of panels: the trailing matrix then streams through HBM once per 64 columns
// instead of once per 32, see lu_solve_blocked_t).  Called by every thread of the CTA after a
// barrier; ends with a barrier.  Kept out of line: one copy of the code for its three call sites
// and its own register allocation.
#define LUB_STAGE_DOUBLES_K64 (64 * LUB_SL + LUB_TC * (64 + 4))
__device__ __noinline__ void lub_update_big(double* __restrict__ A, const int ld, const int kbase, const int K,
                                            const int rb, const int re, const int cb0, const int ce,
                                            double* stage, const int stage_doubles) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = nthr >> 5;
    const int su = K + 4, ksh = (K == 64) ? 6 : 5;               // (stride mod 16 == 4: conflict-free fragments)
    double* Ls = stage;                                          // [K][SL]   Ls[k][r] = -L[r0 + r][kbase + k]
    double* Us = Ls + K * LUB_SL;                                // [TC][su]  Us[c][k] = U[kbase + k][c]
    constexpr int SUBR = LUB_TR / 32, SUBC = LUB_TC / 32;
    // the U slice of the NEXT column tile is copied asynchronously (LDGSTS, no registers) into a
    // second buffer while this tile is multiplied, when the work area holds two of them; the
    // sign of the product therefore lives in the L slice
    const bool dbuf = stage_doubles >= K * LUB_SL + 2 * LUB_TC * su;
    double* Us1 = dbuf ? Us + LUB_TC * su : Us;
    auto stage_u = [&](double* U, const int c0) {
        for (int t = tid; t < LUB_TC * (K >> 1); t += nthr) {
            const int c = t >> (ksh - 1), k = (t & ((K >> 1) - 1)) << 1;
            const bool in = c0 + c <= ce;
            lub_cp_async16(U + c * su + k, in ? A + (size_t)(c0 + c) * ld + kbase + k : A, in ? 16 : 0);
        }
        lub_cp_async_commit();
    };
    const int fr = lane >> 2, fk = lane & 3;                     // fragment row / k index
    for (int r0 = rb; r0 < re; r0 += LUB_TR) {
        for (int t = tid; t < K * LUB_TR; t += nthr) {
            const int k = t / LUB_TR, r = t - k * LUB_TR;
            Ls[k * LUB_SL + r] = (r0 + r < re) ? -A[(size_t)(kbase + k) * ld + r0 + r] : 0.0;
        }
        int buf = 0;
        if (dbuf) stage_u(Us, cb0);
        for (int c0 = cb0; c0 <= ce; c0 += LUB_TC) {
            if (!dbuf) {
                __syncthreads();                                 // previous Us consumers done
                stage_u(Us, c0);
            }
            lub_cp_async_wait();
            __syncthreads();                                     // slices visible; the other buffer is free
            const double* Uc = buf ? Us1 : Us;
            if (dbuf) {
                if (c0 + LUB_TC <= ce) stage_u(buf ? Us : Us1, c0 + LUB_TC);
                buf ^= 1;
            }
            for (int st = warp; st < SUBR * SUBC; st += nw) {
                const int sc = st / SUBR, sr = st - sc * SUBR;
                const int cb = c0 + sc * 32, rbb = r0 + sr * 32;
                if (cb > ce || rbb >= re) continue;
                double acc[4][4][2];
                // C fragment: D[c = cb + 8 ic + fr][r = rbb + 8 ir + 2 fk + {0,1}]
#pragma unroll
                for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                    for (int ir = 0; ir < 4; ++ir) {
                        const int c = cb + 8 * ic + fr, r = rbb + 8 * ir + 2 * fk;
                        if (c <= ce && r + 1 < re) {
                            const double2 v2 = *reinterpret_cast<const double2*>(A + (size_t)c * ld + r);
                            acc[ic][ir][0] = v2.x; acc[ic][ir][1] = v2.y;
                        } else {
                            acc[ic][ir][0] = (c <= ce && r < re) ? A[(size_t)c * ld + r] : 0.0;
                            acc[ic][ir][1] = 0.0;
                        }
                    }
                const double* us = Uc + (sc * 32 + fr) * su + fk;
                const double* ls = Ls + fk * LUB_SL + sr * 32 + fr;
#pragma unroll 2
                for (int kk = 0; kk < K; kk += 4) {
                    double af[4], bf[4];
#pragma unroll
                    for (int ic = 0; ic < 4; ++ic) af[ic] = us[(8 * ic) * su + kk];
#pragma unroll
                    for (int ir = 0; ir < 4; ++ir) bf[ir] = ls[kk * LUB_SL + 8 * ir];
#pragma unroll
                    for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                        for (int ir = 0; ir < 4; ++ir)
                            dmma884(acc[ic][ir][0], acc[ic][ir][1], af[ic], bf[ir]);
                }
#pragma unroll
                for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                    for (int ir = 0; ir < 4; ++ir) {
                        const int c = cb + 8 * ic + fr, r = rbb + 8 * ir + 2 * fk;
                        if (c <= ce && r + 1 < re) {
                            *reinterpret_cast<double2*>(A + (size_t)c * ld + r) =
                                make_double2(acc[ic][ir][0], acc[ic][ir][1]);
                        } else if (c <= ce && r < re) {
                            A[(size_t)c * ld + r] = acc[ic][ir][0];
                        }
                    }
            }
        }
        __syncthreads();                                         // before Ls is overwritten
    }
}

// A: N x (N+1) column-major in global memory, ld even; sm: sm_doubles >= LUB_SMEM_DOUBLES doubles
// of SHARED memory (more = wider panels for large N); sflag: one int of shared memory.
// Two instances: BIG = false for systems whose first (tallest) panel fits the staging buffer - the
// common case, compiled without the in-place global-memory panel path (its 32-element row buffer
// costs the other paths registers: 8 % on the 200-bus configuration); BIG = true otherwise.
template <bool BIG>
__device__ __forceinline__ int lu_solve_blocked_t(double* __restrict__ A, const int N, const int ld,
                                               double* sm, const int sm_doubles, int* sflag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = nthr >> 5;
    double* L11 = sm;                                   // [NB][NB+1]  L11[r][c]
    double* redv = L11 + LUB_NB * (LUB_NB + 1);         // 33
    int* redi = reinterpret_cast<int*>(redv + 34);      // 33 ints
    int* piv = redi + 34;                               // NB ints
    double* prw = redv + 72;                            // [NB] pivot row of the current panel column
    double* stage = sm + LUB_FIXED_DOUBLES;             // panel staging | L21 / U12 slices
    const int stage_doubles = sm_doubles - LUB_FIXED_DOUBLES;
    double* Ls = stage;                                 // [NB][SL]    Ls[k][r] = L21[r][k]
    double* Us = Ls + LUB_NB * LUB_SL;                  // [TC][SU]    Us[c][k] = -U12[k][c]
    double* Pn = stage;                                 // [nb][rows]  panel, column-major
    if (tid == 0) *sflag = 0;
    __syncthreads();

    // permutation of the last register panel as lists (see step 2)
    int* srcTop = reinterpret_cast<int*>(stage + 64);   // [NB] panel-relative source row of top position t
    int* mvSrc = srcTop + LUB_NB;                       // [<= NB] rows that leave the top block ...
    int* mvDst = mvSrc + LUB_NB;                        //         ... and the positions they end at
    int* mvCnt = mvDst + LUB_NB;
    bool deferred = false;                              // (BIG) first panel of a pair factored, its update pending
    for (int k0 = 0; k0 < N;) {
        const int rows = N - k0;
        bool perm_lists = false;
        // the panel is always 32 columns wide (a narrower panel would stream the whole trailing
        // matrix through HBM 2-4 times more often); it is staged in shared memory when it fits
        const bool staged = !BIG || (size_t)rows * LUB_NB <= (size_t)stage_doubles;
        const int nb = min(LUB_NB, N - k0);
        // ---------------- 1. panel ----------------
        if (nb == LUB_NB && rows <= nthr) {
            // REGISTER panel: thread t owns one row of the panel (32 values in registers) for all
            // 32 column steps.  Row interchanges move no data: every thread tracks the panel
            // position `pos` of its row (swap k <-> p: the two owners exchange positions).  Per
            // column step: block argmax (one barrier, every warp reduces the partials redundantly),
            // the pivot row is published through shared memory (second barrier), the active rows
            // update with 31 register FMAs.  (On shared / global memory every step pays ~30
            // dependent round trips per row.)
            double* prow = stage;                                // [32] current pivot row
            double v[LUB_NB];
            int pos = (tid < rows) ? tid : -1;
            if (tid == 0) *mvCnt = 0;                            // (visible after the first barrier below)
            perm_lists = true;
            if (pos >= 0) {
#pragma unroll
                for (int j = 0; j < LUB_NB; ++j) v[j] = A[(size_t)(k0 + j) * ld + k0 + tid];
            } else {
#pragma unroll
                for (int j = 0; j < LUB_NB; ++j) v[j] = 0.0;
            }
            const int aw = (rows + 31) >> 5;                     // warps that own rows
            unsigned* rkey = reinterpret_cast<unsigned*>(redv);  // [aw][2] (hi, lo) of the warp maxima
#pragma unroll
            for (int j = 0; j < LUB_NB; ++j) {
                // argmax of |v[j]| over the rows at positions >= j, idamax tie-breaking (smallest
                // position), with warp REDUX instructions on the bit pattern: for non-negative
                // doubles the (hi, lo) words order like the values, so three 32-bit reductions
                // (max hi, max lo among the hi-maxima, min position among the exact maxima) replace
                // five shuffle rounds on (double, int) pairs
                unsigned mhi = 0u, mlo = 0u;
                int p = 0x7fffffff;
                if (warp < aw) {                                 // (the other warps only keep the barriers)
                    const bool cand = pos >= j;
                    double av = fabs(v[j]);
                    if (av != av) av = CUDART_INF;               // a NaN candidate must surface as a bad pivot
                    const unsigned hi = cand ? (unsigned)__double2hiint(av) : 0u;
                    const unsigned lo = cand ? (unsigned)__double2loint(av) : 0u;
                    mhi = __reduce_max_sync(0xffffffffu, hi);
                    mlo = __reduce_max_sync(0xffffffffu, (hi == mhi) ? lo : 0u);
                    p = (int)__reduce_min_sync(0xffffffffu, (cand && hi == mhi && lo == mlo) ? (unsigned)pos : 0x7fffffffu);
                    if (lane == 0) { rkey[2 * warp] = mhi; rkey[2 * warp + 1] = mlo; redi[warp] = p; }
                }
                __syncthreads();
                if (warp < aw) {
                    // every row-owning warp reduces the partials redundantly (no second barrier)
                    const unsigned hi = (lane < aw) ? rkey[2 * lane] : 0u;
                    const unsigned lo = (lane < aw) ? rkey[2 * lane + 1] : 0u;
                    const unsigned pp = (lane < aw) ? (unsigned)redi[lane] : 0x7fffffffu;
                    mhi = __reduce_max_sync(0xffffffffu, hi);
                    mlo = __reduce_max_sync(0xffffffffu, (hi == mhi) ? lo : 0u);
                    p = (int)__reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? pp : 0x7fffffffu);
                }
                const double best = __hiloint2double((int)mhi, (int)mlo);
                if (!(best > 0.0) || !(best < CUDART_INF)) {
                    if (tid == 0 && *sflag == 0) *sflag = k0 + j + 1;
                }
                if (tid == 0) piv[j] = k0 + p;
                if (pos == p) pos = j;
                else if (pos == j) pos = p;
                if (pos == j) {
#pragma unroll
                    for (int jj = j; jj < LUB_NB; ++jj) prow[jj] = v[jj];
                }
                __syncthreads();
                if (pos > j) {
                    const double l = v[j] * (1.0 / prow[j]);
                    v[j] = l;
#pragma unroll
                    for (int jj = j + 1; jj < LUB_NB; ++jj) v[jj] = fma(-l, prow[jj], v[jj]);
                }
            }
            if (pos >= 0) {
#pragma unroll
                for (int j = 0; j < LUB_NB; ++j) A[(size_t)(k0 + j) * ld + k0 + pos] = v[j];
                // the interchanges as a permutation: rows that end in the top block come from
                // anywhere, rows that end below it always come from the top block
                if (pos < LUB_NB) srcTop[pos] = tid;
                else if (pos != tid) {
                    const int e = atomicAdd(mvCnt, 1);
                    mvSrc[e] = tid;
                    mvDst[e] = pos;
                }
            }
            __syncthreads();
        } else {
            // panel element (row i, panel column j) at PB[j * ps + i - k0]
            double* PB = staged ? Pn : A + (size_t)k0 * ld + k0;
            const size_t ps = staged ? (size_t)rows : (size_t)ld;
            if (staged) {
                for (int t = tid; t < rows * nb; t += nthr) {
                    const int j = t / rows, i = t - j * rows;
                    Pn[t] = A[(size_t)(k0 + j) * ld + k0 + i];
                }
                __syncthreads();
            }
            // pivot of the first column
            double best = -1.0;
            int bi = 0;
            for (int i = tid; i < rows; i += nthr) {
                const double v = fabs(PB[i]);
                if (v > best) { best = v; bi = i; }
            }
            lub_argmax(best, bi, redv, redi);
            for (int j = 0; j < nb; ++j) {
                const int p = bi;                               // panel-relative row of the pivot
                if (!(best > 0.0) || !(best < CUDART_INF)) {
                    if (tid == 0 && *sflag == 0) *sflag = k0 + j + 1;
                }
                if (tid == 0) piv[j] = k0 + p;
                if (p != j && tid < nb) {                       // swap inside the panel
                    double* c0 = PB + (size_t)tid * ps;
                    const double t = c0[j];
                    c0[j] = c0[p];
                    c0[p] = t;
                }
                __syncthreads();
                double* cj0 = PB + (size_t)j * ps;
                const double r = 1.0 / cj0[j];
                // scale column j, rank-1 update of the panel columns to its right; the new
                // column j+1 feeds the next pivot search
                best = -1.0;
                bi = j + 1;
                if (!BIG || staged) {
                    for (int i = j + 1 + tid; i < rows; i += nthr) {
                        const double l = cj0[i] * r;
                        cj0[i] = l;
                        for (int jj = j + 1; jj < nb; ++jj) {
                            double* cj = PB + (size_t)jj * ps;
                            const double v = cj[i] - l * cj[j];
                            cj[i] = v;
                            if (jj == j + 1) {
                                const double av = fabs(v);
                                if (av > best) { best = av; bi = i; }
                            }
                        }
                    }
                } else {
                    // panel in place in global memory (more rows than the stage holds): the pivot row
                    // is first copied to shared memory, then every row issues all its loads before
                    // the first FMA - 31 dependent L2 round trips per row become one
                    __syncthreads();
                    if (tid < nb) prw[tid] = PB[(size_t)tid * ps + j];
                    __syncthreads();
                    for (int i = j + 1 + tid; i < rows; i += nthr) {
                        const double l = cj0[i] * r;
                        cj0[i] = l;
                        double t[LUB_NB];
#pragma unroll
                        for (int jj = 1; jj < LUB_NB; ++jj)
                            if (jj > j && jj < nb) t[jj] = PB[(size_t)jj * ps + i];
#pragma unroll
                        for (int jj = 1; jj < LUB_NB; ++jj)
                            if (jj > j && jj < nb) {
                                t[jj] = fma(-l, prw[jj], t[jj]);
                                PB[(size_t)jj * ps + i] = t[jj];
                                if (jj == j + 1) {
                                    const double av = fabs(t[jj]);
                                    if (av > best) { best = av; bi = i; }
                                }
                            }
                    }
                }
                if (j + 1 < nb) lub_argmax(best, bi, redv, redi);   // (contains the barriers)
                else __syncthreads();
            }
            if (staged) {
                for (int t = tid; t < rows * nb; t += nthr) {
                    const int j = t / rows, i = t - j * rows;
                    A[(size_t)(k0 + j) * ld + k0 + i] = Pn[t];
                }
                __syncthreads();
            }
        }
        const int cr = k0 + nb;                                  // first row / column of the trailing part
        if constexpr (BIG) {
            if (deferred) {
                // SECOND panel of a pair (columns k0 .. k0+31; the first one is k0-32 .. k0-1 and
                // has been applied to this panel's columns only).
                // 2a. this panel's interchanges, on the columns to the right AND on the first
                //     panel's L21 (its rows move with the rows of the trailing matrix)
                for (int t2 = tid; t2 < LUB_NB + (N + 1 - cr); t2 += nthr) {
                    double* col = A + (size_t)(t2 < LUB_NB ? k0 - LUB_NB + t2 : cr + (t2 - LUB_NB)) * ld;
                    for (int t = 0; t < nb; ++t) {
                        const int p = piv[t];
                        if (p != k0 + t) {
                            const double x = col[k0 + t];
                            col[k0 + t] = col[p];
                            col[p] = x;
                        }
                    }
                }
                __syncthreads();
                // 2b. the first panel's contribution to this panel's 32 pivot rows
                lub_update_big(A, ld, k0 - LUB_NB, LUB_NB, k0, cr, cr, N, stage, stage_doubles);
                for (int t = tid; t < nb * nb; t += nthr) {
                    const int c = t / nb, r = t - c * nb;
                    L11[r * (LUB_NB + 1) + c] = A[(size_t)(k0 + c) * ld + k0 + r];
                }
                __syncthreads();
                // 2c. U12 = L11^{-1} A12 for this panel's rows
                for (int c = cr + tid; c <= N; c += nthr) {
                    double* col = A + (size_t)c * ld;
                    double v[LUB_NB];
#pragma unroll
                    for (int t = 0; t < LUB_NB; ++t) v[t] = col[k0 + t];
#pragma unroll
                    for (int t = 1; t < LUB_NB; ++t) {
                        double acc = v[t];
#pragma unroll
                        for (int s2 = 0; s2 < t; ++s2) acc = fma(-L11[t * (LUB_NB + 1) + s2], v[s2], acc);
                        v[t] = acc;
                    }
#pragma unroll
                    for (int t = 1; t < LUB_NB; ++t) col[k0 + t] = v[t];
                }
                __syncthreads();
                // 3. rank-64 update of the trailing matrix with both panels
                lub_update_big(A, ld, k0 - LUB_NB, 2 * LUB_NB, cr, N, cr, N, stage, stage_doubles);
                deferred = false;
                k0 = cr;
                continue;
            }
        }
        // L11 -> smem
        for (int t = tid; t < nb * nb; t += nthr) {
            const int c = t / nb, r = t - c * nb;
            L11[r * (LUB_NB + 1) + c] = A[(size_t)(k0 + c) * ld + k0 + r];
        }
        __syncthreads();
        // ---------------- 2. interchanges + U12 = L11^{-1} A12 (incl. the rhs column N) ----------------
        for (int c = cr + tid; c <= N; c += nthr) {
            double* col = A + (size_t)c * ld;
            double v[LUB_NB];
            if (perm_lists) {
                // independent loads instead of 32 dependent swaps (each an L2 round trip)
#pragma unroll
                for (int t = 0; t < LUB_NB; ++t) v[t] = col[k0 + srcTop[t]];
                const int cnt = *mvCnt;
                for (int e = 0; e < cnt; ++e) col[k0 + mvDst[e]] = col[k0 + mvSrc[e]];
            } else {
                for (int t = 0; t < nb; ++t) {
                    const int p = piv[t];
                    if (p != k0 + t) {
                        const double x = col[k0 + t];
                        col[k0 + t] = col[p];
                        col[p] = x;
                    }
                }
#pragma unroll
                for (int t = 0; t < LUB_NB; ++t) v[t] = (t < nb) ? col[k0 + t] : 0.0;
            }
#pragma unroll
            for (int t = 1; t < LUB_NB; ++t) {
                double acc = v[t];
#pragma unroll
                for (int s2 = 0; s2 < t; ++s2) acc = fma(-L11[t * (LUB_NB + 1) + s2], v[s2], acc);
                v[t] = (t < nb) ? acc : 0.0;
            }
#pragma unroll
            for (int t = 0; t < LUB_NB; ++t)
                if (t < nb && (t > 0 || perm_lists)) col[k0 + t] = v[t];
        }
        __syncthreads();
        if (cr >= N) break;                                      // nothing below the panel
        // ---------------- 3. trailing update on the tensor cores ----------------
        if constexpr (BIG) {
            // Panels are processed in PAIRS while they are factored in place / from the stage (more
            // rows than threads): the first panel is applied to the second panel's 32 columns only,
            // the rest of the trailing matrix then takes both panels in one rank-64 pass - half the
            // HBM traffic of the update, which bounds the large instance.
            const bool pair = nb == LUB_NB && !perm_lists && N - cr > nthr &&
                              stage_doubles >= LUB_STAGE_DOUBLES_K64;
            lub_update_big(A, ld, k0, LUB_NB, cr, N, cr, pair ? cr + LUB_NB - 1 : N, stage, stage_doubles);
            deferred = pair;
        } else {
            for (int r0 = cr; r0 < N; r0 += LUB_TR) {
                // L21 slice: Ls[k][r] = A[r0 + r][k0 + k]
                for (int t = tid; t < LUB_NB * LUB_TR; t += nthr) {
                    const int k = t / LUB_TR, r = t - k * LUB_TR;
                    Ls[k * LUB_SL + r] = (k < nb && r0 + r < N) ? A[(size_t)(k0 + k) * ld + r0 + r] : 0.0;
                }
                for (int c0 = cr; c0 <= N; c0 += LUB_TC) {
                    constexpr int SUBR = LUB_TR / 32, SUBC = LUB_TC / 32;
                    const int fr = lane >> 2, fk = lane & 3;         // fragment row / k index
                    double acc[4][4][2];
                    // C fragment of sub-tile st: D[c = cb + 8 ic + fr][r = rb + 8 ir + 2 fk + {0,1}]
                    auto load_c = [&](const int st) {
                        const int sc = st / SUBR, sr = st - sc * SUBR;
                        const int cb = c0 + sc * 32, rb = r0 + sr * 32;
    #pragma unroll
                        for (int ic = 0; ic < 4; ++ic)
    #pragma unroll
                            for (int ir = 0; ir < 4; ++ir) {
                                const int c = cb + 8 * ic + fr, r = rb + 8 * ir + 2 * fk;
                                if (c <= N && r + 1 < N) {
                                    const double2 v2 = *reinterpret_cast<const double2*>(A + (size_t)c * ld + r);
                                    acc[ic][ir][0] = v2.x; acc[ic][ir][1] = v2.y;
                                } else {
                                    acc[ic][ir][0] = (c <= N && r < N) ? A[(size_t)c * ld + r] : 0.0;
                                    acc[ic][ir][1] = 0.0;
                                }
                            }
                    };
                    __syncthreads();                                 // previous Us consumers done (and Ls visible)
                    for (int t = tid; t < LUB_TC * LUB_NB; t += nthr) {
                        const int c = t / LUB_NB, k = t - c * LUB_NB;
                        Us[c * LUB_SU + k] = (k < nb && c0 + c <= N) ? -A[(size_t)(c0 + c) * ld + k0 + k] : 0.0;
                    }
                    __syncthreads();
                    for (int st = warp; st < SUBR * SUBC; st += nw) {
                        const int sc = st / SUBR, sr = st - sc * SUBR;
                        const int cb = c0 + sc * 32, rb = r0 + sr * 32;
                        if (cb > N || rb >= N) continue;
                        load_c(st);
                        const double* us = Us + (sc * 32 + fr) * LUB_SU + fk;
                        const double* ls = Ls + fk * LUB_SL + sr * 32 + fr;
                        const int nbk = (nb + 3) & ~3;               // (rows k >= nb of the slices are zero)
    #pragma unroll 2
                        for (int kk = 0; kk < nbk; kk += 4) {
                            double af[4], bf[4];
    #pragma unroll
                            for (int ic = 0; ic < 4; ++ic) af[ic] = us[(8 * ic) * LUB_SU + kk];
    #pragma unroll
                            for (int ir = 0; ir < 4; ++ir) bf[ir] = ls[kk * LUB_SL + 8 * ir];
    #pragma unroll
                            for (int ic = 0; ic < 4; ++ic)
    #pragma unroll
                                for (int ir = 0; ir < 4; ++ir)
                                    dmma884(acc[ic][ir][0], acc[ic][ir][1], af[ic], bf[ir]);
                        }
    #pragma unroll
                        for (int ic = 0; ic < 4; ++ic)
    #pragma unroll
                            for (int ir = 0; ir < 4; ++ir) {
                                const int c = cb + 8 * ic + fr, r = rb + 8 * ir + 2 * fk;
                                if (c <= N && r + 1 < N) {
                                    *reinterpret_cast<double2*>(A + (size_t)c * ld + r) =
                                        make_double2(acc[ic][ir][0], acc[ic][ir][1]);
                                } else if (c <= N && r < N) {
                                    A[(size_t)c * ld + r] = acc[ic][ir][0];
                                }
                            }
                    }
                }
                __syncthreads();                                     // before Ls is overwritten
            }
        }
        k0 = cr;
    }
    __syncthreads();
    // ---------------- 4. blocked back substitution U x = y (y = column N) ----------------
    double* y = A + (size_t)N * ld;
    double* xs = Ls;                                             // NB solved values
    for (int kb = ((N - 1) / LUB_NB) * LUB_NB; kb >= 0; kb -= LUB_NB) {
        const int nb = min(LUB_NB, N - kb);
        for (int t = tid; t < nb * nb; t += nthr) {
            const int c = t / nb, r = t - c * nb;
            L11[r * (LUB_NB + 1) + c] = A[(size_t)(kb + c) * ld + kb + r];   // diagonal block of U
        }
        __syncthreads();
        if (warp == 0) {
            double v = (lane < nb) ? y[kb + lane] : 0.0;
            for (int j = nb - 1; j >= 0; --j) {
                const double d = L11[j * (LUB_NB + 1) + j];
                const double xj = __shfl_sync(0xffffffffu, v, j) / d;
                if (lane == j) v = xj;
                else if (lane < j) v = fma(-L11[lane * (LUB_NB + 1) + j], xj, v);
            }
            if (lane < nb) { y[kb + lane] = v; xs[lane] = v; }
        }
        __syncthreads();
        for (int i = tid; i < kb; i += nthr) {
            double acc = y[i];
            for (int j = 0; j < nb; ++j) acc = fma(-A[(size_t)(kb + j) * ld + i], xs[j], acc);
            y[i] = acc;
        }
        __syncthreads();
    }
    return *sflag;
}

// Which instance a system of order N needs with a work area of sm_doubles (host side: the kernels
// are instantiated per instance - linking both into one kernel costs the small one 10 %).
__host__ __device__ inline bool lub_needs_big(int N, size_t sm_doubles) {
    return (size_t)N * LUB_NB > sm_doubles - LUB_FIXED_DOUBLES;
}
