// Kernel 4 of the north star, shared-memory resident: dense LU with partial pivoting + triangular
// solves of the augmented system [A | b] (update_harmonic_state_vec = spsolve(J, f), HG:476-479),
// PANEL-BLOCKED.  Same storage as lu_solve_smem (hpf_device.cuh): column-major, ODD leading
// dimension, column N = right-hand side, solution returned in it.
//
// lu_solve_smem applies one rank-1 update of the whole trailing matrix per pivot: one LDS + one STS
// per DFMA, N passes over the matrix (ncu, round 1: shared-memory wavefronts 55 %, FP64 pipe 11 %).
// Here the columns are taken in panels of NB = 8:
//   P1  panel: ONE warp holds the panel (rows k0.., 8 columns) in registers and runs its 8 pivot
//       steps without any block barrier and without touching shared memory: warp-wide argmax with
//       three REDUX on the bit patterns of |a| (ties -> lowest row, NaNs skipped, exactly the
//       classic search), pivot row broadcast with shuffles, rows are NOT moved - every register row
//       carries the position it would have after the classic interchanges (two integer updates per
//       step instead of 32 register moves) and is scattered to that position at write-out.  The
//       multipliers L21 go to a side buffer Lp [8][N] (the left factor is not needed afterwards:
//       b is carried along as column N).  (All 8 warps running the pivot steps redundantly on
//       shared memory - the first version of this file - spent 68 % of the kernel's instructions
//       there: ncu, profiles/r2_ncu_lu_panel_v1.txt.)  The instance is chosen by the number of
//       32-row chunks left (1..4).  N > 128: lu_solve_smem;
//   P2  one THREAD per trailing column: the panel's 8 interchanges applied at once (net permutation
//       of <= 16 positions, traced backwards through the swaps), then the 8 x 8 unit-lower solve
//       that turns the column's pivot rows into U12;
//   P3  rank-8 update of the trailing matrix: lane = row (up to three rows per thread, multipliers
//       in registers), warp = two columns at a time, U12 read by broadcast; branch-free (clamped
//       loads, predicated stores) so that the 6 FMA chains of a thread interleave.
//   LOOK-AHEAD: a pivot step is a long dependent chain of ONE warp (clock64 phase timing,
//       profiles/r2_lu_phase_timing.txt: P1 was 53 % of the factorisation, P3 24 %, P2 14 %), so
//       the NEXT panel's 8 columns are brought up to date first (P2 + P3 on them, one column per
//       warp), then warp 0 factors them while warps 1..7 do P2 + P3 on the rest of the trailing
//       matrix (named barrier among those 7 warps between P2 and P3); two Lp / table buffers
//       alternate; two full barriers per panel.
// Back substitution in blocks of 8 as well: the 8 x 8 triangle in registers (every warp,
// redundantly), then one thread per remaining row.
// ROUNDING: every matrix element sees exactly the FMAs of the unblocked elimination, in the same
// order (k ascending; back substitution j descending), with the same multipliers and pivots - the
// result is BIT-IDENTICAL to lu_solve_smem (tests/test_gpu_parity.py compares them).
#pragma once
#include "hpf_device.cuh"

#define HPF_LUP_NB 8
#define HPF_LUP_RP 3      // row chunks (of 32) per pass of the rank-8 update
// phase-timing hook of profiles/tools/lu_phase_timing.cu (compiled out of the product)
#ifndef LUP_TICK
#define LUP_TICK(i)
#endif

// extra shared memory (doubles) next to the matrix: 2 x (Lp [NB][N] + permutation table, 16 x int2)
__host__ __device__ inline size_t lup_extra_doubles(int N) { return 2 * ((size_t)HPF_LUP_NB * N + 2 * HPF_LUP_NB) + 2; }

// P1 in registers (N - k0 <= 32 RMAX).  Returns the interchanges p_c (classic positions) in pc[].
template <int RMAX>
__device__ __forceinline__ void lup_panel_regs(double* A, const int N, const int ld, const int k0, const int nb,
                                               double* rinv, int* sflag, double* Lp, int (&pc)[HPF_LUP_NB]) {
    constexpr int NB = HPF_LUP_NB;
    const int lane = threadIdx.x & 31;
    const int k1 = k0 + nb;
    double a[RMAX][NB];
    int vpos[RMAX];                     // classic position of the register row (-1: no row)
#pragma unroll
    for (int ch = 0; ch < RMAX; ++ch) {
        const int i = k0 + lane + 32 * ch;
        vpos[ch] = (i < N) ? i : -1;
#pragma unroll
        for (int c = 0; c < NB; ++c) a[ch][c] = (i < N && c < nb) ? A[(size_t)(k0 + c) * ld + i] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        pc[c] = -1;
        if (c < nb) {                                                  // (warp-uniform)
            const int kc = k0 + c;
            // ---- pivot search: first maximal |a| among the rows at positions >= kc, NaNs skipped.
            // Branch-free: every test below is a select (divergent branches cost BSSY/BSYNC pairs:
            // 28 % of this warp's stall samples in the first version) ----
            double bv = -1.0;
            int bpos = 0x7fffffff;
#pragma unroll
            for (int ch = 0; ch < RMAX; ++ch) {
                const double v = fabs(a[ch][c]);
                const bool better = (vpos[ch] >= kc) && (v > bv || (v == bv && vpos[ch] < bpos));
                bv = better ? v : bv;
                bpos = better ? vpos[ch] : bpos;
            }
            // warp argmax on the bit patterns (monotone for v >= 0; 0 = no candidate), ties -> lowest position
            const unsigned long long bkey = (bv >= 0.0) ? (unsigned long long)__double_as_longlong(bv) + 1ull : 0ull;
            const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(bkey >> 32));
            const unsigned lo = __reduce_max_sync(0xffffffffu, ((unsigned)(bkey >> 32) == hi) ? (unsigned)bkey : 0u);
            const unsigned long long mkey = ((unsigned long long)hi << 32) | lo;
            const int praw = (int)__reduce_min_sync(0xffffffffu, (unsigned)((bkey == mkey) ? bpos : 0x7fffffff));
            const int p = mkey ? praw : kc;
            // (nothing comparable in the column - all NaN: the rank-1 LU keeps row kc; flagged below)
            pc[c] = p;
            const double best = mkey ? __longlong_as_double((long long)(mkey - 1ull)) : -1.0;
            if (!(best > 0.0) || !(best < CUDART_INF)) {                // (warp-uniform)
                if (lane == 0 && *sflag == 0) *sflag = kc + 1;
            }
            // ---- pivot row (columns >= c) from its owner: the chunk that holds it is warp-uniform.
            // (Tried: every lane prepares its own candidate row and reciprocal while the reductions are
            // in flight, the winner's are shuffled - the shorter dependency chain lost to the extra
            // instructions: 880 against 770 cycles per pivot step, clock64) ----
            unsigned own[RMAX];
#pragma unroll
            for (int ch = 0; ch < RMAX; ++ch) own[ch] = __ballot_sync(0xffffffffu, vpos[ch] == p);
            double prow[NB];
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) prow[cc] = 0.0;
            bool found = false;
#pragma unroll
            for (int ch = 0; ch < RMAX; ++ch) {
                if (!found && own[ch]) {                                  // (warp-uniform)
                    found = true;
                    const int ol = __ffs(own[ch]) - 1;
#pragma unroll
                    for (int cc = c; cc < NB; ++cc) prow[cc] = __shfl_sync(0xffffffffu, a[ch][cc], ol);
                }
            }
            const double r = 1.0 / prow[c];
            if (lane == 0) rinv[kc] = r;
            // ---- the interchange (k <-> p) moves no data: the two rows trade positions ----
#pragma unroll
            for (int ch = 0; ch < RMAX; ++ch) {
                const int o = vpos[ch];
                vpos[ch] = (o == kc) ? p : ((o == p) ? kc : o);
            }
            // ---- multipliers and the panel's remaining columns (rows at positions > kc) ----
#pragma unroll
            for (int ch = 0; ch < RMAX; ++ch) {
                const bool act = vpos[ch] > kc;
                const double l = a[ch][c] * r;
                a[ch][c] = act ? l : a[ch][c];
                // (inactive rows: multiplier 0 - a - 0 * u is a for every finite u; a non-finite
                // pivot row has been flagged above)
                const double le = act ? l : 0.0;
#pragma unroll
                for (int cc = c + 1; cc < NB; ++cc) a[ch][cc] -= le * prow[cc];
            }
        }
    }
    // ---- write-out at the classic positions: multipliers -> Lp, the pivot block -> A (U11 for the
    // back substitution; its strictly lower part is the unit-lower L11 that P2 reads from Lp) ----
#pragma unroll
    for (int ch = 0; ch < RMAX; ++ch) {
        const int v = vpos[ch];
        if (v >= 0) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nb) Lp[(size_t)c * N + v] = a[ch][c];
            if (v < k1) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < nb) A[(size_t)(k0 + c) * ld + v] = a[ch][c];
            }
        }
    }
}

// P2 of one trailing column (one thread): the panel's interchanges at once, then the pivot rows -> U12.
__device__ __forceinline__ void lup_p2_column(double* colj, const int2* tab, const double* Lp, const int N,
                                              const int k0, const int nb) {
    constexpr int NB = HPF_LUP_NB;
    double v[2 * NB];
    int ps[NB];
    // (branch-free: unused table entries read row k0 and are not written back)
#pragma unroll
    for (int t = 0; t < 2 * NB; ++t) {
        const int2 e = tab[t];
        v[t] = colj[e.x >= 0 ? e.y : k0];
        if (t >= NB) ps[t - NB] = e.x;
    }
#pragma unroll
    for (int t = 0; t < NB; ++t)
        if (ps[t] >= 0) colj[ps[t]] = v[NB + t];                 // rows pushed out of the pivot block
    // pivot rows: u_r = a_r - sum_{c < r} l_rc u_c (c ascending: the order of the unblocked updates)
    if (nb == NB) {
#pragma unroll
        for (int rr = 0; rr < NB; ++rr) {
#pragma unroll
            for (int c = 0; c < rr; ++c) v[rr] -= Lp[(size_t)c * N + k0 + rr] * v[c];
            colj[k0 + rr] = v[rr];
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < NB; ++rr) {
            if (rr < nb) {
#pragma unroll
                for (int c = 0; c < rr; ++c) v[rr] -= Lp[(size_t)c * N + k0 + rr] * v[c];
                colj[k0 + rr] = v[rr];
            }
        }
    }
}

// P3 of one warp: rows r0 + lane + 32 q (q < RP), columns jfirst, jfirst + jstep, ... <= jlast, two at
// a time.  Branch-free: rows / columns beyond the matrix are clamped for the loads, get multiplier 0
// and are not stored (divergent `if (row < N)` bodies made ptxas serialise the FMA chains of a
// thread, one exposed 8-deep chain at a time).
template <int RP>
__device__ __forceinline__ void lup_p3_warp(double* A, const int ld, const double* Lp, const int N, const int k0,
                                            const int nb, const int r0, const int jfirst, const int jstep,
                                            const int jlast) {
    constexpr int NB = HPF_LUP_NB;
    const int lane = threadIdx.x & 31;
    int ir[RP];
    double l[RP][NB];
#pragma unroll
    for (int q = 0; q < RP; ++q) {
        ir[q] = r0 + lane + 32 * q;
#pragma unroll
        for (int c = 0; c < NB; ++c) l[q][c] = (c < nb && ir[q] < N) ? Lp[(size_t)c * N + ir[q]] : 0.0;
    }
    for (int j = jfirst; j <= jlast; j += 2 * jstep) {
        const bool two = j + jstep <= jlast;
        double* col0 = A + (size_t)j * ld;
        double* col1 = A + (size_t)(two ? j + jstep : j) * ld;
        double u0[NB], u1[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            u0[c] = col0[k0 + (c < nb ? c : 0)];
            u1[c] = col1[k0 + (c < nb ? c : 0)];
        }
        double x0[RP], x1[RP];
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int ic = ir[q] < N ? ir[q] : N - 1;
            x0[q] = col0[ic];
            x1[q] = col1[ic];
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) {
#pragma unroll
            for (int q = 0; q < RP; ++q) {
                x0[q] -= l[q][c] * u0[c];
                x1[q] -= l[q][c] * u1[c];
            }
        }
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            if (ir[q] < N) col0[ir[q]] = x0[q];
            if (two && ir[q] < N) col1[ir[q]] = x1[q];
        }
    }
}

// One column of the NEXT panel, by one warp (it is on the critical path: the panel cannot be factored
// before it): interchanges (lanes 0..15 move one value each), the 8 x 8 unit-lower solve column by
// column with shuffles (lane r = pivot row r), then the rank-nb update of the column's rows with U12
// taken from the lanes instead of shared memory.
__device__ __forceinline__ void lup_next_column(double* __restrict__ colj, const int2* __restrict__ tab,
                                                const double* __restrict__ Lp, const int N, const int k0, const int nb) {
    constexpr int NB = HPF_LUP_NB;
    const int lane = threadIdx.x & 31;
    const int k1 = k0 + nb;
    const int2 e = tab[lane & (2 * NB - 1)];
    double v = colj[(lane < 2 * NB && e.x >= 0) ? e.y : k0];
    double lrow[NB];                                                  // lane r < nb: L11[r][c], c < r
#pragma unroll
    for (int c = 0; c < NB - 1; ++c) lrow[c] = (lane < nb && c < lane) ? Lp[(size_t)c * N + k0 + lane] : 0.0;
    // multipliers of this lane's rows (all passes of up to 4 chunks: N <= 128)
    double l[4][NB];
    double x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = k1 + lane + 32 * q;
        x[q] = colj[i < N ? i : N - 1];
#pragma unroll
        for (int c = 0; c < NB; ++c) l[q][c] = (c < nb && i < N) ? Lp[(size_t)c * N + i] : 0.0;
    }
    __syncwarp();
    if (lane >= NB && lane < 2 * NB && e.x >= 0) colj[e.x] = v;       // rows pushed out of the pivot block
    // (a pushed-out row that lies below the block is one of this lane's x[] rows: refresh after the store)
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = k1 + lane + 32 * q;
        x[q] = colj[i < N ? i : N - 1];
    }
    // unit-lower solve: u_c is final once columns < c have been applied (c ascending, like the rank-1 updates)
    double u[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        u[c] = __shfl_sync(0xffffffffu, v, c);
        if (c < NB - 1) v -= lrow[c] * u[c];                           // (lrow[c] = 0 for lanes <= c and lanes >= nb)
    }
    if (lane < nb) colj[k0 + lane] = v;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] -= l[q][c] * u[c];           // (l = 0 beyond nb / beyond the matrix)
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = k1 + lane + 32 * q;
        if (i < N) colj[i] = x[q];
    }
}

// all row passes of P3 for one warp's columns
__device__ __forceinline__ void lup_p3_rows(double* A, const int ld, const double* Lp, const int N, const int k0,
                                            const int nb, const int jfirst, const int jstep, const int jlast) {
    for (int r0 = k0 + nb; r0 < N; r0 += 96) {
        const int left = N - r0;                                     // (warp-uniform)
        if (left > 64) lup_p3_warp<3>(A, ld, Lp, N, k0, nb, r0, jfirst, jstep, jlast);
        else if (left > 32) lup_p3_warp<2>(A, ld, Lp, N, k0, nb, r0, jfirst, jstep, jlast);
        else lup_p3_warp<1>(A, ld, Lp, N, k0, nb, r0, jfirst, jstep, jlast);
    }
}

// P1 of the panel at k0 by the calling warp + its interchange table (lanes 0..15)
__device__ __forceinline__ void lup_panel_and_table(double* A, const int N, const int ld, const int k0, double* rinv,
                                                    int* sflag, double* Lp, int2* tab) {
    constexpr int NB = HPF_LUP_NB;
    const int lane = threadIdx.x & 31;
    const int nb = (N - k0 < NB) ? N - k0 : NB;
    int pc[NB];
    const int left = N - k0;                                         // (warp-uniform)
    if (left > 96) lup_panel_regs<4>(A, N, ld, k0, nb, rinv, sflag, Lp, pc);
    else if (left > 64) lup_panel_regs<3>(A, N, ld, k0, nb, rinv, sflag, Lp, pc);
    else if (left > 32) lup_panel_regs<2>(A, N, ld, k0, nb, rinv, sflag, Lp, pc);
    else lup_panel_regs<1>(A, N, ld, k0, nb, rinv, sflag, Lp, pc);
    if (lane < 2 * NB) {
        // net effect of the panel's interchanges on the <= 16 positions they touch: the value that
        // ends up at position x comes from s_0(s_1(...s_{nb-1}(x))), s_c = (k0 + c  p_c)
        int pos = (lane < nb) ? k0 + lane : -1;
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (lane == NB + c) pos = pc[c];
        int x = pos;
#pragma unroll
        for (int c = NB - 1; c >= 0; --c) {
            if (c < nb) {
                if (x == k0 + c) x = pc[c];
                else if (x == pc[c]) x = k0 + c;
            }
        }
        tab[lane] = make_int2(pos, x);
    }
}

__device__ __forceinline__ int lu_solve_smem_panel(double* A, const int N, const int ld, double* rinv,
                                                   int* sflag, double* X) {
    constexpr int NB = HPF_LUP_NB;
    if (N > 128 || blockDim.x < 64) return lu_solve_smem(A, N, ld, rinv, sflag);
    const int tid = threadIdx.x, warp = tid >> 5, nw = blockDim.x >> 5;
    const size_t half = (size_t)NB * N + 2 * NB;
    if (tid == 0) *sflag = 0;
    __syncthreads();
    LUP_TICK(0);
    // prologue: the first panel
    if (warp == 0) lup_panel_and_table(A, N, ld, 0, rinv, sflag, X + 2 * NB, reinterpret_cast<int2*>(X));
    LUP_TICK(1);
    __syncthreads();
    int cur = 0;
    for (int k0 = 0; k0 < N; k0 += NB) {
        const int nb = (N - k0 < NB) ? N - k0 : NB, k1 = k0 + nb;
        const bool next = k1 < N;
        const int nbn = next ? ((N - k1 < NB) ? N - k1 : NB) : 0, k2 = k1 + nbn;
        const int2* tab = reinterpret_cast<const int2*>(X + cur * half);
        const double* Lp = X + cur * half + 2 * NB;
        LUP_TICK(2);
        // the next panel's columns first, one per warp: interchanges + U12 (lane 0), rank-nb update
        if (next) {
            for (int j = k1 + warp; j < k2; j += nw) lup_next_column(A + (size_t)j * ld, tab, Lp, N, k0, nb);
        }
        __syncthreads();
        LUP_TICK(3);
        if (warp == 0) {
            // warp 0 factors the next panel while the other warps update the rest of the trailing matrix
            if (next)
                lup_panel_and_table(A, N, ld, k1, rinv, sflag, X + (cur ^ 1) * half + 2 * NB,
                                    reinterpret_cast<int2*>(X + (cur ^ 1) * half));
            LUP_TICK(4);
        } else {
            const int others = (nw - 1) * 32;
            for (int j = k2 + (tid - 32); j <= N; j += others) lup_p2_column(A + (size_t)j * ld, tab, Lp, N, k0, nb);
            asm volatile("bar.sync 1, %0;" ::"r"(others) : "memory");
            lup_p3_rows(A, ld, Lp, N, k0, nb, k2 + (warp - 1), nw - 1, N);
            LUP_TICK(4);
        }
        __syncthreads();
        LUP_TICK(5);
        cur ^= 1;
    }
    LUP_TICK(7);
    // ================= back substitution U x = b, blocks of NB from the bottom =================
    {
        double* b = A + (size_t)N * ld;
        const int last = ((N - 1) / NB) * NB;
        for (int jb = last; jb >= 0; jb -= NB) {
            const int nbb = (N - jb < NB) ? N - jb : NB;
            // the triangle of the block, redundantly in every warp that owns rows above it (broadcast loads)
            double x[NB];
            if (warp * 32 < jb || warp == 0) {
#pragma unroll
                for (int c = 0; c < NB; ++c) x[c] = (c < nbb) ? b[jb + c] : 0.0;
#pragma unroll
                for (int c = NB - 1; c >= 0; --c) {
                    if (c < nbb) {
                        x[c] *= rinv[jb + c];
                        const double* colj = A + (size_t)(jb + c) * ld + jb;
#pragma unroll
                        for (int i = 0; i < c; ++i) x[i] -= colj[i] * x[c];
                    }
                }
                // rows above the block: one thread per row, j descending like the column sweep
                for (int i = tid; i < jb; i += blockDim.x) {
                    double bi = b[i];
#pragma unroll
                    for (int c = NB - 1; c >= 0; --c)
                        if (c < nbb) bi -= A[(size_t)(jb + c) * ld + i] * x[c];
                    b[i] = bi;
                }
            }
            __syncthreads();
            if (tid < nbb) {
                double xv = x[0];
#pragma unroll
                for (int c = 1; c < NB; ++c)
                    if (tid == c) xv = x[c];
                b[jb + tid] = xv;
            }
        }
    }
    __syncthreads();
    LUP_TICK(8);
    return *sflag;
}
