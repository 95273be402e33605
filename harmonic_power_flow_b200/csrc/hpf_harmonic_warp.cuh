// Harmonic Newton-Raphson (HG:511-560, structured step - see hpf_structured.cuh), the kernel the
// BASELINE 4-bus shapes run on: ONE WARP PER HARMONIC, lane = scenario.
//
// harm_tile_kernel spreads the stacked rows of a scenario over 7 warps, so every operand of
// every complex MAC is a shared-memory load (ncu, round 1: shared-memory wavefronts 62 % of
// peak, FP64 pipe 32 %, 16 warps per SM).  Here thread (lane, h) OWNS the n phasors of harmonic
// h of its scenario:
//   * state V_m, V_a [n] and the phasors V [n] live in REGISTERS for the whole Newton loop;
//   * Y(h), Y_N, G and the reduced fundamental operator are kernel parameters
//     (__grid_constant__, constant bank): the harmonic index is made warp-uniform with a
//     warp reduction (CREDUX writes a uniform register), so every constant is a
//     `c[0x0][UR + imm]` operand of the DFMA itself - no load instruction, no shared memory;
//   * the rows (Y_h V_h)_i of a harmonic need nothing but the thread's own registers; the only
//     exchange between the warps of a tile is the Norton contraction (HG:313-323), which reads
//     the nonlinear buses' phasors of all harmonics (q H complex per scenario, shared memory),
//     V_F + u_F of the linear buses (m complex) and the mismatch norm (one shared atomic max
//     on the sign-cleared bit patterns per thread).
// The fundamental warp (h = 0) also owns the power rows (HG:372-380) and the border system for
// the polar unknowns x_F of the linear buses.  That system is assembled in its reduced complex
// form: with u_b = u0_k - sum_j G[k][j] u_j for the fundamental nonlinear buses,
//     dS_i = u_i conj(I_i) + V_i conj(sum_{j<m} Yeff_ij u_j) + V_i conj(sum_k Y_{i,m+k} u0_k),
//     Yeff_ij = Y1_ij - sum_k Y1_{i,m+k} G[k][j]        (constant, computed once per network),
// so that with t_ij = V_i conj(Yeff_ij V_j) and s_i = V_i conj(I_i)
//     d/dtheta_j : j ([i=j] s_i - t_ij)        d/dV_m,j : ([i=j] s_i + t_ij) / V_m,j
//     rhs_i = -(P_i + jQ_i) - V_i conj(sum_{j<m} Yeff_ij V_j - sum_k Y1_{i,m+k} w_N,k)
// - the same Newton step dx = J^{-1} f as hpf_structured.cuh describes, a quarter of the
// instructions of the entry-by-entry assembly (no per-entry polar conversion of u_b).
//
// One round = one Newton iteration of every lane:
//   A  refilled lanes load their scenario; sin/cos of the thread's n angles; publish the
//      nonlinear buses' phasors                                                    | barrier
//   B  mismatch rows of harmonic h (HG:326-357) incl. the Norton contraction; warp 0: power
//      rows + border system -> V_F + u_F; ||f||_inf by shared atomic max           | barrier
//   C  decisions (every warp computes the same ones); finished lanes write their results;
//      u_z = -V_z - G (V_F + u_F) - w_N, polar conversion, state update in registers;
//      finished lanes are refilled from the global queue                           | barrier (only if a lane finished)
// Lanes are refilled individually because iteration counts differ (8..40).
#pragma once
#include <type_traits>
#include "hpf_structured.cuh"
// (needs absmax_bits of hpf_lane.cuh: include after it)

template <class D, bool COUPLED>
struct HwConsts {
    static constexpr int nZ = D::n * D::H - D::m;
    double2 Y[D::H * D::n * D::n];                               // Y(h) [H][n][n]
    double2 YNk[COUPLED ? D::q * D::H * D::H : D::q * D::H];     // Y_N per nonlinear bus [q][H][H] / [q][H]
    double2 G[nZ * D::m];                                        // G = A_ZZ^-1 A_ZF  [nZ][m]
    double2 Yeff[D::m * D::m];                                   // reduced fundamental operator (rows 0..m-1)
};

template <class D>
__host__ __device__ constexpr size_t harm_hw_smem_bytes() {
    // sVnl [q][H][32] c128 | sW [nZ][32] c128 | sTot [m][32] c128 | sP, sQ [m][32] | red [2][32] | sbase
    return ((size_t)(2 * D::q * D::H + 2 * (D::n * D::H - D::m) + 2 * D::m + 2 * D::m + 2) * 32) * sizeof(double) + 64 +
           sizeof(double2) * (D::H * D::n * D::n + D::q * D::H * D::H + (D::n * D::H - D::m) * D::m);
}

// 1 / x without the library's special-case path (a CALL, which also clobbers the uniform
// registers that hold the constant-bank addresses): MUFU.RCP64H seed (2^-23) and two Newton steps,
// <= 1 ulp for normal x; 0, Inf and denormals give NaN (such a state is reported as non-finite anyway).
__device__ __forceinline__ double rcp_fast(const double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// numpy's a % (2 pi) (HG:548) without fmod: for |a| < 1e6 the quotient k = trunc(a / 2pi) is exact
// up to one unit, a - k * fl(2pi) is exactly representable (it is fmod's result or off by one
// period), so ONE fma reproduces fmod bit for bit after the off-by-one correction.
__device__ __forceinline__ double mod_twopi_fast(const double va) {
    const double twopi = 2.0 * CUDART_PI;
    if (!(fabs(va) < 1.0e6)) return mod_twopi(va);        // huge, Inf, NaN: library path
    const double k = trunc(va * 0.15915494309189535);
    double r = fma(-k, twopi, va);
    if (va >= 0.0) {                                       // fmod: sign of the dividend, |r| < 2 pi
        if (r < 0.0) r += twopi;
        else if (r >= twopi) r -= twopi;
    } else {
        if (r > 0.0) r -= twopi;
        else if (r <= -twopi) r += twopi;
    }
    if (r != 0.0) { if (r < 0.0) r += twopi; } else r = 0.0;   // python %: the divisor's sign
    return r;
}

// Gaussian elimination with partial pivoting on a register-resident augmented system (every index
// is a compile-time constant after unrolling: the pivot row is bubbled up with selects).
template <int NX>
__device__ __forceinline__ int gauss_regs(double (&A)[NX][NX + 1], double (&x)[NX]) {
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NX; ++k) {
#pragma unroll
        for (int i = k + 1; i < NX; ++i) {
            const bool sw = fabs(A[i][k]) > fabs(A[k][k]);
#pragma unroll
            for (int cc = k; cc <= NX; ++cc) {
                const double t0 = A[k][cc], t1 = A[i][cc];
                A[k][cc] = sw ? t1 : t0;
                A[i][cc] = sw ? t0 : t1;
            }
        }
        const double pv = fabs(A[k][k]);
        if (!(pv > 0.0) || !(pv < CUDART_INF)) bad = 1;
        const double r = rcp_fast(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < NX; ++i) {
            const double l = A[i][k] * r;
#pragma unroll
            for (int cc = k + 1; cc <= NX; ++cc) A[i][cc] -= l * A[k][cc];
        }
        A[k][k] = r;                                         // keep the reciprocal for the back substitution
    }
#pragma unroll
    for (int k = NX - 1; k >= 0; --k) {
        double sv = A[k][NX];
#pragma unroll
        for (int cc = k + 1; cc < NX; ++cc) sv -= A[k][cc] * x[cc];
        x[k] = sv * A[k][k];
    }
    return bad;
}

__device__ __forceinline__ double2 hw_lds128(const double2* sptr) {
    double2 v;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sptr);
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(sa));
    return v;
}

__device__ __forceinline__ void hw_cp_async16(void* sdst, const void* gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sa), "l"(gsrc) : "memory");
}

// The Newton loop of one warp.  FUND = true: the fundamental warp (h = 0 is a compile-time constant:
// its constants are immediate constant-bank operands); FUND = false: a harmonic warp, h >= 1 is
// re-derived into a uniform register in every phase.  Both instances execute the same sequence
// of CTA barriers (the third one depends on donemask, which every warp computes identically).
template <class D, bool COUPLED, bool FUND>
__device__ __forceinline__ void harm_hw_loop(const HwConsts<D, COUPLED>& C, const HarmTileArgs& a, double* smem,
                                             const int hv) {
    constexpr int n = D::n, m = D::m, c = D::c, H = D::H, q = D::q, nH = n * H, nZ = nH - m;
    constexpr int nth = m - 1, nv = m - c, nx = nth + nv;
    constexpr int T = 32;
    double2* sVnl = reinterpret_cast<double2*>(smem);             // [q][H][T] phasors of the nonlinear buses
    double2* sW = sVnl + q * H * T;                               // [nZ][T]   w_N (row z = s - m)
    double2* sTot = sW + nZ * T;                                  // [m][T]    V_F + u_F
    double* sP = reinterpret_cast<double*>(sTot + m * T);         // [m][T]
    double* sQ = sP + m * T;                                      // [m][T]
    unsigned long long* red = reinterpret_cast<unsigned long long*>(sQ + m * T);   // [2][T] ||f||_inf bit patterns
    int* sbase = reinterpret_cast<int*>(red + 2 * T);
    // the harmonic warps read their constants from a shared-memory copy (one broadcast wavefront per
    // LDS.128): with a run-time harmonic index ptxas turns constant-bank reads into register-indexed
    // LDC, whose throughput (not the FP64 pipe) then bounds the kernel - measured
    const double2* cY = reinterpret_cast<const double2*>(sbase + 16);
    const double2* cYN = cY + H * n * n;
    const double2* cG = cYN + (COUPLED ? q * H * H : q * H);

    const int lane = threadIdx.x & 31;
    // The harmonic index is warp-uniform.  It is re-derived with a warp reduction (CREDUX writes a
    // UNIFORM register) at the start of every phase so that the constants are c[0x0][UR + imm]
    // operands of the DFMAs themselves: a uniform register does not survive a CALL, and the
    // phases below contain none (rcp_fast / mod_twopi_fast instead of the library routines).
    const size_t B = (size_t)a.B;
    const unsigned lt = (1u << lane) - 1u;
    // loop-invariant row bases of this thread (rows s = h n + i, i = 0..n-1; consecutive rows are B
    // elements apart): only the scenario offset is added at a refill / write-out
    const size_t row0 = (size_t)hv * n * B;
    const double2* const wN0 = a.wN + row0 - (size_t)m * B;       // row z = h n + i - m (never read for z < 0)
    const double2* const IN0 = a.I_N + (size_t)hv * B;             // row k H + h

    int sc = sbase[0] + lane;
    if ((size_t)sc >= B) sc = -1;
    bool isnew = true;
    int itc = 0, stat = 0, cur = 0, rc = 0;
    double Vm[n], Va[n];
    double2 V[n], IN[q], inj[q];
#pragma unroll
    for (int k = 0; k < q; ++k) inj[k] = make_double2(0.0, 0.0);

    for (;;) {
        // ================= A: refill, phasors =================
        if (isnew) {
            itc = 0;
            const bool live = sc >= 0;
            const size_t so = live ? (size_t)sc : 0;
            if (FUND) {
                const double* pm = a.V_m + so;
                const double* pa = a.V_a + so;
                const double* pp = a.P + so;
                const double* pq = a.Q + so;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    Vm[i] = live ? __ldcs(pm) : 1.0;
                    Va[i] = live ? __ldcs(pa) : 0.0;
                    pm += B; pa += B;
                }
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    sP[i * T + lane] = live ? __ldcs(pp) : 0.0;
                    sQ[i * T + lane] = live ? __ldcs(pq) : 0.0;
                    pp += B; pq += B;
                }
                stat = live ? a.status[so] : 0;
            } else {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    Vm[i] = live ? 0.1 : 1.0;                       // flat start of the harmonics (HG:183)
                    Va[i] = 0.0;
                }
            }
            {
                const double2* pin = IN0 + so;
#pragma unroll
                for (int k = 0; k < q; ++k) {
                    IN[k] = live ? __ldcs(pin) : make_double2(0.0, 0.0);
                    pin += (size_t)H * B;
                }
            }
            // w_N rows of this thread: thread-private columns of sW, fetched asynchronously (consumed in
            // phase C; warp 0's rows z < q in phase B)
            {
                const double2* pw = wN0 + so;
                double2* sw = sW + (hv * n - m) * T + lane;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    if (!FUND || i >= m) {
                        if (live) hw_cp_async16(sw, pw);
                        else *sw = make_double2(0.0, 0.0);
                    }
                    pw += B; sw += T;
                }
            }
        }
        cp_async_commit();
        {
            double sn_[n], cs_[n];
            sincos_group<n>(Va, sn_, cs_);
#pragma unroll
            for (int i = 0; i < n; ++i) V[i] = make_double2(Vm[i] * cs_[i], Vm[i] * sn_[i]);   // V = V_m e^{j theta} (HG:403)
        }
        if (COUPLED) {
#pragma unroll
            for (int k = 0; k < q; ++k) sVnl[(k * H + hv) * T + lane] = V[m + k];
        }
        __syncthreads();
        // ================= B: mismatch rows (HG:360-388); border system =================
        long long mxb = 0;
        auto yrow = [&](const int h, const int i) {
            double2 f = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < n; ++j) f = cfma(f, FUND ? C.Y[(h * n + i) * n + j] : cY[(h * n + i) * n + j], V[j]);
            return f;
        };
        // nonlinear bus k: (Y_h V_h)_i + I_N - sum_p Y_N[h][p] V_p,i   (HG:313-323,335-354)
        auto nlrow = [&](const int h, const int k, double2 f) {
            double2 acc;
            if (COUPLED) {
                // two independent accumulators (even / odd p): the 2 H-deep dependent DFMA chain of a
                // single accumulator is what the warp would otherwise wait on
                double2 a0 = make_double2(0.0, 0.0), a1 = a0;
                const double2* sv = sVnl + (k * H) * T + lane;
                const double2* yn = FUND ? C.YNk + (k * H + h) * H : cYN + (k * H + h) * H;
                if constexpr (FUND) {
#pragma unroll
                    for (int p = 0; p + 1 < H; p += 2) {
                        a0 = cfma(a0, yn[p], sv[p * T]);
                        a1 = cfma(a1, yn[p + 1], sv[(p + 1) * T]);
                    }
                    if (H & 1) a0 = cfma(a0, yn[H - 1], sv[(H - 1) * T]);
                } else {
                    // software-pipelined: the LDS.128 pairs of terms p + 2, p + 3 are issued before the FMAs of
                    // terms p, p + 1 (ptxas otherwise emits every pair right in front of its consumer: 41 % of
                    // the harmonic warps' stall samples were short-scoreboard waits in this phase); volatile loads
                    // keep their program order.  Measured: 0.645 -> 0.641 ms per 65,536 scenarios - the other 25
                    // warps of the SM were already covering those waits; bit-identical
                    double2 y0 = hw_lds128(yn), s0 = hw_lds128(sv), y1 = hw_lds128(yn + 1), s1 = hw_lds128(sv + T);
#pragma unroll
                    for (int p = 0; p + 1 < H; p += 2) {
                        double2 y2 = y0, s2 = s0, y3 = y1, s3 = s1;
                        if (p + 2 < H) { y2 = hw_lds128(yn + p + 2); s2 = hw_lds128(sv + (p + 2) * T); }
                        if (p + 3 < H) { y3 = hw_lds128(yn + p + 3); s3 = hw_lds128(sv + (p + 3) * T); }
                        a0 = cfma(a0, y0, s0);
                        a1 = cfma(a1, y1, s1);
                        y0 = y2; s0 = s2; y1 = y3; s1 = s3;
                    }
                    if (H & 1) a0 = cfma(a0, y0, s0);
                }
                acc = cadd(a0, a1);
            } else {
                acc = cmul(FUND ? C.YNk[k * H + h] : cYN[k * H + h], V[m + k]);
            }
            inj[k] = make_double2(IN[k].x - acc.x, IN[k].y - acc.y);
            f = cadd(f, inj[k]);
            absmax_bits(mxb, f.x);
            absmax_bits(mxb, f.y);
        };
        double xF[nx > 0 ? nx : 1];
        if constexpr (!FUND) {
            // (a per-harmonic compile-time copy of this block - immediate constant operands, no loads -
            // was measured 2.3x SLOWER: 13 code paths per CTA thrash the instruction cache)
            const int hh = hv;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const double2 f = yrow(hh, i);
                if (i >= m) {
                    nlrow(hh, i - m, f);
                } else {
                    absmax_bits(mxb, f.x);
                    absmax_bits(mxb, f.y);
                }
            }
        } else {
            constexpr int hh = 0;
#pragma unroll
            for (int k = 0; k < q; ++k) nlrow(hh, k, yrow(hh, m + k));
            cp_async_wait<0>();                                  // w_N rows z < q of a refilled lane
            double M[nx > 0 ? nx : 1][nx + 1];
#pragma unroll
            for (int r = 0; r < nx; ++r)
#pragma unroll
                for (int cc = 0; cc <= nx; ++cc) M[r][cc] = 0.0;
            double rvm[m];
#pragma unroll
            for (int j = c; j < m; ++j) rvm[j] = rcp_fast(Vm[j]);
#pragma unroll
            for (int i = 1; i < m; ++i) {
                const double2 I1 = yrow(hh, i);
                const double2 s_i = cmul(V[i], cconj(I1));                   // V_i conj(I_i)
                const double pl = sP[i * T + lane], ql = sQ[i * T + lane];
                const double2 fs = make_double2(pl + s_i.x, ql + s_i.y);    // power mismatch (HG:372-380)
                absmax_bits(mxb, fs.x);
                if (i >= c) absmax_bits(mxb, fs.y);
                const int rp = i - 1, rq = nth + (i - c);
                double2 ri = make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    const double2 pv = cmul(C.Yeff[i * m + j], V[j]);
                    ri = cadd(ri, pv);
                    if (j >= 1) {
                        const double2 t = cmul(V[i], cconj(pv));             // t_ij = V_i conj(Yeff_ij V_j)
                        const double2 d = (i == j) ? csub(s_i, t) : cneg(t);
                        M[rp][j - 1] = -d.y;                                 // j ([i=j] s_i - t_ij)
                        if (i >= c) M[rq][j - 1] = d.x;
                        if (j >= c) {
                            const double2 e = (i == j) ? cadd(s_i, t) : t;
                            M[rp][nth + j - c] = e.x * rvm[j];
                            if (i >= c) M[rq][nth + j - c] = e.y * rvm[j];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < q; ++k) {
                    const double2 w = sW[k * T + lane];
                    ri = cfma(ri, cneg(C.Y[i * n + m + k]), w);
                }
                const double2 vr = cmul(V[i], cconj(ri));
                M[rp][nx] = -pl - vr.x;
                if (i >= c) M[rq][nx] = -ql - vr.y;
            }
            int badp = 0;
            if constexpr (nx > 0) badp = gauss_regs<nx>(M, xF);
            if (badp) stat |= 0x200;
            sTot[lane] = V[0];
#pragma unroll
            for (int i = 1; i < m; ++i) {
                const double dth = xF[i - 1];
                const double dvm = (i >= c) ? xF[nth + i - c] : 0.0;
                const double ex = (i >= c) ? V[i].x * rvm[i] : 0.0, ey = (i >= c) ? V[i].y * rvm[i] : 0.0;
                // V_i + u_F,i,  u_F = (j V_i) dtheta + E_i dV_m
                sTot[i * T + lane] = make_double2(V[i].x - V[i].y * dth + ex * dvm, V[i].y + V[i].x * dth + ey * dvm);
            }
        }
        atomicMax(&red[cur * T + lane], (unsigned long long)mxb);
        __syncthreads();
        // ================= C: decisions, finished lanes, update, refill =================
        const double err = __longlong_as_double((long long)red[cur * T + lane]);
        const bool active = sc >= 0;
        const bool cont = active && (err > a.thresh_h) && (itc < a.max_h);
        const unsigned actmask = __ballot_sync(0xffffffffu, active);
        if (actmask == 0u) break;                                     // same lanes in every warp: uniform over the CTA
        // Finished lanes are written out and refilled in BATCHES: every warp executes the write-out /
        // refill code (and the CTA the queue atomic and a third barrier) whenever ANY lane is serviced,
        // which would be 4 rounds out of 5 with lanes serviced one by one.  A finished lane keeps its
        // state frozen (cont stays false, its mismatch is recomputed bit-identically) until the next
        // service round: every `epoch`-th round, or at once when 8 lanes wait or nothing else is left.
        unsigned donemask = __ballot_sync(0xffffffffu, active && !cont);
        ++rc;
        if (!(rc >= a.epoch || donemask == actmask || __popc(donemask) >= 8)) donemask = 0u;
        if (donemask) rc = 0;
        const bool done = (donemask >> lane) & 1u;
        // the queue position of the refills is requested first: its latency hides behind the update
        int claimed = 0;
        if (donemask && threadIdx.x == 0) claimed = atomicAdd(a.work_counter, __popc(donemask));
        if (FUND) {
            red[(cur ^ 1) * T + lane] = 0ull;                         // next round's accumulator
            if (a.hist_h && active) a.hist_h[(size_t)itc * B + sc] = err;
        }
        if (done) {
            // post-processing (HG:547-549) + write-out of the thread's n rows
            double* om = a.V_m + row0 + sc;
            double* oa = a.V_a + row0 + sc;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                double vm = Vm[i], va = Va[i];
                if (!(a.flags & HPF_SOLVE_RAW)) {
                    if (vm < 0.0) va += CUDART_PI;
                    va = mod_twopi_fast(va);
                    if (vm < 0.0) vm = -vm;
                }
                *om = vm;
                *oa = va;
                om += B; oa += B;
            }
            if (a.I_inj) {
                double2* oj = a.I_inj + (size_t)hv * B + sc;
#pragma unroll
                for (int k = 0; k < q; ++k) {
                    *oj = inj[k];
                    oj += (size_t)H * B;
                }
            }
            if (FUND) {
                int st = stat & 0xff;
                if ((stat & 0x200) && st == HPF_ST_CONVERGED) st = HPF_ST_SINGULAR;
                if (itc >= a.max_h && st == HPF_ST_CONVERGED) st = HPF_ST_MAXITER;
                if (!(err < CUDART_INF)) st = HPF_ST_NONFINITE;           // NaN or Inf mismatch
                a.n_iter_h[sc] = itc;
                a.err_h[sc] = err;
                a.status[sc] = st;
            }
        }
        if (!FUND) cp_async_wait<0>();                               // w_N rows of a refilled lane
        {
            double2 tot[m];
#pragma unroll
            for (int i = 0; i < m; ++i) tot[i] = sTot[i * T + lane];
            auto update = [&](const int hc) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const int z0 = hc * n - m;
                    if (!FUND || i >= m) {
                        const int z = z0 + i;
                        // u_z = -V_z - sum_i G[z][i] (V_i + u_F,i) - w_N,z ; polar conversion; update
                        const double2 w = sW[z * T + lane];
                        double2 acc = make_double2(V[i].x + w.x, V[i].y + w.y);
#pragma unroll
                        for (int j = 0; j < m; ++j) acc = cfma(acc, FUND ? C.G[z * m + j] : cG[z * m + j], tot[j]);
                        // conj(E) u with E = V / V_m:  dV_m = Re(conj(V) u) / V_m,  dtheta = Im(conj(V) u) / V_m^2
                        const double2 cv = cmul(make_double2(V[i].x, -V[i].y), cneg(acc));
                        const double rv = rcp_fast(Vm[i]);
                        if (cont) {
                            Va[i] += (cv.y * rv) * rv;
                            Vm[i] += cv.x * rv;
                        }
                    } else if (i >= 1) {                                   // linear bus at the fundamental
                        if (cont) {
                            Va[i] += xF[i - 1];
                            if (i >= c) Vm[i] += xF[nth + i - c];
                        }
                    }
                }
            };
            if constexpr (FUND) update(0);
            else update(hv);
            if (cont) ++itc;
        }
        isnew = false;
        if (donemask) {
            if (threadIdx.x == 0) sbase[0] = claimed;
            __syncthreads();
            if (done) {
                const int idx = sbase[0] + __popc(donemask & lt);
                sc = ((size_t)idx < B) ? idx : -1;
                isnew = true;
            }
        }
        cur ^= 1;
    }
}

template <class D, bool COUPLED, int MINB>
__global__ void __launch_bounds__(D::H * 32, MINB)
harm_hw_kernel(const __grid_constant__ HwConsts<D, COUPLED> C, const HarmTileArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int T = 32;
    unsigned long long* red = reinterpret_cast<unsigned long long*>(
        smem + (2 * D::q * D::H + 2 * (D::n * D::H - D::m) + 2 * D::m + 2 * D::m) * T);
    int* sbase = reinterpret_cast<int*>(red + 2 * T);
    if (threadIdx.x == 0) sbase[0] = atomicAdd(a.work_counter, T);
    if (threadIdx.x < 2 * T) red[threadIdx.x] = 0ull;
    {
        double2* cc = reinterpret_cast<double2*>(sbase + 16);
        constexpr int nY = D::H * D::n * D::n, nYN = COUPLED ? D::q * D::H * D::H : D::q * D::H;
        constexpr int nG = (D::n * D::H - D::m) * D::m;
        for (int t = threadIdx.x; t < nY; t += D::H * 32) cc[t] = C.Y[t];
        for (int t = threadIdx.x; t < nYN; t += D::H * 32) cc[nY + t] = C.YNk[t];
        for (int t = threadIdx.x; t < nG; t += D::H * 32) cc[nY + nYN + t] = C.G[t];
    }
    __syncthreads();
    const int hv = threadIdx.x >> 5;
    if (hv == 0) harm_hw_loop<D, COUPLED, true>(C, a, smem, hv);
    else harm_hw_loop<D, COUPLED, false>(C, a, smem, hv);
}
