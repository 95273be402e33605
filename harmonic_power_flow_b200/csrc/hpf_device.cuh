// Device-side building blocks shared by every kernel of the harmonic power-flow path.
//
// Everything here is FP64.  A "scenario view" addresses the per-scenario state with a
// stride VS and an offset `off`:
//   * fused / per-CTA kernels:  VS = 1,  off = 0   (one scenario per CTA, state in smem)
//   * tile kernels:             VS = 32, off = lane (32 scenarios per CTA, lane = scenario)
// so the same arithmetic (same operation order => same bits) is used by the standalone
// mismatch kernel and by the fused Newton kernel.
//
// Reference citations: HG = "Harmonic Power Flow/hcne_generalized.py" of the reference.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

struct DevNet {
    int n, m, c, H, q, nH, N, Nf, coupled;
    const double2* Y;       // [H][n][n]   bus admittance per harmonic (kernel 1 output)
    const double2* YN;      // [n_dev][H][H] (coupled) or [n_dev][H]
    const int* dev_of_nl;   // [q]
    // optional sparsity pattern of Y(h) (union over the harmonics), ELL format: the columns of
    // row i are ell_col[i * ell_w + e], ascending, padded with -1; nullptr = use the dense rows.
    // Built for the larger networks only; skipping exact zeros of a sequential sum leaves
    // every finite result bit-identical.
    const int* ell_col;
    int ell_w;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + a*b with fused multiply-adds (4 FP64 instructions per complex MAC)
__device__ __forceinline__ double2 cfma(double2 acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
    return acc;
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cneg(double2 a) { return make_double2(-a.x, -a.y); }
// multiplication by j is exact: j(a + jb) = -b + ja
__device__ __forceinline__ double2 cmulj(double2 a) { return make_double2(-a.y, a.x); }
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

// sin / cos with the network's phasor angles: Cody-Waite reduction by pi/2 in three parts and
// the classic degree-13 / degree-14 minimax polynomials on [-pi/4, pi/4] (fdlibm's
// coefficients), error < 1 ulp.  Coefficients are constant-bank operands (the library
// routine materialises them as immediates, two extra issue slots per coefficient).  Beyond
// |x| >= 105615 the three-part reduction loses accuracy: the library routine is called.
__constant__ double HPF_SC[20] = {
    6.36619772367581382433e-01,   // 0  2/pi
    6755399441055744.0,           // 1  1.5 * 2^52 (round-to-nearest-integer by addition)
    1.57079632679489655800e+00,   // 2  pi/2 hi
    6.12323399573676603587e-17,   // 3  pi/2 mid
    -1.49738490485916983169e-33,  // 4  pi/2 lo (pi/2 - hi - mid)
    -1.66666666666666324348e-01,  // 5  S1
    8.33333333332248946124e-03,   // 6  S2
    -1.98412698298579493134e-04,  // 7  S3
    2.75573137070700676789e-06,   // 8  S4
    -2.50507602534068634195e-08,  // 9  S5
    1.58969099521155010221e-10,   // 10 S6
    4.16666666666666019037e-02,   // 11 C1
    -1.38888888888741095749e-03,  // 12 C2
    2.48015872894767294178e-05,   // 13 C3
    -2.75573143513906633035e-07,  // 14 C4
    2.08757232129817482790e-09,   // 15 C5
    -1.13596475577881948265e-11,  // 16 C6
    105615.0,                     // 17 fast-path bound
    0.0, 0.0};

__device__ __noinline__ double2 sincos_slow(double x) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(s, c);
}

// Branch-free fast path (valid for |x| < 105615; garbage-in-garbage-out beyond, NaN for NaN).
__device__ __forceinline__ void sincos_core(const double x, double& sn, double& cs) {
    const double t = fma(x, HPF_SC[0], HPF_SC[1]);
    const int k = __double2loint(t);
    const double qd = t - HPF_SC[1];
    double r = fma(-qd, HPF_SC[2], x);
    r = fma(-qd, HPF_SC[3], r);
    r = fma(-qd, HPF_SC[4], r);
    const double z = r * r;
    double ps = fma(z, HPF_SC[10], HPF_SC[9]);
    double pc = fma(z, HPF_SC[16], HPF_SC[15]);
    ps = fma(z, ps, HPF_SC[8]);
    pc = fma(z, pc, HPF_SC[14]);
    ps = fma(z, ps, HPF_SC[7]);
    pc = fma(z, pc, HPF_SC[13]);
    ps = fma(z, ps, HPF_SC[6]);
    pc = fma(z, pc, HPF_SC[12]);
    ps = fma(z, ps, HPF_SC[5]);
    pc = fma(z, pc, HPF_SC[11]);
    const double s0 = fma(z * r, ps, r);                  // r + r^3 (S1 + z S2 + ...)
    const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));  // 1 - z/2 + z^2 (C1 + z C2 + ...)
    const bool swap = (k & 1) != 0;
    const double a = swap ? c0 : s0;                      // sin(x) up to sign
    const double b = swap ? s0 : c0;                      // cos(x) up to sign
    // signs by integer XOR on the high word (keeps the FP64 pipe free)
    sn = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
    cs = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
}

// N independent sin/cos pairs: all fast paths first (independent dependency chains that the
// scheduler interleaves), one rare fix-up branch for large arguments afterwards.
template <int N>
__device__ __forceinline__ void sincos_group(const double* x, double* sn, double* cs) {
    bool big = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        sincos_core(x[i], sn[i], cs[i]);
        big |= (fabs(x[i]) >= HPF_SC[17]);                // also +-Inf; NaN propagates through the core
    }
    if (big) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (fabs(x[i]) >= HPF_SC[17]) {
                const double2 r2 = sincos_slow(x[i]);
                sn[i] = r2.x;
                cs[i] = r2.y;
            }
    }
}

// ---------------------------------------------------------------------------------------
// Phasors.  V = V_m e^{j theta};  E = V / V_m (harmonic stage, SIGNED magnitude, HG:405,455)
// or E = V / |V| (fundamental stage, HG:210).
template <int VS>
__device__ __forceinline__ void phasor_one(int s, int off, const double* Vm, const double* Va,
                                           double* Vre, double* Vim, double* Ere, double* Eim,
                                           bool abs_norm) {
    const double vm = Vm[s * VS + off];
    double sn, cs;
    sincos(Va[s * VS + off], &sn, &cs);
    const double re = vm * cs, im = vm * sn;
    Vre[s * VS + off] = re;
    Vim[s * VS + off] = im;
    const double d = abs_norm ? hypot(re, im) : vm;
    const double r = 1.0 / d;          // numpy divides complex by real as x * (1/d)
    Ere[s * VS + off] = re * r;
    Eim[s * VS + off] = im * r;
}

// ---------------------------------------------------------------------------------------
// Row dot products.  I1_i = sum_j Y1[i][j] V1[j]  (HG:208,339,379,453); sequential in j.
template <int VS>
__device__ __forceinline__ double2 ydotv(const DevNet& net, int h, int i, int off,
                                         const double* Vre, const double* Vim) {
    const int n = net.n;
    const double2* Yrow = net.Y + ((size_t)h * n + i) * n;
    double2 acc = make_double2(0.0, 0.0);
    if (net.ell_col) {
        const int* cols = net.ell_col + (size_t)i * net.ell_w;
        for (int e = 0; e < net.ell_w; ++e) {
            const int j = __ldg(cols + e);
            if (j < 0) break;
            const double2 y = ldg2(Yrow + j);
            const double2 v = make_double2(Vre[(h * n + j) * VS + off], Vim[(h * n + j) * VS + off]);
            acc = cfma(acc, y, v);
        }
        return acc;
    }
    for (int j = 0; j < n; ++j) {
        const double2 y = ldg2(Yrow + j);
        const double2 v = make_double2(Vre[(h * n + j) * VS + off], Vim[(h * n + j) * VS + off]);
        acc = cfma(acc, y, v);
    }
    return acc;
}

// Norton contraction: I_inj[k][h] = I_N[k][h] - sum_p Y_N[dev(k)][h][p] V[p][m+k]  (HG:313-323)
template <int VS>
__device__ __forceinline__ double2 norton_injection(const DevNet& net, int k, int h, int off,
                                                    const double* Vre, const double* Vim,
                                                    double2 I_N) {
    const int n = net.n, H = net.H, bus = net.m + k;
    const int dev = net.dev_of_nl[k];
    double2 acc;
    if (net.coupled) {
        const double2* row = net.YN + ((size_t)dev * H + h) * H;
        acc = make_double2(0.0, 0.0);
        for (int p = 0; p < H; ++p) {
            const double2 v = make_double2(Vre[(p * n + bus) * VS + off], Vim[(p * n + bus) * VS + off]);
            acc = cfma(acc, ldg2(row + p), v);
        }
    } else {
        const double2 v = make_double2(Vre[(h * n + bus) * VS + off], Vim[(h * n + bus) * VS + off]);
        acc = cmul(ldg2(net.YN + (size_t)dev * H + h), v);
    }
    return csub(I_N, acc);
}

// One complex entry e of the harmonic mismatch f_c (HG:360-388); stacked index s = e + 1.
//   s <  m : dS = (P + jQ)_s + V_s conj(I1_s)                       (HG:372-380)
//   s >= m : dI = (Y_h V_h)_i + I_inj[i-m][h] at nonlinear buses     (HG:335-354)
// I1 [n] and Iinj [q*H] are per-scenario arrays laid out like the state (stride VS).
template <int VS>
__device__ __forceinline__ double2 harmonic_mismatch_entry(const DevNet& net, int e, int off,
                                                           const double* Vre, const double* Vim,
                                                           const double2* I1, const double2* Iinj,
                                                           const double* P, const double* Q) {
    const int n = net.n, m = net.m, s = e + 1;
    if (s < m) {
        const double2 v = make_double2(Vre[s * VS + off], Vim[s * VS + off]);
        const double2 sl = cmul(v, cconj(I1[s * VS + off]));
        return make_double2(P[s * VS + off] + sl.x, Q[s * VS + off] + sl.y);
    }
    const int h = s / n, i = s - h * n;
    double2 acc = (h == 0) ? I1[i * VS + off] : ydotv<VS>(net, h, i, off, Vre, Vim);
    if (i >= m) acc = cadd(acc, Iinj[((i - m) * net.H + h) * VS + off]);
    return acc;
}

// Row/column maps of the harmonic system (HG:388,397,445-446,464-472).
__device__ __forceinline__ int h_row_re(const DevNet& net, int e) { return e; }
__device__ __forceinline__ int h_row_im(const DevNet& net, int e) { return (net.nH - 1) + e - (net.c - 1); }
__device__ __forceinline__ int h_col_a(const DevNet& net, int t) { return t - 1; }
__device__ __forceinline__ int h_col_v(const DevNet& net, int t) { return (net.nH - 1) + t - net.c; }

// ---------------------------------------------------------------------------------------
// Jacobian element generator (HG:401-473).  Work item (e, slot):
//   slot <  n : column phasor (h, j = slot) of the same harmonic block
//   slot >= n : column phasor (p = slot - n, i) - the Norton coupling Y_N[h][p], p != h
// writes up to 4 real entries through `put(row, col, value)`.
template <class Put>
__device__ __forceinline__ void harmonic_jacobian_item(const DevNet& net, int e, int slot,
                                                       const double* Vre, const double* Vim,
                                                       const double* Ere, const double* Eim,
                                                       const double2* I1, Put put) {
    const int n = net.n, m = net.m, c = net.c, H = net.H;
    const int s = e + 1, h = s / n, i = s - h * n;
    int t;
    double2 dA, dV;
    if (slot < n) {
        const int j = slot;
        t = h * n + j;
        const double2 y = ldg2(net.Y + ((size_t)h * n + i) * n + j);
        const double2 vt = make_double2(Vre[t], Vim[t]);
        const double2 et = make_double2(Ere[t], Eim[t]);
        if (s < m) {
            // power rows (HG:451-467): dSdA = jV_i conj(I_i d_ij - Y_ij V_j);
            //                          dSdV = E_i conj(I_i) d_ij + V_i conj(Y_ij E_j)
            if (i != j && y.x == 0.0 && y.y == 0.0) return;
            const double2 vi = make_double2(Vre[s], Vim[s]);
            const double2 ei = make_double2(Ere[s], Eim[s]);
            const double2 yv = cmul(y, vt);
            const double2 d = (i == j) ? csub(I1[i], yv) : cneg(yv);
            dA = cmul(cmulj(vi), cconj(d));
            dV = cmul(vi, cconj(cmul(y, et)));
            if (i == j) dV = cadd(cmul(ei, cconj(I1[i])), dV);
        } else {
            // current rows (HG:403-411,425-443): dIdV = Y E - Y_N E ; dIdA = (jY) V - (jY_N) V
            const bool nl_diag = (i >= m) && (j == i);
            if (!nl_diag && y.x == 0.0 && y.y == 0.0) return;
            dV = cmul(y, et);
            dA = cmul(cmulj(y), vt);
            if (nl_diag) {
                const int dev = net.dev_of_nl[i - m];
                const double2 yn = net.coupled ? ldg2(net.YN + ((size_t)dev * H + h) * H + h)
                                               : ldg2(net.YN + (size_t)dev * H + h);
                dV = csub(dV, cmul(yn, et));
                dA = csub(dA, cmul(cmulj(yn), vt));
            }
        }
    } else {
        const int p = slot - n;
        if (!net.coupled || s < m || i < m || p == h) return;
        t = p * n + i;
        const int dev = net.dev_of_nl[i - m];
        const double2 yn = ldg2(net.YN + ((size_t)dev * H + h) * H + p);
        const double2 vt = make_double2(Vre[t], Vim[t]);
        const double2 et = make_double2(Ere[t], Eim[t]);
        dV = cneg(cmul(yn, et));
        dA = cneg(cmul(cmulj(yn), vt));
    }
    const bool has_im = (e >= c - 1);
    const int rr = h_row_re(net, e), ri = h_row_im(net, e);
    if (t >= 1) {
        put(rr, h_col_a(net, t), dA.x);
        if (has_im) put(ri, h_col_a(net, t), dA.y);
    }
    if (t >= c) {
        put(rr, h_col_v(net, t), dV.x);
        if (has_im) put(ri, h_col_v(net, t), dV.y);
    }
}

// Fundamental Jacobian element (i, j) (HG:205-223), abs-normalised E.
template <class Put>
__device__ __forceinline__ void fund_jacobian_item(const DevNet& net, int i, int j,
                                                   const double* Vre, const double* Vim,
                                                   const double* Ere, const double* Eim,
                                                   const double2* I1, Put put) {
    const int n = net.n, c = net.c;
    const double2 y = ldg2(net.Y + (size_t)i * n + j);
    if (i != j && y.x == 0.0 && y.y == 0.0) return;
    const double2 vi = make_double2(Vre[i], Vim[i]), ei = make_double2(Ere[i], Eim[i]);
    const double2 vj = make_double2(Vre[j], Vim[j]), ej = make_double2(Ere[j], Eim[j]);
    const double2 yv = cmul(y, vj);
    const double2 d = (i == j) ? csub(I1[i], yv) : cneg(yv);
    const double2 dA = cmul(cmulj(vi), cconj(d));
    double2 dV = cmul(vi, cconj(cmul(y, ej)));
    if (i == j) dV = cadd(cmul(ei, cconj(I1[i])), dV);
    const int rr = i - 1, ri = (n - 1) + i - c;
    const int ca = j - 1, cv = (n - 1) + j - c;
    if (i >= 1) {
        if (j >= 1) put(rr, ca, dA.x);
        if (j >= c) put(rr, cv, dV.x);
    }
    if (i >= c) {
        if (j >= 1) put(ri, ca, dA.y);
        if (j >= c) put(ri, cv, dV.y);
    }
}

// ---------------------------------------------------------------------------------------
// Block-wide max-reduction that PROPAGATES NaN (numpy's max / norm(inf) do, HG:201,389).
__device__ __forceinline__ double block_max_nan(double v, double* red /* >= 33 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int bad = (v != v);
    double x = bad ? 0.0 : v;
    for (int o = 16; o; o >>= 1) {
        x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __syncthreads();                      // red[] may still be read from a previous call
    if (lane == 0) red[warp] = bad ? CUDART_NAN : x;
    __syncthreads();
    if (warp == 0) {
        double y = (lane < nw) ? red[lane] : 0.0;
        int b2 = (y != y);
        y = b2 ? 0.0 : y;
        for (int o = 16; o; o >>= 1) {
            y = fmax(y, __shfl_xor_sync(0xffffffffu, y, o));
            b2 |= __shfl_xor_sync(0xffffffffu, b2, o);
        }
        if (lane == 0) red[32] = b2 ? CUDART_NAN : y;
    }
    __syncthreads();
    return red[32];
}

// ---------------------------------------------------------------------------------------
// Dense LU with partial pivoting + triangular solve of the augmented system [A | b] held
// in shared memory, COLUMN-MAJOR with an ODD leading dimension (odd stride in doubles =>
// row-direction and column-direction accesses are both bank-conflict free for 64-bit
// words).  Column N of A is the right-hand side; on return it holds the solution.
//   * pivot search: every warp redundantly scans column k (lanes over rows) and reduces
//     (|value|, row) with warp shuffles - no block barrier is needed to publish the pivot;
//   * row swap is fused into the trailing update: the warp that owns column j swaps its
//     two entries, then applies the rank-1 update with multipliers kept in registers;
//   * ONE __syncthreads per elimination step.
// Returns (to all threads) 0 or k+1 for a zero / non-finite pivot at step k.
#define HPF_LU_MAXCHUNK 6   // rows per lane: N <= 32*6 = 192 (shared memory caps N near 165)

__device__ __forceinline__ int lu_solve_smem(double* A, const int N, const int ld, double* rinv,
                                             int* sflag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) *sflag = 0;
    for (int k = 0; k < N; ++k) {
        __syncthreads();
        const double* colk = A + (size_t)k * ld;
        // --- pivot search (idamax semantics: first maximal |a_ik|) ---
        double best = -1.0;
        int bi = k;
        for (int i = k + lane; i < N; i += 32) {
            const double v = fabs(colk[i]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        const int p = bi;
        const double piv = colk[p];
        const double akk = colk[k];
        if (!(best > 0.0) || !(best < CUDART_INF)) {
            if (tid == 0 && *sflag == 0) *sflag = k + 1;
        }
        const double r = 1.0 / piv;
        if (tid == 0) rinv[k] = r;
        // multipliers of my rows (after the swap row p holds the old row k)
        double l[HPF_LU_MAXCHUNK];
#pragma unroll
        for (int ch = 0; ch < HPF_LU_MAXCHUNK; ++ch) {
            const int i = k + 1 + lane + 32 * ch;
            l[ch] = (i < N) ? ((i == p) ? akk : colk[i]) * r : 0.0;
        }
        // --- trailing update, one column per warp at a time (incl. the rhs column N) ---
        for (int j = k + 1 + warp; j <= N; j += nw) {
            double* colj = A + (size_t)j * ld;
            const double ukj = colj[p];
            const double okj = colj[k];
            __syncwarp();
            if (lane == 0) { colj[k] = ukj; colj[p] = okj; }
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < HPF_LU_MAXCHUNK; ++ch) {
                const int i = k + 1 + lane + 32 * ch;
                if (i < N) colj[i] -= l[ch] * ukj;
            }
        }
    }
    __syncthreads();
    // --- back substitution U x = b (column sweep, warp 0; U's diagonal as reciprocals) ---
    if (warp == 0) {
        double* b = A + (size_t)N * ld;
        for (int j = N - 1; j >= 0; --j) {
            const double xj = b[j] * rinv[j];
            __syncwarp();
            if (lane == 0) b[j] = xj;
            const double* colj = A + (size_t)j * ld;
            for (int i = lane; i < j; i += 32) b[i] -= colj[i] * xj;
            __syncwarp();
        }
    }
    __syncthreads();
    return *sflag;
}
