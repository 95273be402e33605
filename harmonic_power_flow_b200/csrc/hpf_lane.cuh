// Standalone mismatch kernel (kernel 2), "lane" variant: ONE THREAD PER SCENARIO, the whole
// row set of harmonic_mismatch() (HG:360-390) streamed harmonic by harmonic through registers.
//
// Why: the 32-scenario tile kernel (mismatch_tile_kernel) spreads the rows of a scenario over
// 8 warps, so every operand of every complex MAC is a shared-memory load and the kernel ends
// up bound by instruction issue (ncu: 11.8 k instructions per scenario, 30 % of them FP64).
// Here a thread owns its scenario: the phasors of the nonlinear buses stay in registers for
// the Norton contraction (HG:313-323), the phasors of one harmonic are live for the Y(h) V
// rows of that harmonic only (HG:335-354,372-380), and every network constant is a
// constant-bank operand of the FP64 instruction itself (kernel parameter space,
// __grid_constant__): no shared memory, no barrier, no load instruction for Y(h) / Y_N.
// HBM traffic is unchanged: batch-innermost arrays, a warp touches 32 consecutive scenarios,
// every global access is one 256-byte (f64) or 512-byte (c128) segment.
//
// Only shape-specialised instances exist (register arrays need compile-time extents); other
// networks use the tile kernel.
#pragma once
#include <type_traits>
#include "hpf_device.cuh"

// NaN-propagating running maximum of |v| on the INTEGER pipe: for sign-cleared IEEE doubles the
// bit patterns order like the magnitudes, and every NaN pattern is above +Inf, so an integer
// max keeps a NaN once it has seen one (numpy's norm(inf) / max do the same, HG:389).
__device__ __forceinline__ void absmax_bits(long long& mxb, const double v) {
    // (inline PTX: written as an AND on the 64-bit pattern the compiler turns the mask into an
    // FP64 |v| = DADD, which is not bit-preserving for NaN and occupies the FP64 pipe)
    long long bits;
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tand.b32 hi, hi, 0x7fffffff;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}" : "=l"(bits) : "d"(v));
    mxb = bits > mxb ? bits : mxb;
}

// Network constants of a shape-specialised instance, passed BY VALUE as a kernel parameter
// (constant bank 0): Y(h) [H][n][n] and the Norton admittances pre-gathered per nonlinear
// bus, YNk [q][H][H] (coupled) or [q][H] in the first q*H entries (uncoupled).
template <class D>
struct LaneConsts {
    double2 Y[D::H * D::n * D::n];
    double2 YNk[D::q * D::H * D::H];
};

#define HPF_LANE_THREADS 64
#define HPF_LANE_STAGES 4

__device__ __forceinline__ void lane_cp_async8(double* sdst, const double* gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void lane_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lane_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

template <class D, bool COUPLED>
__global__ void __launch_bounds__(HPF_LANE_THREADS, 7)
mismatch_lane_kernel(const __grid_constant__ LaneConsts<D> C, const MismatchArgs a) {
    constexpr int n = D::n, m = D::m, c = D::c, H = D::H, q = D::q, nH = n * H;
    // phasors of the nonlinear buses, one column per thread: read back with a RUNTIME harmonic
    // index inside the rolled harmonic loop (the register copy needs compile-time indices)
    __shared__ double2 sVnl[q * H][HPF_LANE_THREADS];
    __shared__ double ring[HPF_LANE_STAGES][2 * m + 2 * q][HPF_LANE_THREADS];
    const size_t B = (size_t)a.B;
    const size_t b0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b0 >= B) return;                                 // no barrier in this kernel: spare lanes just leave
    const size_t b = b0;
    const size_t nB = (size_t)n * B;

    // Input ring: the raw state (V_m, V_a of the linear buses, I_N of the nonlinear ones) of the
    // next HPF_LANE_STAGES harmonics is kept in flight with asynchronous global->shared copies
    // (LDGSTS, one column per thread, no register cost).  With one prefetched harmonic in
    // registers a single resident wave (65,536 scenarios = 14 warps per SM) cannot cover the
    // HBM latency: all warps are in the same phase at the same time.
    const double* pvm = a.V_m + b;                       // row (h + STAGES) n of V_m / V_a (next to request)
    const double* pva = a.V_a + b;
    const double2* pin = a.I_N + b;
    double2* pij = a.I_inj ? a.I_inj + b : nullptr;
    auto request = [&](const int stage) {
        double* slot = &ring[stage][0][threadIdx.x];
#pragma unroll
        for (int j = 0; j < m; ++j) {
            lane_cp_async8(slot + (2 * j) * HPF_LANE_THREADS, pvm + (size_t)j * B);
            lane_cp_async8(slot + (2 * j + 1) * HPF_LANE_THREADS, pva + (size_t)j * B);
        }
#pragma unroll
        for (int k = 0; k < q; ++k) {
            const double* src = reinterpret_cast<const double*>(pin + (size_t)(k * H) * B);
            lane_cp_async8(slot + (2 * m + 2 * k) * HPF_LANE_THREADS, src);
            lane_cp_async8(slot + (2 * m + 2 * k + 1) * HPF_LANE_THREADS, src + 1);
        }
        pvm += nB;
        pva += nB;
        pin += B;
    };
#pragma unroll
    for (int st = 0; st < HPF_LANE_STAGES; ++st) {
        if (st < H) request(st);
        lane_cp_async_commit();
    }
    double Pl[m], Ql[m];
#pragma unroll
    for (int i = 1; i < m; ++i) {
        Pl[i] = __ldcs(a.P + (size_t)i * B + b);
        Ql[i] = __ldcs(a.Q + (size_t)i * B + b);
    }

    // ---- pass 1: phasors of the nonlinear buses at every harmonic (registers + smem) ----
    double2 Vnl[q][H];
#pragma unroll
    for (int k = 0; k < q; ++k) {
        double rm[H], ra[H], sn[H], cs[H];
        const double* pm = a.V_m + (size_t)(m + k) * B + b;
        const double* pa = a.V_a + (size_t)(m + k) * B + b;
#pragma unroll
        for (int p = 0; p < H; ++p) {
            rm[p] = __ldcs(pm);
            ra[p] = __ldcs(pa);
            pm += nB;
            pa += nB;
        }
        constexpr int G4 = 4;
#pragma unroll
        for (int p0 = 0; p0 < H; p0 += G4) {
            if (p0 + G4 <= H) sincos_group<G4>(ra + p0, sn + p0, cs + p0);
            else sincos_group<(H % G4) ? (H % G4) : G4>(ra + p0, sn + p0, cs + p0);
        }
#pragma unroll
        for (int p = 0; p < H; ++p) {
            Vnl[k][p] = make_double2(rm[p] * cs[p], rm[p] * sn[p]);
            sVnl[k * H + p][threadIdx.x] = Vnl[k][p];
        }
    }

    long long mxb = 0;
    double* pf = a.f + b;                                // row e of f (real parts), advanced row by row
    double* pfi = pf + (size_t)((nH - 1) - (c - 1)) * B; // row of the imaginary part of entry e (HG:388)

    // ---- pass 2: harmonic by harmonic (the h >= 1 blocks are ONE rolled loop: fully unrolled
    //      the kernel is 120 KB of straight-line code and starves on instruction fetch); the
    //      raw state of harmonic h+1 is in flight while harmonic h is computed ----
    auto harmonic = [&](auto first_tag, const int h) {
        constexpr bool FIRST = decltype(first_tag)::value;
        double cm[m], ca[m];
        double2 cin[q];
        lane_cp_async_wait<HPF_LANE_STAGES - 1>();        // the group of harmonic h has landed
        {
            const int stage = h % HPF_LANE_STAGES;
            const double* slot = &ring[stage][0][threadIdx.x];
#pragma unroll
            for (int j = 0; j < m; ++j) {
                cm[j] = slot[(2 * j) * HPF_LANE_THREADS];
                ca[j] = slot[(2 * j + 1) * HPF_LANE_THREADS];
            }
#pragma unroll
            for (int k = 0; k < q; ++k)
                cin[k] = make_double2(slot[(2 * m + 2 * k) * HPF_LANE_THREADS],
                                      slot[(2 * m + 2 * k + 1) * HPF_LANE_THREADS]);
            if (h + HPF_LANE_STAGES < H) request(stage);  // refill the slot just consumed
            lane_cp_async_commit();                       // (possibly empty: keeps the group count uniform)
        }
        double2 V[n];
        {
            double sn[m], cs[m];
            sincos_group<m>(ca, sn, cs);
#pragma unroll
            for (int j = 0; j < m; ++j) V[j] = make_double2(cm[j] * cs[j], cm[j] * sn[j]);
        }
#pragma unroll
        for (int k = 0; k < q; ++k) V[m + k] = FIRST ? Vnl[k][0] : sVnl[k * H + h][threadIdx.x];
        const double2* Yh = C.Y + h * (n * n);
        // rows of this harmonic, bus order
#pragma unroll
        for (int i = 0; i < n; ++i) {
            if (FIRST && i == 0) continue;                               // slack: no row
            double2 f = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < n; ++j) f = cfma(f, Yh[i * n + j], V[j]);
            if (i < m) {
                if (FIRST) {                                             // dS = (P + jQ) + V conj(Y1 V)  (HG:372-380)
                    const double2 sl = cmul(V[i], cconj(f));
                    f = make_double2(Pl[i] + sl.x, Ql[i] + sl.y);
                }
            } else {
                // nonlinear bus: (Y_h V_h)_i + I_N - sum_p Y_N[h][p] V_p,i   (HG:313-323,335-354)
                const int k = i - m;
                double2 acc;
                if (COUPLED) {
                    const double2* row = C.YNk + (k * H + h) * H;
                    acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int p = 0; p < H; ++p) acc = cfma(acc, row[p], Vnl[k][p]);
                } else {
                    acc = cmul(C.YNk[k * H + h], V[i]);
                }
                const double2 inj = csub(cin[k], acc);
                if (pij) __stcs(pij + (size_t)(k * H) * B, inj);
                f = cadd(f, inj);
            }
            // emit row e = h n + i - 1 (rows are produced in order)
            __stcs(pf, f.x);
            absmax_bits(mxb, f.x);
            if (!FIRST || i - 1 >= c - 1) {
                __stcs(pfi, f.y);
                absmax_bits(mxb, f.y);
            }
            pf += B;
            pfi += B;
        }
        if (pij) pij += B;
    };
    harmonic(std::true_type{}, 0);
#pragma unroll 1
    for (int h = 1; h < H; ++h) harmonic(std::false_type{}, h);
    a.err[b] = __longlong_as_double(mxb);
}

// Per-scenario constant of the harmonic stage, w_N = W_NL I_N (see hpf_structured.cuh), for the
// shape-specialised 4-bus networks: one thread per scenario, its Norton currents in registers,
// the (nZ x qH) operator a kernel parameter (constant bank), rows in a rolled loop.  Same
// summation order as wn_tile_kernel (bit-identical results), a quarter of its time: the tile
// kernel re-reads every current from shared memory for every row.
template <class D>
struct WnConsts {
    double2 W[(D::n * D::H - D::m) * (D::q * D::H)];
};

template <class D>
__global__ void __launch_bounds__(128)
wn_lane_kernel(const __grid_constant__ WnConsts<D> C, const WnArgs a) {
    constexpr int nZ = D::n * D::H - D::m, qH = D::q * D::H;
    const size_t B = (size_t)a.B;
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double2 in[qH];
#pragma unroll
    for (int u = 0; u < qH; ++u) in[u] = a.I_N[(size_t)u * B + b];
    double2* out = a.wN + b;
#pragma unroll 1
    for (int z = 0; z < nZ; ++z) {
        const double2* row = C.W + z * qH;
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < qH; ++u) acc = cadd(acc, cmul(row[u], in[u]));
        *out = acc;
        out += B;
    }
}
