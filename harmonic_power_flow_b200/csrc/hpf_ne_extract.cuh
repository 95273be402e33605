// Norton-equivalent extraction from simulated measurements, the step before the solve path
// (reference: "Circuit Simulation/NE_from_sim.py" = NE), batched over
// devices / operating points: one CTA per device.
//
//   uncoupled (Thunberg, NE:86-118): per frequency  Y = (I_2 - I_1) / (V_1 - V_2),  I_N = Y V_1 + I_1
//   coupled   (Almeida,  NE:141-173): the N+1 measurements (fundamental 1, one harmonic source
//     at a time, fundamental 2) satisfy  I_k[f] = I_N[f] - sum_p Y_N[f][p] V_k[p];  with the
//     measurement matrix M = [-V | 1] this is  M X = I,  X = [Y_N^T ; I_N]: ONE (N+1) x (N+1)
//     complex system with N right-hand sides.  The reference inverts M (numpy.linalg.inv) and
//     multiplies; here the augmented system [M | I] is reduced by Gauss-Jordan elimination with
//     partial pivoting in shared memory.
// Complex numbers are interleaved double2.
#pragma once
#include "hpf_device.cuh"
#include "hpf_structured.cuh"     // crecip

struct NeExtractArgs {
    int D, N;                       // devices, frequencies (N = K + 1)
    const double2* Vf;              // [D][2]
    const double2* Vh;              // [D][2][K]
    const double2* I_f;             // [D][2][N]
    const double2* I_h;             // [D][2][K][N]
    double2 *Y_N_c, *I_N_c, *Y_N_uc, *I_N_uc;   // [D][N][N], [D][N], [D][N], [D][N]
    int* info;                      // [D] 0 or k+1 (zero pivot at step k)
};

__host__ __device__ inline size_t ne_extract_smem_bytes(int N) {
    return (size_t)(N + 1) * (2 * N + 1) * sizeof(double2) + 64;
}

__global__ void __launch_bounds__(256)
ne_extract_kernel(const NeExtractArgs a) {
    extern __shared__ __align__(16) double smem_ne[];
    const int N = a.N, K = N - 1, R = N + 1, W = 2 * N + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double2* A = reinterpret_cast<double2*>(smem_ne);          // [R][W] row-major: [M | I]
    int* sflag = reinterpret_cast<int*>(A + (size_t)R * W);   // [0] pivot row, [1] info
    for (int d = blockIdx.x; d < a.D; d += gridDim.x) {
        const double2* Vf = a.Vf + (size_t)d * 2;
        const double2* Vh = a.Vh + (size_t)d * 2 * K;
        const double2* I_f = a.I_f + (size_t)d * 2 * N;
        const double2* I_h = a.I_h + (size_t)d * 2 * K * N;
        __syncthreads();
        // ---- uncoupled equivalents (NE:86-118) ----
        for (int f = tid; f < N; f += blockDim.x) {
            double2 v1, v2, i1, i2;
            if (f == 0) { v1 = Vf[0]; v2 = Vf[1]; i1 = I_f[0]; i2 = I_f[N]; }
            else {
                const int k = f - 1;
                v1 = Vh[k]; v2 = Vh[K + k];
                i1 = I_h[(size_t)k * N + f]; i2 = I_h[((size_t)K + k) * N + f];
            }
            const double2 y = cmul(csub(i2, i1), crecip(csub(v1, v2)));
            a.Y_N_uc[(size_t)d * N + f] = y;
            a.I_N_uc[(size_t)d * N + f] = cadd(cmul(y, v1), i1);
        }
        // ---- augmented system [M | I] (NE:145-167) ----
        for (int t = tid; t < R * W; t += blockDim.x) {
            const int r = t / W, cc = t - r * W;
            double2 v = make_double2(0.0, 0.0);
            if (cc < N) {                                          // -V_mes
                if (cc == 0) v = cneg((r == N) ? Vf[1] : Vf[0]);
                else if (r == cc) v = cneg(Vh[cc - 1]);
            } else if (cc == N) {
                v = make_double2(1.0, 0.0);
            } else {                                               // measured spectra
                const int f = cc - R;
                v = (r == 0) ? I_f[f] : (r == N) ? I_f[N + f] : I_h[(size_t)(r - 1) * N + f];
            }
            A[t] = v;
        }
        if (tid == 0) sflag[1] = 0;
        __syncthreads();
        // ---- Gauss-Jordan with partial pivoting on the R x R block, carrying the N rhs columns ----
        for (int k = 0; k < R; ++k) {
            if (warp == 0) {
                double best = -1.0;
                int bi = k;
                for (int i = k + lane; i < R; i += 32) {
                    const double2 x = A[(size_t)i * W + k];
                    const double v = hypot(x.x, x.y);
                    if (v > best) { best = v; bi = i; }
                }
                for (int o = 16; o; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) {
                    sflag[0] = bi;
                    if ((!(best > 0.0) || !(best < CUDART_INF)) && sflag[1] == 0) sflag[1] = k + 1;
                }
            }
            __syncthreads();
            const int p = sflag[0];
            if (p != k)
                for (int c2 = k + tid; c2 < W; c2 += blockDim.x) {
                    const double2 t2 = A[(size_t)k * W + c2];
                    A[(size_t)k * W + c2] = A[(size_t)p * W + c2];
                    A[(size_t)p * W + c2] = t2;
                }
            __syncthreads();
            const double2 pinv = crecip(A[(size_t)k * W + k]);
            __syncthreads();
            for (int c2 = k + tid; c2 < W; c2 += blockDim.x) A[(size_t)k * W + c2] = cmul(A[(size_t)k * W + c2], pinv);
            __syncthreads();
            // eliminate column k from every other row (columns > k only; column k is never read again)
            const int wc = W - (k + 1);
            for (int t = tid; t < R * wc; t += blockDim.x) {
                const int i = t / wc, c2 = k + 1 + (t - i * wc);
                if (i == k) continue;
                const double2 f = A[(size_t)i * W + k];
                A[(size_t)i * W + c2] = csub(A[(size_t)i * W + c2], cmul(f, A[(size_t)k * W + c2]));
            }
            __syncthreads();
        }
        // X = columns R.. of the reduced system: X[p][f];  Y_N_c[f][p] = X[p][f], I_N_c[f] = X[N][f]
        for (int t = tid; t < N * N; t += blockDim.x) {
            const int f = t / N, p = t - f * N;
            a.Y_N_c[(size_t)d * N * N + t] = A[(size_t)p * W + R + f];
        }
        for (int f = tid; f < N; f += blockDim.x) a.I_N_c[(size_t)d * N + f] = A[(size_t)N * W + R + f];
        if (tid == 0) a.info[d] = sflag[1];
    }
}
