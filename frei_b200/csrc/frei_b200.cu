// frei_b200.cu — sm_100a kernels and the C ABI (include/frei_b200.h) of the
// radiative-equilibrium hot path: (P,T) bracket + weights (K0), opacity gather
// (K1), two-stream layer response (K2), the fused layer sweep with
// wavelength-integral partials (K2+K3), the fixed-order reduction and the
// per-layer temperature update (K4).
//
// Reference behaviour restated here (paths into the reference checkout):
//   frei/opacity.py:173-269   kappa(), Rayleigh
//   frei/twostream.py:16-287  bolometric_flux, BB, E, propagate_fluxes, layer thermodynamics
//   frei/twostream.py:351-416 emit loop body,  :486-545 absorb loop body
// No tensor cores: nothing here is a dense contraction.  The sweep is a
// one-thread-per-wavelength serial recurrence over layers held in registers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/frei_b200.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess)                                                     \
            return set_err(FREI_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define ARG_TRY(cond)                                                               \
    do {                                                                            \
        if (!(cond)) return set_err(FREI_E_ARG, "bad argument: %s%s", #cond);       \
    } while (0)

// ---------------------------------------------------------------------------
// constants (CGS, CODATA 2018 as shipped by astropy >= 4.0)
// ---------------------------------------------------------------------------
#define FREI_KB      1.380649e-16
#define FREI_MP      1.67262192369e-24
#define FREI_H       6.62607015e-27
#define FREI_C       2.99792458e10
#define FREI_SIGSB   5.6703744191844314e-5
#define FREI_BAR     1e6
#define FREI_PI      3.141592653589793

constexpr int kThreads = 128;            // threads per sweep CTA (4 warps)
constexpr int kWarps = kThreads / 32;
constexpr int kMaxS = 32;

// ---------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------
struct LayerParams {            // views into ws->layer_params
    double*  dpg;               // [B][L]      (p1 - p2) / g
    double*  invT;              // [B][L]      1 / T_i
    double*  W;                 // [B][L][S][4] mmr-premultiplied corner weights
    int32_t* base;              // [B][L][S]   table row of corner (iP, iT)
};

static inline int64_t round16(int64_t x) { return (x + 15) & ~int64_t(15); }

static inline int64_t layer_params_bytes(int B, int L, int S) {
    return round16((int64_t)B * L * (16 + 36 * (int64_t)S));
}
static inline LayerParams layer_params_view(void* p, int B, int L, int S) {
    LayerParams v;
    char* c = (char*)p;
    int64_t n = (int64_t)B * L;
    v.dpg = (double*)c;
    v.invT = (double*)(c + n * 8);
    v.W = (double*)(c + n * 16);
    v.base = (int32_t*)(c + n * 16 + n * S * 32);
    return v;
}
static inline int64_t sweep_blocks(int64_t n_lam) { return (n_lam + kThreads - 1) / kThreads; }

// ---------------------------------------------------------------------------
// K0: brackets, weights, per-layer scalars
// ---------------------------------------------------------------------------
// scipy find_indices rule: below grid -> 0; >= last node -> n-2; else x[i] <= v < x[i+1].
__device__ __forceinline__ int bracket_index(const double* __restrict__ x, int n, double v) {
    if (v < x[0]) return 0;
    if (v >= x[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;              // invariant x[lo] <= v < x[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

struct PrepArgs {
    const double* axis_P; const double* axis_T; const int32_t* has_T;
    const double* T; const double* P; const double* mmr; const double* g;
    LayerParams lp;
    int32_t* iP; int32_t* iT; double* wP; double* wT; uint8_t* oob;
    int B, L, S, N_P, N_T;
};

__device__ __forceinline__ void prep_one(const PrepArgs& a, int b, int i) {
    const int L = a.L, S = a.S;
    const double* T = a.T + (int64_t)b * L;
    const double* P = a.P + (int64_t)b * L;
    const double g = a.g[b];
    const double p1 = P[i] * FREI_BAR;
    double p2;
    if (i == L - 1) p2 = p1 * (P[L - 2] * FREI_BAR) / (P[L - 3] * FREI_BAR);   // twostream.py:359
    else p2 = P[i + 1] * FREI_BAR;
    const int64_t li = (int64_t)b * L + i;
    a.lp.dpg[li] = (p1 - p2) / g;                                              // twostream.py:231
    a.lp.invT[li] = 1.0 / T[i];
    for (int s = 0; s < S; ++s) {
        const double* xp = a.axis_P + (int64_t)s * a.N_P;
        const double* xt = a.axis_T + (int64_t)s * a.N_T;
        const double vp = P[i], vt = T[i];
        int ip = bracket_index(xp, a.N_P, vp);
        double wp = (vp - xp[ip]) / (xp[ip + 1] - xp[ip]);
        bool out = (vp < xp[0]) || (vp > xp[a.N_P - 1]);
        int it = 0; double wt = 0.0;
        if (a.has_T[s]) {
            it = bracket_index(xt, a.N_T, vt);
            wt = (vt - xt[it]) / (xt[it + 1] - xt[it]);
            out = out || (vt < xt[0]) || (vt > xt[a.N_T - 1]);
        }
        const double m = a.mmr[li * S + s];
        double w00 = (1.0 - wp) * (1.0 - wt), w01 = (1.0 - wp) * wt;
        double w10 = wp * (1.0 - wt), w11 = wp * wt;
        if (out) { w00 = w01 = w10 = w11 = 0.0; }       // fill_value=0, opacity.py:243
        double* W = a.lp.W + (li * S + s) * 4;
        W[0] = m * w00; W[1] = m * w01; W[2] = m * w10; W[3] = m * w11;
        a.lp.base[li * S + s] = (s * a.N_P + ip) * a.N_T + it;
        if (a.iP) a.iP[li * S + s] = ip;
        if (a.iT) a.iT[li * S + s] = it;
        if (a.wP) a.wP[li * S + s] = wp;
        if (a.wT) a.wT[li * S + s] = wt;
        if (a.oob) a.oob[li * S + s] = out ? 1 : 0;
    }
}

__global__ void prep_kernel(PrepArgs a) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.L) return;
    prep_one(a, idx / a.L, idx % a.L);
}

// ---------------------------------------------------------------------------
// per-wavelength constants: Planck prefactors, Rayleigh sigma, trapezoid weights, F_TOA
// ---------------------------------------------------------------------------
__global__ void spectral_kernel(const double* __restrict__ lam_um, int64_t n_global, int64_t off,
                                int64_t n_local, double m_bar, double T_star, double a_rstar, double f,
                                double* __restrict__ c1, double* __restrict__ c2,
                                double* __restrict__ sigma, double* __restrict__ w,
                                double* __restrict__ f_toa) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_local) return;
    const int64_t jg = off + j;
    const double lu = lam_um[jg], l = lu * 1e-4;
    const double k1 = 2.0 * FREI_H * FREI_C * FREI_C / pow(l, 5.0);
    c1[j] = k1;
    c2[j] = FREI_H * FREI_C / (l * FREI_KB);
    // Rayleigh, opacity.py:173-200
    const double nh2 = 13.58e-5 * (1.0 + 7.52e-11 / (l * l)) + 1.0;
    const double nhe = 1e-8 * (2283.0 + (1.8102e13 / (1.5342e10 - 1.0 / (lu * lu)))) + 1.0;
    const double pi3 = FREI_PI * FREI_PI * FREI_PI, l4 = (l * l) * (l * l);
    const double rh = (nh2 * nh2 - 1.0) / (nh2 * nh2 + 2.0), re = (nhe * nhe - 1.0) / (nhe * nhe + 2.0);
    const double s_h2 = (24.0 * pi3 / (2.68678e19 * 2.68678e19) / l4 * (rh * rh)) / m_bar;
    const double s_he = (24.0 * pi3 / (2.546899e19 * 2.546899e19) / l4 * (re * re)) / m_bar;
    sigma[j] = s_h2 + s_he;
    // trapezoid weights of the global grid (in cm)
    double wj;
    if (n_global == 1) wj = 0.0;
    else if (jg == 0) wj = 0.5 * (lam_um[1] * 1e-4 - l);
    else if (jg == n_global - 1) wj = 0.5 * (l - lam_um[jg - 1] * 1e-4);
    else wj = 0.5 * (lam_um[jg + 1] * 1e-4 - lam_um[jg - 1] * 1e-4);
    w[j] = wj;
    // F_TOA, core.py:48-62
    const double Bs = k1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T_star));
    f_toa[j] = f * (1.0 / (a_rstar * a_rstar)) * 1.0 / (2.0 * FREI_PI) * (FREI_PI * Bs);
}

// ---------------------------------------------------------------------------
// two-stream layer response, g_0 = 0   (twostream.py:139-176)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double planck(double c1, double c2, double invT) {
    return c1 / expm1(c2 * invT);                        // twostream.py:64-67
}

__device__ __forceinline__ void two_stream(double dtau, double w0, double F1u, double F2d,
                                           double B1, double B2, double& F2u, double& F1d) {
    const double Ew = (w0 > 0.1) ? (1.225 - 0.1777 * w0 - 0.05582 * (w0 * w0)) : 1.0;   // :89-94
    const double EmW = Ew - w0;
    const double Tr = exp(-2.0 * sqrt(Ew * EmW) * dtau);                                // :139
    const double r = sqrt(EmW / Ew);
    const double zp = 0.5 * (1.0 + r), zm = 0.5 * (1.0 - r);                            // :143-146
    const double Tr2 = Tr * Tr;
    const double chi = zm * zm * Tr2 - zp * zp;                                         // :149
    const double xi = zp * zm * (1.0 - Tr2);                                            // :150
    const double psi = (zm * zm - zp * zp) * Tr;                                        // :151
    const double pit = FREI_PI * (1.0 - w0) / EmW;                                      // :152
    const double q = ((B1 - B2) / dtau) / (2.0 * Ew);                                   // :158, :165
    const double inv_chi = 1.0 / chi;
    F2u = inv_chi * (psi * F1u - xi * F2d +
                     pit * (B2 * (chi + xi) - psi * B1 + q * (chi - psi - xi)));        // :161-168
    F1d = inv_chi * (psi * F2d - xi * F1u +
                     pit * (B1 * (chi + xi) - psi * B2 + q * (xi + psi - chi)));        // :169-176
}

// ---------------------------------------------------------------------------
// K1 standalone: k and sigma for every level
// ---------------------------------------------------------------------------
template <typename TabT>
__global__ void kappa_kernel(const TabT* __restrict__ tab, const double* __restrict__ sigma,
                             const double* __restrict__ sigma_scale, LayerParams lp,
                             double* __restrict__ k_out, double* __restrict__ sigma_out,
                             int L, int S, int N_T, int64_t n_lam) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= n_lam) return;
    const int64_t li = (int64_t)b * L + i;
    const double sg = sigma[j] * (sigma_scale ? sigma_scale[b] : 1.0);
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const double* W = lp.W + (li * S + s) * 4;
        const TabT* r0 = tab + (int64_t)lp.base[li * S + s] * n_lam + j;
        double v = 0.0;
        v += (double)r0[0] * W[0];
        v += (double)r0[n_lam] * W[1];
        v += (double)r0[(int64_t)N_T * n_lam] * W[2];
        v += (double)r0[(int64_t)(N_T + 1) * n_lam] * W[3];
        acc += v;
    }
    k_out[li * n_lam + j] = acc + sg;                    // opacity.py:269 (k includes sigma)
    if (i == 0) sigma_out[(int64_t)b * n_lam + j] = sg;
}

// ---------------------------------------------------------------------------
// K2 standalone: propagate_fluxes elementwise
// ---------------------------------------------------------------------------
__global__ void propagate_kernel(const double* __restrict__ lam, const double* __restrict__ F1u,
                                 const double* __restrict__ F2d, double T1, double T2,
                                 const double* __restrict__ dtau, const double* __restrict__ w0,
                                 double* __restrict__ F2u, double* __restrict__ F1d, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double l = lam[j];
    const double c1 = 2.0 * FREI_H * FREI_C * FREI_C / pow(l, 5.0);
    const double B1 = c1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T1));
    const double B2 = c1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T2));
    double a, d;
    two_stream(dtau[j], w0[j], F1u[j], F2d[j], B1, B2, a, d);
    F2u[j] = a; F1d[j] = d;
}

// ---------------------------------------------------------------------------
// K2+K3: the layer sweep
// ---------------------------------------------------------------------------
struct SweepArgs {
    const void* tab;
    const double* c1; const double* c2; const double* sigma; const double* w; const double* f_toa;
    const double* sigma_scale; const double* ftoa_scale;
    LayerParams lp;
    void* F_up; void* F_down; void* dtaus;
    double* partials;           // [B][nblk][L][4]
    int64_t n_lam;
    int B, L, S, N_T;
};

// Sum four per-lane values across the warp; on return lanes 0, 8, 16, 24 hold the
// totals of v0, v1, v2, v3 respectively.  Fixed butterfly -> deterministic.
__device__ __forceinline__ double warp_reduce4(double v0, double v1, double v2, double v3, int lane) {
    const unsigned full = 0xffffffffu;
    // step 1 (xor 16): lower half keeps (v0, v1), upper half keeps (v2, v3)
    const bool up16 = lane & 16;
    double s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3;     // what I send
    double k0 = up16 ? v2 : v0, k1 = up16 ? v3 : v1;     // what I keep
    k0 += __shfl_xor_sync(full, s0, 16);
    k1 += __shfl_xor_sync(full, s1, 16);
    // step 2 (xor 8): within each half, lanes with bit 3 clear keep k0, set keep k1
    const bool up8 = lane & 8;
    double s = up8 ? k0 : k1, k = up8 ? k1 : k0;
    k += __shfl_xor_sync(full, s, 8);
    k += __shfl_xor_sync(full, k, 4);
    k += __shfl_xor_sync(full, k, 2);
    k += __shfl_xor_sync(full, k, 1);
    return k;     // lane 0: v0, lane 8: v1, lane 16: v2, lane 24: v3
}

template <typename TabT, int S_T>
__device__ __forceinline__ double gather_k(const TabT* __restrict__ tab, const LayerParams& lp,
                                           int64_t li, int S, int N_T, int64_t n_lam, int64_t j) {
    const int SS = (S_T > 0) ? S_T : S;
    double acc = 0.0;
    if (S_T > 0) {
        TabT t[S_T > 0 ? S_T : 1][4];
#pragma unroll
        for (int s = 0; s < SS; ++s) {
            const TabT* r0 = tab + (int64_t)__ldg(lp.base + li * SS + s) * n_lam + j;
            t[s][0] = __ldg(r0);
            t[s][1] = __ldg(r0 + n_lam);
            t[s][2] = __ldg(r0 + (int64_t)N_T * n_lam);
            t[s][3] = __ldg(r0 + (int64_t)(N_T + 1) * n_lam);
        }
#pragma unroll
        for (int s = 0; s < SS; ++s) {
            const double4* Wp = reinterpret_cast<const double4*>(lp.W + (li * SS + s) * 4);
            const double2 wa = __ldg(reinterpret_cast<const double2*>(Wp));
            const double2 wb = __ldg(reinterpret_cast<const double2*>(Wp) + 1);
            double v = (double)t[s][0] * wa.x;
            v = fma((double)t[s][1], wa.y, v);
            v = fma((double)t[s][2], wb.x, v);
            v = fma((double)t[s][3], wb.y, v);
            acc += v;
        }
    } else {
        for (int s = 0; s < SS; ++s) {
            const TabT* r0 = tab + (int64_t)__ldg(lp.base + li * SS + s) * n_lam + j;
            const double* W = lp.W + (li * SS + s) * 4;
            double v = (double)__ldg(r0) * __ldg(W);
            v = fma((double)__ldg(r0 + n_lam), __ldg(W + 1), v);
            v = fma((double)__ldg(r0 + (int64_t)N_T * n_lam), __ldg(W + 2), v);
            v = fma((double)__ldg(r0 + (int64_t)(N_T + 1) * n_lam), __ldg(W + 3), v);
            acc += v;
        }
    }
    return acc;
}

template <typename TabT, int S_T, int DIR>
__global__ void __launch_bounds__(kThreads) sweep_kernel(SweepArgs a) {
    extern __shared__ double sm_part[];          // [L][kWarps][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int L = a.L, S = a.S;
    const int64_t n_lam = a.n_lam;
    const int64_t j_raw = (int64_t)blockIdx.x * kThreads + tid;
    const bool live = j_raw < n_lam;
    const int64_t j = live ? j_raw : n_lam - 1;

    const TabT* tab = static_cast<const TabT*>(a.tab);
    double* Fu = static_cast<double*>(a.F_up) + (int64_t)b * L * n_lam + j;
    double* Fd = static_cast<double*>(a.F_down) + (int64_t)b * L * n_lam + j;
    double* dt_out = a.dtaus ? static_cast<double*>(a.dtaus) + (int64_t)b * L * n_lam + j : nullptr;

    const double c1 = a.c1[j], c2 = a.c2[j];
    const double sg = a.sigma[j] * (a.sigma_scale ? a.sigma_scale[b] : 1.0);
    const double wj = live ? a.w[j] : 0.0;
    const int64_t lb = (int64_t)b * L;

    if (dt_out && live) dt_out[0] = 1.0;         // leading row of ones, twostream.py:352/:487

    if (DIR == FREI_EMIT) {
        const double ftoa = a.f_toa[j] * (a.ftoa_scale ? a.ftoa_scale[b] : 1.0);
        double F1u = Fu[n_lam];                                  // fluxes_up[1], stale
        double B1 = planck(c1, c2, __ldg(a.lp.invT + lb + 1));
        if (warp == 0 && lane < 4) sm_part[(0 * kWarps + 0) * 4 + lane] = 0.0;
        for (int i = 1; i < L; ++i) {
            const bool top = (i == L - 1);
            const double F2d = top ? ftoa : Fd[(int64_t)(i + 1) * n_lam];     // :379-382
            const double k = gather_k<TabT, S_T>(tab, a.lp, lb + i, S, a.N_T, n_lam, j) + sg;
            const double B2 = top ? B1 : planck(c1, c2, __ldg(a.lp.invT + lb + i + 1));
            const double dtau = __ldg(a.lp.dpg + lb + i) * k;                 // :371-373
            const double w0 = sg / (sg + k);                                  // :376-378
            double F2u, F1d;
            two_stream(dtau, w0, F1u, F2d, B1, B2, F2u, F1d);
            if (live) {
                if (!top) Fu[(int64_t)(i + 1) * n_lam] = F2u;                 // :392-394
                Fd[(int64_t)i * n_lam] = F1d;
                if (dt_out) dt_out[(int64_t)i * n_lam] = dtau;
            }
            const double red = warp_reduce4(wj * F2u, wj * F2d, wj * F1u, wj * F1d, lane);
            if ((lane & 7) == 0) sm_part[(i * kWarps + warp) * 4 + (lane >> 3)] = red;
            F1u = F2u; B1 = B2;
        }
    } else {
        double F2d = Fd[(int64_t)(L - 1) * n_lam];                            // fluxes_down[L-1]
        double B2 = planck(c1, c2, __ldg(a.lp.invT + lb + L - 1));
        for (int i = L - 2; i >= 0; --i) {
            const double F1u = Fu[(int64_t)i * n_lam];                        // stale, :512
            const double k = gather_k<TabT, S_T>(tab, a.lp, lb + i, S, a.N_T, n_lam, j) + sg;
            const double B1 = planck(c1, c2, __ldg(a.lp.invT + lb + i));
            const double dtau = __ldg(a.lp.dpg + lb + i) * k;
            const double w0 = sg / (sg + k);
            double F2u, F1d;
            two_stream(dtau, w0, F1u, F2d, B1, B2, F2u, F1d);
            if (live) {
                Fu[(int64_t)(i + 1) * n_lam] = F2u;                           // :521-522
                Fd[(int64_t)i * n_lam] = F1d;
                if (dt_out) dt_out[(int64_t)(L - 1 - i) * n_lam] = dtau;      // visiting order
            }
            const double red = warp_reduce4(wj * F2u, wj * F2d, wj * F1u, wj * F1d, lane);
            if ((lane & 7) == 0) sm_part[(i * kWarps + warp) * 4 + (lane >> 3)] = red;
            F2d = F1d; B2 = B1;
        }
    }
    __syncthreads();
    // combine the warps of this CTA in fixed order and publish [L][4]
    double* out = a.partials + ((int64_t)b * gridDim.x + blockIdx.x) * L * 4;
    const int i_lo = (DIR == FREI_EMIT) ? 1 : 0, i_hi = (DIR == FREI_EMIT) ? L : L - 1;
    for (int e = tid; e < L * 4; e += kThreads) {
        const int i = e >> 2, cidx = e & 3;
        double s = 0.0;
        if (i >= i_lo && i < i_hi) {
#pragma unroll
            for (int wq = 0; wq < kWarps; ++wq) s += sm_part[(i * kWarps + wq) * 4 + cidx];
        }
        out[e] = s;
    }
}

// ---------------------------------------------------------------------------
// fixed-order reduction of block partials: one warp per (b, layer, component)
// ---------------------------------------------------------------------------
__global__ void reduce_kernel(const double* __restrict__ partials, double* __restrict__ sums,
                              int B, int L, int nblk) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= B * L * 4) return;
    const int b = gw / (L * 4), e = gw % (L * 4);
    const double* p = partials + (int64_t)b * nblk * L * 4 + e;
    double s = 0.0;
    for (int k = lane; k < nblk; k += 32) s += p[(int64_t)k * L * 4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sums[gw] = s;
}

// ---------------------------------------------------------------------------
// K4: per-layer thermodynamics and the temperature update
// ---------------------------------------------------------------------------
__device__ __forceinline__ double cp_of(double m_bar) { return (2.0 + 5.0) / (2.0 * m_bar) * FREI_KB; }   // :220-224
__device__ __forceinline__ double dz_of(double T, double p1, double p2, double g, double m_bar) {
    return (FREI_KB * T) / (m_bar * g) * log(p1 / p2);                                                    // :186-187
}

// One CTA per atmosphere, thread i = layer i: every read of T precedes the barrier, every
// write follows it.
__global__ void update_T_kernel(double* __restrict__ T, const double* __restrict__ P,
                                const double* __restrict__ g_arr, const double* __restrict__ mbar_arr,
                                const double* __restrict__ alpha_arr, const double* __restrict__ sums,
                                double* __restrict__ dT_out, double* __restrict__ T_hist,
                                int L, int direction, double alpha_override) {
    const int b = blockIdx.x, i = threadIdx.x;
    double* Tb = T + (int64_t)b * L;
    const double* Pb = P + (int64_t)b * L;
    double dT = 0.0, T1 = 0.0;
    if (i < L) {
        T1 = Tb[i];
        const bool active = (direction == FREI_EMIT) ? (i >= 1) : (i <= L - 2);
        if (active) {
            const double g = g_arr[b], m_bar = mbar_arr[b];
            const double alpha = (alpha_override >= 0.0) ? alpha_override : alpha_arr[b];
            const double p1 = Pb[i] * FREI_BAR;
            double p2, T2;
            if (i == L - 1) { p2 = p1 * (Pb[L - 2] * FREI_BAR) / (Pb[L - 3] * FREI_BAR); T2 = T1; }   // :358-363
            else { p2 = Pb[i + 1] * FREI_BAR; T2 = Tb[i + 1]; }
            const double* s = sums + ((int64_t)b * L + i) * 4;
            const double dF_rad = (s[0] - s[1]) - (s[2] - s[3]);                      // :199
            const double cp = cp_of(m_bar);
            const double dz = dz_of(T1, p1, p2, g, m_bar);
            const double rho = ((p1 - p2) / g) / dz;                                  // :238
            const double dgam = (T1 - T2) / dz - g / cp;                              // :241-266
            const double lmix = alpha * FREI_KB * T1 / (m_bar * g);                   // :270
            double F_conv = 0.0;
            if (dgam > 0.0) F_conv = rho * cp * (lmix * lmix) * sqrt(g / T1) * pow(dgam, 1.5);   // :285-287
            const double div = (dF_rad + F_conv) / dz;                                // :205
            const double X = div * dz;
            const double f_pre = (X != 0.0) ? 1e5 / pow(fabs(X), 0.9) : 1.0;          // :32-35
            const double dt_rad = cp * p1 / FREI_SIGSB / g / (T1 * T1 * T1);          // :37
            double dt = f_pre * dt_rad;
            if (dgam > 0.0) dt = f_pre * fmin(dt_rad, sqrt(T1 / g / dgam));           // :39-43
            // delta_temperature is called without m_bar: defaults 2.4 m_p, n_dof 5 (:403-405)
            const double m_def = 2.4 * FREI_MP;
            const double rho_def = ((p1 - p2) / g) / dz_of(T1, p1, p2, g, m_def);
            dT = 1.0 / rho_def / cp_of(m_def) * div * dt;                             // :216-217
        }
    }
    __syncthreads();
    if (i < L) {
        const double Tn = T1 - dT;                                                    // :407, :536
        dT_out[(int64_t)b * L + i] = dT;
        Tb[i] = Tn;
        if (T_hist) T_hist[(int64_t)b * L + i] = Tn;
    }
}

// ---------------------------------------------------------------------------
// host side of the ABI
// ---------------------------------------------------------------------------
template <typename TabT, int S_T>
static int launch_sweep_dir(const SweepArgs& a, int direction, dim3 grid, size_t smem, cudaStream_t st) {
    if (direction == FREI_EMIT) {
        CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<TabT, S_T, FREI_EMIT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep_kernel<TabT, S_T, FREI_EMIT><<<grid, kThreads, smem, st>>>(a);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<TabT, S_T, FREI_ABSORB>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep_kernel<TabT, S_T, FREI_ABSORB><<<grid, kThreads, smem, st>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

template <typename TabT>
static int launch_sweep(const SweepArgs& a, int direction, dim3 grid, size_t smem, cudaStream_t st) {
    switch (a.S) {
        case 1: return launch_sweep_dir<TabT, 1>(a, direction, grid, smem, st);
        case 2: return launch_sweep_dir<TabT, 2>(a, direction, grid, smem, st);
        case 3: return launch_sweep_dir<TabT, 3>(a, direction, grid, smem, st);
        case 4: return launch_sweep_dir<TabT, 4>(a, direction, grid, smem, st);
        case 8: return launch_sweep_dir<TabT, 8>(a, direction, grid, smem, st);
        default: return launch_sweep_dir<TabT, 0>(a, direction, grid, smem, st);
    }
}

extern "C" {

const char* frei_b200_last_error(void) { return g_err; }
int frei_b200_abi_version(void) { return FREI_B200_ABI_VERSION; }

int frei_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int frei_b200_workspace_bytes(int32_t B, int32_t L, int32_t S, int64_t n_lam,
                              int64_t* layer_params, int64_t* partials, int64_t* sums, int64_t* dT) {
    ARG_TRY(B > 0 && L >= 3 && S > 0 && S <= kMaxS && n_lam > 0);
    if (layer_params) *layer_params = layer_params_bytes(B, L, S);
    if (partials) *partials = (int64_t)B * sweep_blocks(n_lam) * L * 4 * 8;
    if (sums) *sums = (int64_t)B * L * 4 * 8;
    if (dT) *dT = (int64_t)B * L * 8;
    return FREI_OK;
}

static int check_common(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws) {
    ARG_TRY(tab && atm && ws);
    ARG_TRY(tab->values && tab->axis_P && tab->axis_T && tab->has_T);
    ARG_TRY(tab->S > 0 && tab->S <= kMaxS && tab->N_P >= 2 && tab->N_T >= 2 && tab->n_lam > 0);
    ARG_TRY(tab->dtype == FREI_F32 || tab->dtype == FREI_F64);
    ARG_TRY(atm->T && atm->P && atm->mmr && atm->g && atm->m_bar && atm->alpha);
    ARG_TRY(atm->B > 0 && atm->L >= 3);
    ARG_TRY(ws->layer_params);
    return FREI_OK;
}

int frei_b200_spectral_setup(const double* d_lam_um, int64_t n_global, int64_t offset, int64_t n_local,
                             double m_bar, double T_star, double a_rstar, double f,
                             double* d_c1, double* d_c2, double* d_sigma, double* d_w, double* d_f_toa,
                             void* stream) {
    ARG_TRY(d_lam_um && d_c1 && d_c2 && d_sigma && d_w && d_f_toa);
    ARG_TRY(n_global > 0 && n_local > 0 && offset >= 0 && offset + n_local <= n_global);
    ARG_TRY(m_bar > 0 && T_star > 0 && a_rstar > 0);
    spectral_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_lam_um, n_global, offset, n_local, m_bar, T_star, a_rstar, f,
        d_c1, d_c2, d_sigma, d_w, d_f_toa);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_layer_prep(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                         int32_t* d_iP, int32_t* d_iT, double* d_wP, double* d_wT, uint8_t* d_oob,
                         void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    PrepArgs a;
    a.axis_P = tab->axis_P; a.axis_T = tab->axis_T; a.has_T = tab->has_T;
    a.T = atm->T; a.P = atm->P; a.mmr = atm->mmr; a.g = atm->g;
    a.lp = layer_params_view(ws->layer_params, atm->B, atm->L, tab->S);
    a.iP = d_iP; a.iT = d_iT; a.wP = d_wP; a.wT = d_wT; a.oob = d_oob;
    a.B = atm->B; a.L = atm->L; a.S = tab->S; a.N_P = tab->N_P; a.N_T = tab->N_T;
    const int n = atm->B * atm->L;
    prep_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_kappa(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                    const frei_workspace* ws, double* d_k, double* d_sigma, void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    ARG_TRY(spec && spec->sigma && d_k && d_sigma && spec->n_lam == tab->n_lam);
    ARG_TRY(atm->L <= 65535 && atm->B <= 65535);
    LayerParams lp = layer_params_view(ws->layer_params, atm->B, atm->L, tab->S);
    dim3 grid((unsigned)((tab->n_lam + 255) / 256), atm->L, atm->B);
    if (tab->dtype == FREI_F32)
        kappa_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float*)tab->values, spec->sigma, atm->sigma_scale, lp, d_k, d_sigma,
            atm->L, tab->S, tab->N_T, tab->n_lam);
    else
        kappa_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const double*)tab->values, spec->sigma, atm->sigma_scale, lp, d_k, d_sigma,
            atm->L, tab->S, tab->N_T, tab->n_lam);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_propagate(const double* d_lam_cm, const double* d_F1_up, const double* d_F2_down,
                        double T1, double T2, const double* d_delta_tau, const double* d_omega0,
                        double* d_F2_up, double* d_F1_down, int64_t n, void* stream) {
    ARG_TRY(d_lam_cm && d_F1_up && d_F2_down && d_delta_tau && d_omega0 && d_F2_up && d_F1_down);
    ARG_TRY(n > 0);
    propagate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_lam_cm, d_F1_up, d_F2_down, T1, T2, d_delta_tau, d_omega0, d_F2_up, d_F1_down, n);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_sweep(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                    const frei_flux* flux, int32_t direction, const frei_workspace* ws, void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    ARG_TRY(spec && spec->c1 && spec->c2 && spec->sigma && spec->w && spec->f_toa);
    ARG_TRY(spec->n_lam == tab->n_lam);
    ARG_TRY(flux && flux->F_up && flux->F_down && ws->partials);
    ARG_TRY(direction == FREI_EMIT || direction == FREI_ABSORB);
    if (flux->dtype != FREI_F64)
        return set_err(FREI_E_UNSUPPORTED, "flux dtype %s not supported%s", "f32");
    ARG_TRY(atm->B <= 65535);
    SweepArgs a;
    a.tab = tab->values;
    a.c1 = spec->c1; a.c2 = spec->c2; a.sigma = spec->sigma; a.w = spec->w; a.f_toa = spec->f_toa;
    a.sigma_scale = atm->sigma_scale; a.ftoa_scale = atm->ftoa_scale;
    a.lp = layer_params_view(ws->layer_params, atm->B, atm->L, tab->S);
    a.F_up = flux->F_up; a.F_down = flux->F_down; a.dtaus = flux->dtaus;
    a.partials = ws->partials;
    a.n_lam = tab->n_lam; a.B = atm->B; a.L = atm->L; a.S = tab->S; a.N_T = tab->N_T;
    dim3 grid((unsigned)sweep_blocks(tab->n_lam), atm->B);
    const size_t smem = (size_t)atm->L * kWarps * 4 * sizeof(double);
    if (smem > 200 * 1024) return set_err(FREI_E_UNSUPPORTED, "too many layers for shared memory%s%s");
    if (tab->dtype == FREI_F32) return launch_sweep<float>(a, direction, grid, smem, (cudaStream_t)stream);
    return launch_sweep<double>(a, direction, grid, smem, (cudaStream_t)stream);
}

int frei_b200_reduce(const frei_atmosphere* atm, const frei_workspace* ws, int64_t n_lam, void* stream) {
    ARG_TRY(atm && ws && ws->partials && ws->sums && n_lam > 0);
    const int nw = atm->B * atm->L * 4;
    reduce_kernel<<<(nw * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        ws->partials, ws->sums, atm->B, atm->L, (int)sweep_blocks(n_lam));
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_update_T(const frei_atmosphere* atm, const frei_workspace* ws, int32_t direction,
                       double alpha_override, double* d_T_hist, void* stream) {
    ARG_TRY(atm && ws && ws->sums && ws->dT && atm->T && atm->P && atm->g && atm->m_bar && atm->alpha);
    ARG_TRY(atm->L >= 3 && atm->L <= 1024);
    ARG_TRY(direction == FREI_EMIT || direction == FREI_ABSORB);
    const int threads = ((atm->L + 31) / 32) * 32;
    update_T_kernel<<<atm->B, threads, 0, (cudaStream_t)stream>>>(
        atm->T, atm->P, atm->g, atm->m_bar, atm->alpha, ws->sums, ws->dT, d_T_hist,
        atm->L, direction, alpha_override);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_sweep_step(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                         const frei_flux* flux, int32_t direction, double alpha_override,
                         const frei_workspace* ws, double* d_T_hist, void* stream) {
    int rc = frei_b200_layer_prep(tab, atm, ws, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    rc = frei_b200_sweep(tab, spec, atm, flux, direction, ws, stream);
    if (rc) return rc;
    rc = frei_b200_reduce(atm, ws, tab->n_lam, stream);
    if (rc) return rc;
    return frei_b200_update_T(atm, ws, direction, alpha_override, d_T_hist, stream);
}

}  // extern "C"
