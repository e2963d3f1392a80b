// diagnostics.cu — device-side post-processing of a finished solve:
//   * Milne photosphere pressure per wavelength and the flux-weighted sums behind
//     effective_temperature_milne()            (frei/core.py:386-405)
//   * the bolometric trapezoid integral behind effective_temperature_planck() (frei/core.py:408-414)
//   * the normalised contribution function of the dashboard (frei/plot.py:63-79)
// The reference walks the wavelengths in a Python loop (one np.interp call each); here one thread
// owns one wavelength column of dtaus [L][n_lam] and the spectrum / dtaus never leave HBM.
// HBM-bound: 8 L bytes read per wavelength for T_eff, 16 L read + 8 L written with the
// contribution function (the column is walked twice: sum, then normalised values).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kDiagThreads = 256;
constexpr int kLikelyInCache = 8;          // numpy's LIKELY_IN_CACHE_SIZE

struct DiagArgs {
    const double* dtaus;    // [L][n]
    const double* spec;     // [n]  F_up of the top level
    const double* lam_um;   // [n]  this device's wavelength slice
    const double* w;        // [n]  trapezoid weights of the global grid (cm)
    const double* P;        // [L]  bar, bottom -> top
    const double* T;        // [L]
    double* p_milne;        // [n] or null
    double* cf;             // [L][n] or null
    double* block_sums;     // [blocks][3]
    int L;
    int64_t n;
};

// xp[l] of np.interp(2/3, np.exp(-dtaus[:, j]), P): evaluated where the search looks
__device__ __forceinline__ double xp_at(const DiagArgs& a, int64_t j, int l) {
    return exp(-a.dtaus[(int64_t)l * a.n + j]);
}

// numpy's binary_search_with_guess (numpy/_core/src/multiarray/compiled_base.c) with guess = 0,
// restated because the reference calls np.interp on a sequence that is NOT sorted in general
// (row 0 of dtaus is the row of ones, frei/twostream.py:352): the result is then defined by
// the search path, not by the mathematical interpolant.  Returns -1 (below xp[0]), len (above
// xp[len-1]) or the index of the bracket's left node.
__device__ __forceinline__ int numpy_search(const DiagArgs& a, int64_t j, double key) {
    const int len = a.L;
    if (key > xp_at(a, j, len - 1)) return len;
    if (key < xp_at(a, j, 0)) return -1;
    if (len <= 4) {
        int i = 1;
        while (i < len && key >= xp_at(a, j, i)) ++i;
        return i - 1;
    }
    int guess = 0, imin = 0, imax = len;
    if (guess > len - 3) guess = len - 3;
    if (guess < 1) guess = 1;
    if (key < xp_at(a, j, guess)) {
        if (key < xp_at(a, j, guess - 1)) {
            imax = guess - 1;
            if (guess > kLikelyInCache && key >= xp_at(a, j, guess - kLikelyInCache))
                imin = guess - kLikelyInCache;
        } else {
            return guess - 1;
        }
    } else {
        if (key < xp_at(a, j, guess + 1)) return guess;
        if (key < xp_at(a, j, guess + 2)) return guess + 1;
        imin = guess + 2;
        if (guess < len - kLikelyInCache - 1 && key < xp_at(a, j, guess + kLikelyInCache))
            imax = guess + kLikelyInCache;
    }
    while (imin < imax) {
        const int imid = imin + ((imax - imin) >> 1);
        if (key >= xp_at(a, j, imid)) imin = imid + 1; else imax = imid;
    }
    return imin - 1;
}

// np.interp(key, xp, P) for one column (arr_interp of the same file, scalar x)
__device__ __forceinline__ double numpy_interp_column(const DiagArgs& a, int64_t j, double key) {
    const int len = a.L;
    const int k = numpy_search(a, j, key);
    if (k == -1) return a.P[0];
    if (k == len) return a.P[len - 1];
    if (k == len - 1) return a.P[k];
    const double x0 = xp_at(a, j, k);
    if (x0 == key) return a.P[k];
    const double x1 = xp_at(a, j, k + 1);
    const double slope = (a.P[k + 1] - a.P[k]) / (x1 - x0);
    double r = slope * (key - x0) + a.P[k];
    if (isnan(r)) {
        r = slope * (key - x1) + a.P[k + 1];
        if (isnan(r) && a.P[k] == a.P[k + 1]) r = a.P[k];
    }
    return r;
}

// fixed-order block sum of three values: shuffle butterfly inside the warps, then the warps in
// warp order -> deterministic
__device__ __forceinline__ void block_sum3(double v[3], double* sm /* [3][warps] */, double out[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double x = v[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[c * nw + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < nw; ++k) t += sm[threadIdx.x * nw + k];
        out[threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kDiagThreads) diag_kernel(DiagArgs a) {
    __shared__ double sm[3 * (kDiagThreads / 32)];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    if (j < a.n) {
        const double F = a.spec[j];
        const double lam_cm = a.lam_um[j] * 1e-4;
        const double pm = numpy_interp_column(a, j, 2.0 / 3.0);              // core.py:392-395
        if (a.p_milne) a.p_milne[j] = pm;
        const double wt = F * lam_cm;                // erg s^-1 cm^-2 via spectral_density, core.py:399-400
        v[0] = pm * wt;
        v[1] = wt;
        v[2] = a.w[j] * F;                           // np.trapz(spec.flux, lam), core.py:413
        if (a.cf) {
            // plot.py:63-79 with r = L-1-l walking top -> bottom:
            //   cf_r = exp(-tau_r) dtau_r (P_r / dP_r) nu^3 / expm1(h c nu / (k_B T_r)),
            //   tau = cumsum(dtaus[::-1]),  dP = (1 - 10^-dlogP) P,  normalised over the column.
            // Stored in level order (the reference plots cf[::-1], plot.py:83).
            const int L = a.L;
            const double nu = 1.0 / lam_cm;
            const double nu3 = nu * nu * nu;
            const double hcperk = FREI_H * FREI_C / FREI_KB;
            double pmax = a.P[0], pmin = a.P[0];
            for (int l = 1; l < L; ++l) { pmax = fmax(pmax, a.P[l]); pmin = fmin(pmin, a.P[l]); }
            const double dlogP = (log10(pmax) - log10(pmin)) / (L - 1);
            const double omk = 1.0 - pow(10.0, -dlogP);
            double tau = 0.0, tot = 0.0;
            for (int l = L - 1; l >= 0; --l) {
                const double dt = a.dtaus[(int64_t)l * a.n + j];
                tau += dt;
                const double Pl = a.P[l];
                tot += exp(-tau) * dt * (Pl / (omk * Pl)) * nu3 / expm1(hcperk * nu / a.T[l]);
            }
            tau = 0.0;
            for (int l = L - 1; l >= 0; --l) {
                const double dt = a.dtaus[(int64_t)l * a.n + j];
                tau += dt;
                const double Pl = a.P[l];
                const double c = exp(-tau) * dt * (Pl / (omk * Pl)) * nu3 / expm1(hcperk * nu / a.T[l]);
                a.cf[(int64_t)l * a.n + j] = c / tot;
            }
        }
    }
    double out[3];
    block_sum3(v, sm, out);
    if (threadIdx.x < 3) a.block_sums[(int64_t)blockIdx.x * 3 + threadIdx.x] = out[threadIdx.x];
}

// block sums -> three totals, in block order (one CTA; thread c strides are summed in a fixed
// tree: 256 interleaved partial sums per component, then sequentially)
__global__ void __launch_bounds__(kDiagThreads) diag_finish_kernel(const double* __restrict__ block_sums,
                                                                    int64_t blocks, double* __restrict__ sums) {
    __shared__ double sm[3][kDiagThreads];
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t k = threadIdx.x; k < blocks; k += kDiagThreads) {
        v[0] += block_sums[k * 3 + 0];
        v[1] += block_sums[k * 3 + 1];
        v[2] += block_sums[k * 3 + 2];
    }
    sm[0][threadIdx.x] = v[0]; sm[1][threadIdx.x] = v[1]; sm[2][threadIdx.x] = v[2];
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < kDiagThreads; ++k) t += sm[threadIdx.x][k];
        sums[threadIdx.x] = t;
    }
}

}  // namespace

extern "C" {

int64_t frei_b200_diagnostics_scratch_bytes(int64_t n_lam) {
    if (n_lam <= 0) return 0;
    return ((n_lam + kDiagThreads - 1) / kDiagThreads) * 3 * (int64_t)sizeof(double);
}

int frei_b200_diagnostics(const double* d_dtaus, const double* d_spec, const double* d_lam_um,
                          const double* d_w, const double* d_P_bar, const double* d_T, int32_t L,
                          int64_t n_lam, double* d_pressure_milne, double* d_cf, double* d_scratch,
                          double* d_sums, void* stream) {
    if (!d_dtaus || !d_spec || !d_lam_um || !d_w || !d_P_bar || !d_T || !d_scratch || !d_sums)
        return frei_set_err(FREI_E_ARG, "bad argument: null pointer (frei_b200_diagnostics)");
    if (L < 2 || n_lam <= 0)
        return frei_set_err(FREI_E_ARG, "bad argument: L >= 2 && n_lam > 0 (frei_b200_diagnostics)");
    DiagArgs a;
    a.dtaus = d_dtaus; a.spec = d_spec; a.lam_um = d_lam_um; a.w = d_w; a.P = d_P_bar; a.T = d_T;
    a.p_milne = d_pressure_milne; a.cf = d_cf; a.block_sums = d_scratch; a.L = L; a.n = n_lam;
    const int64_t blocks = (n_lam + kDiagThreads - 1) / kDiagThreads;
    if (blocks > 0x7fffffff)
        return frei_set_err(FREI_E_UNSUPPORTED, "too many wavelengths for one launch (frei_b200_diagnostics)");
    cudaStream_t st = (cudaStream_t)stream;
    diag_kernel<<<(unsigned)blocks, kDiagThreads, 0, st>>>(a);
    diag_finish_kernel<<<1, kDiagThreads, 0, st>>>(d_scratch, blocks, d_sums);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return frei_set_err(FREI_E_CUDA, cudaGetErrorString(e));
    return FREI_OK;
}

}  // extern "C"
