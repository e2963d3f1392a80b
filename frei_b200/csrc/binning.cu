// binning.cu — load-time wavelength binning of line-by-line opacities.
//
// Restates frei/interp.py:156-202 (AggregateTrapz._loop, Trapz._inner) as used by
// groupby_bins_agg (frei/interp.py:270-307) from binned_opacity (frei/opacity.py:137-139):
// for every leading index (temperature, pressure) and every wavelength bin b,
//     out[row][b] = sum over consecutive samples i, i+1 that both fall in bin b of
//                   (a[row][i] + a[row][i+1]) / 2            (unit sample spacing, x = None)
// The host turns the bin code of every sample (pandas.cut, right-closed bins) into runs of
// equal consecutive codes; a run [s, e) contributes sum_{i=s}^{e-2} (a[i] + a[i+1]) / 2.
// HBM-bound segmented reduction: G lanes per run read the run's samples contiguously.
//
// With sample positions x (the groupies=False branch of binned_opacity, frei/opacity.py:29-40,
// 150-162: xarray's integrate('wavelength') per bin) the trapezoids carry their widths,
//     out[row][b] = sum (a[i] + a[i+1]) / 2 * (x[i+1] - x[i]).
//
// regrid_kernel finishes the load: nearest-neighbour lookup of the Grid's
// (temperature, pressure) nodes in the source grid (frei/opacity.py:141-146, 31-33) and, for the
// groupies=False branch, linear interpolation with extrapolation along wavelength onto the bin
// centres (frei/opacity.py:163-166), in scipy.interpolate.interp1d's arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"



template <typename T, int G>
__global__ void bin_trapz_kernel(const T* __restrict__ a, int64_t n_samples, int64_t row_stride,
                                 const int64_t* __restrict__ run_start, const int64_t* __restrict__ run_end,
                                 const int32_t* __restrict__ bin_first_run, int32_t n_bins,
                                 double* __restrict__ out) {
    // one group of G lanes per (row, bin); a bin owns the runs [bin_first_run[b], bin_first_run[b+1])
    const int lane = threadIdx.x & 31, sub = lane % G;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int row = blockIdx.y;
    const bool valid = grp < n_bins;
    const int b = valid ? (int)grp : n_bins - 1;
    const T* ar = a + (int64_t)row * row_stride;
    double acc = 0.0;
    for (int r = bin_first_run[b]; r < bin_first_run[b + 1]; ++r) {
        const int64_t s = run_start[r], e = run_end[r];
        // sum_{i=s}^{e-2} (a[i] + a[i+1]) / 2 = sum_{i=s}^{e-1} a[i] - (a[s] + a[e-1]) / 2:
        // every sample is read once, lanes stride over i
        for (int64_t i = s + sub; i < e; i += G) acc += (double)ar[i];
        if (sub == 0) acc -= 0.5 * ((double)ar[s] + (double)ar[e - 1]);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
    if (valid && sub == 0) out[(int64_t)row * n_bins + b] = acc;
}

template <typename T, int G>
__global__ void bin_trapz_x_kernel(const T* __restrict__ a, const double* __restrict__ x, int64_t row_stride,
                                   const int64_t* __restrict__ run_start, const int64_t* __restrict__ run_end,
                                   const int32_t* __restrict__ bin_first_run, int32_t n_bins,
                                   double* __restrict__ out) {
    const int lane = threadIdx.x & 31, sub = lane % G;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int row = blockIdx.y;
    const bool valid = grp < n_bins;
    const int b = valid ? (int)grp : n_bins - 1;
    const T* ar = a + (int64_t)row * row_stride;
    double acc = 0.0;
    for (int r = bin_first_run[b]; r < bin_first_run[b + 1]; ++r) {
        const int64_t s = run_start[r], e = run_end[r];
        for (int64_t i = s + sub; i + 1 < e; i += G)
            acc += ((double)ar[i] + (double)ar[i + 1]) * 0.5 * (x[i + 1] - x[i]);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
    if (valid && sub == 0) out[(int64_t)row * n_bins + b] = acc;
}

// out[it][ip][j] for the Grid's nodes (it, ip): source row (src_T[it], src_P[ip]) of
// binned [nT][nP][nb]; j0 == null: element j (copy), else scipy's linear form between the source
// columns j0[j] and j0[j] + 1: slope = (y_hi - y_lo) / dx[j], y = slope * t[j] + y_lo.
__global__ void regrid_kernel(const double* __restrict__ binned, int nP, int64_t nb,
                              const int32_t* __restrict__ src_T, const int32_t* __restrict__ src_P,
                              int mP, const int32_t* __restrict__ j0, const double* __restrict__ dx,
                              const double* __restrict__ t, int64_t m, double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int ip = blockIdx.y, it = blockIdx.z;
    const double* src = binned + ((int64_t)src_T[it] * nP + src_P[ip]) * nb;
    double v;
    if (j0) {
        const double lo = src[j0[j]], hi = src[j0[j] + 1];
        v = (hi - lo) / dx[j] * t[j] + lo;
    } else {
        v = src[j];
    }
    out[((int64_t)it * mP + ip) * m + j] = v;
}

template <typename T>
static int launch_bins_x(const T* a, const double* x, int64_t n_rows, int64_t row_stride,
                         const int64_t* rs, const int64_t* re, const int32_t* bfr, int32_t n_bins,
                         double* out, int G, cudaStream_t st) {
    const int threads = 256;
    auto grid = [&](int g) { return dim3((unsigned)(((int64_t)n_bins * g + threads - 1) / threads), (unsigned)n_rows); };
    switch (G) {
        case 4: bin_trapz_x_kernel<T, 4><<<grid(4), threads, 0, st>>>(a, x, row_stride, rs, re, bfr, n_bins, out); break;
        case 8: bin_trapz_x_kernel<T, 8><<<grid(8), threads, 0, st>>>(a, x, row_stride, rs, re, bfr, n_bins, out); break;
        case 16: bin_trapz_x_kernel<T, 16><<<grid(16), threads, 0, st>>>(a, x, row_stride, rs, re, bfr, n_bins, out); break;
        default: bin_trapz_x_kernel<T, 32><<<grid(32), threads, 0, st>>>(a, x, row_stride, rs, re, bfr, n_bins, out); break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <typename T>
static int launch_bins(const T* a, int64_t n_rows, int64_t n_samples, int64_t row_stride,
                       const int64_t* rs, const int64_t* re, const int32_t* bfr, int32_t n_bins,
                       double* out, int G, cudaStream_t st) {
    const int threads = 256;
    auto grid = [&](int g) { return dim3((unsigned)(((int64_t)n_bins * g + threads - 1) / threads), (unsigned)n_rows); };
    switch (G) {
        case 4: bin_trapz_kernel<T, 4><<<grid(4), threads, 0, st>>>(a, n_samples, row_stride, rs, re, bfr, n_bins, out); break;
        case 8: bin_trapz_kernel<T, 8><<<grid(8), threads, 0, st>>>(a, n_samples, row_stride, rs, re, bfr, n_bins, out); break;
        case 16: bin_trapz_kernel<T, 16><<<grid(16), threads, 0, st>>>(a, n_samples, row_stride, rs, re, bfr, n_bins, out); break;
        default: bin_trapz_kernel<T, 32><<<grid(32), threads, 0, st>>>(a, n_samples, row_stride, rs, re, bfr, n_bins, out); break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int frei_b200_regrid(const double* d_binned, int32_t nT, int32_t nP, int64_t nb,
                                const int32_t* d_src_T, int32_t mT, const int32_t* d_src_P, int32_t mP,
                                const int32_t* d_j0, const double* d_dx, const double* d_t, int64_t m,
                                double* d_out, void* stream) {
    if (!d_binned || !d_src_T || !d_src_P || !d_out || nT <= 0 || nP <= 0 || nb <= 0 || mT <= 0 ||
        mP <= 0 || m <= 0 || mT > 65535 || mP > 65535 || (d_j0 && (!d_dx || !d_t)) || (!d_j0 && m != nb))
        return frei_set_err(FREI_E_ARG, "bad argument to frei_b200_regrid");
    regrid_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)mP, (unsigned)mT), 256, 0, (cudaStream_t)stream>>>(
        d_binned, nP, nb, d_src_T, d_src_P, mP, d_j0, d_dx, d_t, m, d_out);
    return cudaGetLastError() == cudaSuccess ? FREI_OK : frei_set_err(FREI_E_CUDA, "regrid launch failed");
}

extern "C" int frei_b200_bin_trapz(const void* d_a, int32_t dtype, const double* d_x, int64_t n_rows,
                                   int64_t n_samples, int64_t row_stride, const int64_t* d_run_start,
                                   const int64_t* d_run_end, const int32_t* d_bin_first_run,
                                   int32_t n_bins, double* d_out, void* stream) {
    if (!d_a || !d_run_start || !d_run_end || !d_bin_first_run || !d_out || n_rows <= 0 ||
        n_rows > 65535 || n_samples <= 0 || n_bins <= 0 || (dtype != FREI_F32 && dtype != FREI_F64))
        return frei_set_err(FREI_E_ARG, "bad argument to frei_b200_bin_trapz");
    const double per_bin = (double)n_samples / n_bins;
    const int G = per_bin >= 64 ? 32 : per_bin >= 24 ? 16 : per_bin >= 10 ? 8 : 4;
    if (d_x) {
        const int rcx = (dtype == FREI_F32)
            ? launch_bins_x<float>((const float*)d_a, d_x, n_rows, row_stride, d_run_start, d_run_end,
                                   d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream)
            : launch_bins_x<double>((const double*)d_a, d_x, n_rows, row_stride, d_run_start, d_run_end,
                                    d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream);
        return rcx ? frei_set_err(FREI_E_CUDA, "bin_trapz launch failed") : FREI_OK;
    }
    const int rc = (dtype == FREI_F32)
        ? launch_bins<float>((const float*)d_a, n_rows, n_samples, row_stride, d_run_start, d_run_end,
                             d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream)
        : launch_bins<double>((const double*)d_a, n_rows, n_samples, row_stride, d_run_start, d_run_end,
                              d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream);
    return rc ? frei_set_err(FREI_E_CUDA, "bin_trapz launch failed") : FREI_OK;
}
