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

extern "C" int frei_b200_bin_trapz(const void* d_a, int32_t dtype, int64_t n_rows, int64_t n_samples,
                                   int64_t row_stride, const int64_t* d_run_start,
                                   const int64_t* d_run_end, const int32_t* d_bin_first_run,
                                   int32_t n_bins, double* d_out, void* stream) {
    if (!d_a || !d_run_start || !d_run_end || !d_bin_first_run || !d_out || n_rows <= 0 ||
        n_rows > 65535 || n_samples <= 0 || n_bins <= 0 || (dtype != FREI_F32 && dtype != FREI_F64))
        return frei_set_err(FREI_E_ARG, "bad argument to frei_b200_bin_trapz");
    const double per_bin = (double)n_samples / n_bins;
    const int G = per_bin >= 64 ? 32 : per_bin >= 24 ? 16 : per_bin >= 10 ? 8 : 4;
    const int rc = (dtype == FREI_F32)
        ? launch_bins<float>((const float*)d_a, n_rows, n_samples, row_stride, d_run_start, d_run_end,
                             d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream)
        : launch_bins<double>((const double*)d_a, n_rows, n_samples, row_stride, d_run_start, d_run_end,
                              d_bin_first_run, n_bins, d_out, G, (cudaStream_t)stream);
    return rc ? frei_set_err(FREI_E_CUDA, "bin_trapz launch failed") : FREI_OK;
}
