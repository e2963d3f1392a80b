// peaks.cu — measured ceiling of the resource that bounds the fp64 sweep: the fp64 pipe.
// MEASURED_PEAKS.json (driver-written) holds an HBM copy bandwidth and a bf16 tensor rate but no
// fp64 figure, so bench.py measures one in the same run, on the same clocks, with this kernel:
// every thread runs ILP independent chains of dependent DFMAs (register operands only), one CTA
// of 512 threads per SM x 4 (= 16 warps per scheduler: the pipe, not latency, is the limit).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kIlp = 8, kInner = 16;

__global__ void __launch_bounds__(512) dfma_peak_kernel(double* out, double a, double b, int iters) {
    double x[kIlp];
#pragma unroll
    for (int i = 0; i < kIlp; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < kInner; ++r) {
#pragma unroll
            for (int i = 0; i < kIlp; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) s += x[i];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int frei_b200_fp64_peak(double* d_scratch, int64_t scratch_doubles, double* h_dfma_per_s,
                                   void* stream) {
    int dev = 0, sms = 0;
    if (!d_scratch || !h_dfma_per_s) return frei_set_err(FREI_E_ARG, "bad argument: null pointer");
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        return frei_set_err(FREI_E_CUDA, "cannot query the device");
    const int blocks = sms * 4, threads = 512, iters = 2000;
    if (scratch_doubles < (int64_t)blocks * threads)
        return frei_set_err(FREI_E_ARG, "bad argument: scratch too small (need 2048 doubles per SM)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
        return frei_set_err(FREI_E_CUDA, "cudaEventCreate failed");
    double best = 0.0;
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(d_scratch, 1.0000001, 1e-9, 50);      // warm-up
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        dfma_peak_kernel<<<blocks, threads, 0, st>>>(d_scratch, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)blocks * threads * iters * kInner * kIlp;               // thread-level DFMAs
        if (ms > 0.f && n / (ms * 1e-3) > best) best = n / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess || best <= 0.0)
        return frei_set_err(FREI_E_CUDA, "fp64 peak kernel failed");
    *h_dfma_per_s = best;
    return FREI_OK;
}
