// B200 fp64 pipe characterisation: dependent-DFMA latency and throughput vs ILP and warps/SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64mb scripts/fp64_microbench.cu && /tmp/fp64mb
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_sm, double* d_out) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int sms = p.multiProcessorCount;
    int threads = 32 * warps_per_sm;     // one CTA per SM
    if (threads > 1024) return;
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP><<<sms, threads>>>(d_out, 1.0000001, 1e-9, 10);
    cudaEventRecord(e0);
    chain<ILP><<<sms, threads>>>(d_out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst_per_warp = (double)iters * 16 * ILP;
    double clk = 1.965e9 * ms * 1e-3;
    double thread_fma_per_clk_sm = inst_per_warp * warps_per_sm * 32 / clk;
    printf("ILP %d warps/SM %2d: %.3f ms  DFMA lanes/clk/SM %.1f  cycles per dependent step (1 warp view) %.2f\n",
           ILP, warps_per_sm, ms, thread_fma_per_clk_sm, clk / ((double)iters * 16));
}

int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {1, 2, 4, 8, 16, 32}) {
        run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d);
    }
    cudaFree(d);
    return 0;
}
