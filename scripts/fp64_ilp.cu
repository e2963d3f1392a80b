// B200 fp64 pipe: DFMA throughput of W warps per SM with ILP independent dependent chains per thread
// (the sweep kernel runs 16 warps per SM with two wavelengths = two chains per thread).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ilp scripts/fp64_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int REGS3>      // REGS3 = 1: three distinct register operands
__global__ void chain(double* out, double a, double b, int iters) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = 1.0 + threadIdx.x * 1e-12 + i * 1e-13; z[i] = 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = REGS3 ? fma(x[i], y[i], z[i]) : fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, int REGS3>
void run(int warps, double* d) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP, REGS3><<<148, 32 * warps>>>(d, 1.0000001, 1e-9, 10);
    cudaEventRecord(e0);
    chain<ILP, REGS3><<<148, 32 * warps>>>(d, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = 1.965e9 * ms * 1e-3;
    printf("warps/SM %2d ILP %d regs3 %d: DFMA lanes/clk/SM %.1f (peak 64)\n", warps, ILP, REGS3,
           (double)iters * 16 * ILP * warps * 32 / clk);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 12, 16}) {
        run<1, 0>(w, d); run<2, 0>(w, d); run<4, 0>(w, d);
        run<1, 1>(w, d); run<2, 1>(w, d); run<4, 1>(w, d);
    }
    return 0;
}
