// B200 fp64 pipe: does DFMA throughput depend on how many distinct register operands it reads?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64op scripts/fp64_operands.cu && /tmp/fp64op
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x = fma(x, a, b)   a, b kernel constants (uniform)        -> 1 register operand
// MODE 1: x = fma(x, y, b)   y per-thread register                   -> 2 register operands
// MODE 2: x = fma(x, y, z)   y, z per-thread registers               -> 3 register operands
// MODE 3: x = x * y (DMUL, 2 regs)     MODE 4: x = x + y (DADD, 2 regs)
template <int MODE, int ILP>
__global__ void chain(double* out, double a, double b, int iters) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = threadIdx.x * 1e-9 + i;
        y[i] = 1.0 + threadIdx.x * 1e-12 + i * 1e-13;
        z[i] = 1e-9 * (threadIdx.x + i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                if (MODE == 1) x[i] = fma(x[i], y[i], b);
                if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);
                if (MODE == 3) x[i] = x[i] * y[i];
                if (MODE == 4) x[i] = x[i] + z[i];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(int warps, double* d) {
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<MODE, ILP><<<148, 32 * warps>>>(d, 1.0000001, 1e-9, 10);
    cudaEventRecord(e0);
    chain<MODE, ILP><<<148, 32 * warps>>>(d, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clk = 1.965e9 * ms * 1e-3;
    printf("mode %d ILP %d warps/SM %2d: lanes/clk/SM %.1f\n", MODE, ILP, warps,
           (double)iters * 16 * ILP * warps * 32 / clk);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {8, 16, 32}) {
        run<0, 4>(w, d); run<1, 4>(w, d); run<2, 4>(w, d); run<3, 4>(w, d); run<4, 4>(w, d);
    }
    return 0;
}
