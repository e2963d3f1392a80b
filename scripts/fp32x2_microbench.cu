// B200: throughput of packed FFMA2 against scalar FFMA (fp32 lanes per clock per SM) and how much
// issue room each leaves for other work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 scripts/fp32x2_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>      // 0: scalar FFMA, 1: FFMA2, 2: FFMA2 + one IADD per FFMA2, 3: scalar FFMA + one IADD per 2 FFMA
__global__ void chain(float* out, float a, float b, int iters) {
    float2 x[ILP];
    unsigned c = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f + i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0 || MODE == 3) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
                else x[i] = __ffma2_rn(x[i], a2, b2);
                if (MODE >= 2) c = c * 3u + (unsigned)r;
            }
        }
    }
    float s = (float)c;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(int warps, float* d) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<MODE, ILP><<<148, 32 * warps>>>(d, 1.0000001f, 1e-9f, 10);
    cudaEventRecord(e0);
    chain<MODE, ILP><<<148, 32 * warps>>>(d, 1.0000001f, 1e-9f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = 1.965e9 * ms * 1e-3;
    printf("mode %d ILP %d warps/SM %2d: fp32 FMA lanes/clk/SM %.1f\n", MODE, ILP, warps,
           (double)iters * 8 * ILP * 2 * warps * 32 / clk);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
    for (int w : {8, 16, 32}) { run<0, 8>(w, d); run<1, 8>(w, d); run<2, 8>(w, d); run<3, 8>(w, d); }
    return 0;
}
