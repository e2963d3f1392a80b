// B200: does non-fp64 work (integer, select, shared-memory loads) issued between DFMAs cost fp64
// throughput?  8 independent DFMA chains per thread + NI filler instructions per 8 DFMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64mix scripts/fp64_mix.cu && /tmp/fp64mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int KIND>   // KIND 0: IMAD, 1: FSEL-like selects, 2: LDS.64
__global__ void mix(double* out, double a, double b, int iters, int m) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double x[8];
    unsigned y[16];
    float z[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { y[i] = threadIdx.x + i; z[i] = i; }
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (KIND == 0) y[i] = y[i] * m + 1u;
                if (KIND == 1) z[i] = (y[i] > (unsigned)it) ? z[i] : z[(i + 1) % 16] + 0.0f * 0 ;
                if (KIND == 2) acc += sm[(threadIdx.x + y[i] + it + r) & 1023];
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NI, int KIND>
void run(int warps_per_sm, double* d_out) {
    int threads = 32 * warps_per_sm, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NI, KIND><<<148, threads>>>(d_out, 1.0000001, 1e-9, 10, 3);
    cudaEventRecord(e0);
    mix<NI, KIND><<<148, threads>>>(d_out, 1.0000001, 1e-9, iters, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clk = 1.965e9 * ms * 1e-3;
    double dfma_warp_instr_per_smsp = (double)iters * 64 * warps_per_sm / 4;
    printf("kind %d filler %2d per 8 DFMA, warps/SM %2d: cycles per DFMA warp-instr per SMSP %.2f\n",
           KIND, NI, warps_per_sm, clk / dfma_warp_instr_per_smsp);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {4, 16}) {
        run<0, 0>(w, d); run<4, 0>(w, d); run<8, 0>(w, d); run<16, 0>(w, d);
        run<8, 1>(w, d); run<16, 1>(w, d);
        run<2, 2>(w, d); run<4, 2>(w, d);
    }
    return 0;
}
