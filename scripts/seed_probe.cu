// accuracy of the fp64 hardware seeds (MUFU.RCP64H / RSQ64H) and of 1 vs 2 Newton steps
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* x, double* o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double v = x[i], y, z;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(v));
    o[i] = y; o[n + i] = z;
    double e = fma(-v, y, 1.0); double y1 = fma(y, e, y); o[2 * n + i] = y1;
    e = fma(-v, y1, 1.0); o[3 * n + i] = fma(y1, e, y1);
    double h = 0.5 * v; double e2 = fma(-h * z, z, 0.5); double z1 = fma(z, e2, z); o[4 * n + i] = z1;
    e2 = fma(-h * z1, z1, 0.5); o[5 * n + i] = fma(z1, e2, z1);
}
int main() {
    const int n = 1 << 20; double *hx = new double[n], *ho = new double[6 * n], *dx, *dout;
    for (int i = 0; i < n; ++i) hx[i] = pow(10.0, -8.0 + 16.0 * (i + 0.5) / n) * (1.0 + 0.37 * ((i * 2654435761u) % 1000) / 1000.0);
    cudaMalloc(&dx, n * 8); cudaMalloc(&dout, 6 * n * 8); cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dout, n); cudaMemcpy(ho, dout, 6 * n * 8, cudaMemcpyDeviceToHost);
    double m[6] = {0};
    for (int i = 0; i < n; ++i) {
        double r = 1.0 / hx[i], s = 1.0 / sqrt(hx[i]);
        double ref[6] = {r, s, r, r, s, s};
        for (int c = 0; c < 6; ++c) m[c] = fmax(m[c], fabs(ho[c * n + i] - ref[c]) / ref[c]);
    }
    printf("rcp seed %.3e  rsqrt seed %.3e | rcp 1 step %.3e 2 steps %.3e | rsqrt 1 step %.3e 2 steps %.3e\n", m[0], m[1], m[2], m[3], m[4], m[5]);
    return 0;
}
