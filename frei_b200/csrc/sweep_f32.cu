// sweep_f32.cu — the layer sweep in fp32 arithmetic (flux state and opacity table in fp32,
// wavelength integrals accumulated in fp64).  BASELINE.json allows 1e-4 relative per-wavelength
// flux error in fp32; the kernel is the same algorithm as sweep_kernel in frei_b200.cu
// (frei/twostream.py:351-405, 486-533) with
//   * hardware approximations for 1/x, 1/sqrt(x) and 2^x (MUFU.RCP / RSQ / EX2, ~2^-22),
//   * the cancellation-free grouping of the two-stream expressions (mandatory in fp32: the
//     reference's grouping loses all digits for delta_tau < 1e-3 in single precision),
//   * short series for 1 - exp(-u) and 1 - (1 - exp(-u))/u below u = 1/8,
//   * 4 wavelengths per thread (16-byte loads/stores, 16-byte cp.async of table rows).
// Algorithmic traffic 4 S * 4 + 3 * 4 = 60 B per evaluation for S = 3: HBM-bound.
#include "common.cuh"

namespace {

__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqf(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// T = exp(-u), m = 1 - exp(-u), e1 = 1 - m/u  (u >= 0)
__device__ __forceinline__ void exp_neg_f(float u, float& T, float& m, float& e1) {
    T = ex2f(-1.4426950408889634f * u);
    const bool small = u < 0.125f;
    const float h = 1.0f - u * (1.0f / 3.0f) * (1.0f - u * 0.25f * (1.0f - u * 0.2f * (1.0f - u * (1.0f / 6.0f))));
    const float e1s = 0.5f * u * h;                      // u/2 - u^2/6 + u^3/24 - u^4/120 + u^5/720
    const float ms = u * (1.0f - e1s);                   // m = u (1 - e1)
    m = small ? ms : 1.0f - T;
    e1 = small ? e1s : 1.0f - m * rcpf(u);
}

__device__ __forceinline__ float planck_f(float c1, float c2, float invT) {
    float e, m, e1;
    exp_neg_f(c2 * invT, e, m, e1);                      // c1 / expm1(x) = c1 e^-x / (1 - e^-x)
    return c1 * e * rcpf(m);                             // twostream.py:64-67
}

// Same regrouping as two_stream_k (frei_b200.cu) in single precision.
__device__ __forceinline__ void two_stream_f(float k, float sg, float dpg, float F1u, float F2d,
                                             float B1, float B2, float& F2u, float& F1d, float& dtau) {
    dtau = dpg * k;                                                      // :371-373
    const float R1 = rcpf(sg + k);
    const float w0 = sg * R1, omw = k * R1;                              // omega0, 1 - omega0
    const bool hi = w0 > 0.1f;                                           // :89-94
    const float Ep = 1.225f - 0.1777f * w0 - 0.05582f * (w0 * w0);
    const float Ew = hi ? Ep : 1.0f;
    const float invE = hi ? rcpf(Ep) : 1.0f;
    const float EmW = Ew - w0;
    const float q = Ew * EmW;
    const float a = q * rsqf(q);
    const float r = a * invE;                                            // :143
    const float u = 2.0f * a * dtau;                                     // T = exp(-u), :139
    float Tr, m, e1;
    exp_neg_f(u, Tr, m, e1);
    const float z = 0.5f * (w0 * invE) * rcpf(1.0f + r);                 // zeta_minus, :145
    const float zm = z * m, omzm = 1.0f - zm;
    const float chi = -(r + zm) * omzm;                                  // :149
    const float xi = (1.0f - z) * zm * (2.0f - m);                       // :150
    const float psi = -r * Tr;                                           // :151
    const float A = 2.0f * xi - m * omzm;                                // chi + xi - psi
    const float H = (B1 - B2) * r * (zm * (1.0f - e1) + (e1 - m));       // psi D + B'/(2E)(chi-psi-xi)
    const float R3 = rcpf(EmW * chi);
    const float ic = EmW * R3;                                           // 1 / chi
    const float pc = (3.14159265358979f * omw) * R3;                     // :152
    F2u = ic * (psi * F1u - xi * F2d) + pc * (B2 * A + H);               // :161-168
    F1d = ic * (psi * F2d - xi * F1u) + pc * (B1 * A - H);               // :169-176
}

template <typename T, int V>
struct alignas((sizeof(T) * V) > 16 ? 16 : sizeof(T) * V) Pack { T v[V]; };

template <int V, typename T>
__device__ __forceinline__ void ldv(const T* p, float* o) {
    const Pack<T, V> t = *reinterpret_cast<const Pack<T, V>*>(p);
#pragma unroll
    for (int v = 0; v < V; ++v) o[v] = (float)t.v[v];
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float* x) {
    Pack<float, V> t;
#pragma unroll
    for (int v = 0; v < V; ++v) t.v[v] = x[v];
    *reinterpret_cast<Pack<float, V>*>(p) = t;
}

template <int BYTES>
__device__ __forceinline__ void cp_async_n(uint32_t dst, const void* src) {
    if (BYTES == 32) { cp_async<16>(dst, src); cp_async<16>(dst + 16, (const char*)src + 16); }
    else cp_async<BYTES>(dst, src);
}

template <int V, int THREADS>
struct LaneF { float c1[V], c2[V], sg[V], wj[V], Fcar[V], Bcar[V]; };

template <typename TabT, int V, int THREADS>
__device__ __forceinline__ void stage_rows_f(const TabT* __restrict__ tabj, const double* rec, int S,
                                             int64_t n_lam, int64_t rowT, uint32_t stage) {
    constexpr int kSlot = V * (int)sizeof(TabT);
    constexpr uint32_t kRow = (uint32_t)THREADS * kSlot;
    const int64_t* off = reinterpret_cast<const int64_t*>(rec) + 2 + 4 * S;
    for (int s = 0; s < S; ++s) {
        const TabT* r0 = tabj + off[s];
        cp_async_n<kSlot>(stage + (4 * s + 0) * kRow, r0);
        cp_async_n<kSlot>(stage + (4 * s + 1) * kRow, r0 + n_lam);
        cp_async_n<kSlot>(stage + (4 * s + 2) * kRow, r0 + rowT);
        cp_async_n<kSlot>(stage + (4 * s + 3) * kRow, r0 + rowT + n_lam);
    }
    cp_async_commit();
}

template <typename TabT, int V, int THREADS>
__device__ __forceinline__ void gather_f(const TabT* slot, const double* rec, int S, const float* sg, float* k) {
    constexpr int kRowElems = THREADS * V;
#pragma unroll
    for (int v = 0; v < V; ++v) k[v] = 0.0f;
    for (int s = 0; s < S; ++s) {
        const double2 wa = *reinterpret_cast<const double2*>(rec + 2 + 4 * s);
        const double2 wb = *reinterpret_cast<const double2*>(rec + 4 + 4 * s);
        const float w0 = (float)wa.x, w1 = (float)wa.y, w2 = (float)wb.x, w3 = (float)wb.y;
        float t0[V], t1[V], t2[V], t3[V];
        ldv<V>(slot + (4 * s + 0) * kRowElems, t0);
        ldv<V>(slot + (4 * s + 1) * kRowElems, t1);
        ldv<V>(slot + (4 * s + 2) * kRowElems, t2);
        ldv<V>(slot + (4 * s + 3) * kRowElems, t3);
#pragma unroll
        for (int v = 0; v < V; ++v)
            k[v] += fmaf(t3[v], w3, fmaf(t2[v], w2, fmaf(t1[v], w1, t0[v] * w0)));
    }
#pragma unroll
    for (int v = 0; v < V; ++v) k[v] += sg[v];            // k includes sigma, opacity.py:269
}

// fixed butterfly sum of four values over the warp (see warp_reduce4 in frei_b200.cu)
__device__ __forceinline__ double reduce4(double v0, double v1, double v2, double v3, int lane) {
    const unsigned full = 0xffffffffu;
    const bool up16 = lane & 16;
    double s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3, k0 = up16 ? v2 : v0, k1 = up16 ? v3 : v1;
    k0 += __shfl_xor_sync(full, s0, 16);
    k1 += __shfl_xor_sync(full, s1, 16);
    const bool up8 = lane & 8;
    double s = up8 ? k0 : k1, k = up8 ? k1 : k0;
    k += __shfl_xor_sync(full, s, 8);
    k += __shfl_xor_sync(full, k, 4);
    k += __shfl_xor_sync(full, k, 2);
    k += __shfl_xor_sync(full, k, 1);
    return k;
}

template <int DIR, int V, int THREADS, bool SAME_T>
__device__ __forceinline__ void step_f(LaneF<V, THREADS>& t, const float* k, float dpg, const float* other,
                                       float invTn, float* F2u, float* F1d, float* dtau, double* red) {
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float Bn = SAME_T ? t.Bcar[v] : planck_f(t.c1[v], t.c2[v], invTn);
        if (DIR == FREI_EMIT) {
            two_stream_f(k[v], t.sg[v], dpg, t.Fcar[v], other[v], t.Bcar[v], Bn, F2u[v], F1d[v], dtau[v]);
            r0 = fmaf(t.wj[v], F2u[v], r0); r1 = fmaf(t.wj[v], other[v], r1);
            r2 = fmaf(t.wj[v], t.Fcar[v], r2); r3 = fmaf(t.wj[v], F1d[v], r3);
            t.Fcar[v] = F2u[v];
        } else {
            two_stream_f(k[v], t.sg[v], dpg, other[v], t.Fcar[v], Bn, t.Bcar[v], F2u[v], F1d[v], dtau[v]);
            r0 = fmaf(t.wj[v], F2u[v], r0); r1 = fmaf(t.wj[v], t.Fcar[v], r1);
            r2 = fmaf(t.wj[v], other[v], r2); r3 = fmaf(t.wj[v], F1d[v], r3);
            t.Fcar[v] = F1d[v];
        }
        t.Bcar[v] = Bn;
    }
    red[0] = r0; red[1] = r1; red[2] = r2; red[3] = r3;
}

// One CTA = THREADS * V = 256 consecutive wavelengths (128 when V == 1), the same wavelengths per
// CTA as the fp64 kernel, so both write the same [rows][L][4] partials layout: a warp here covers
// RW = 4 / (THREADS / 32) rows of the fp64 layout and zero-fills the ones it does not use.
template <typename TabT, int DIR, int V, int THREADS, bool DTAUS>
__global__ void __launch_bounds__(THREADS) sweep_f32_kernel(SweepArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    constexpr int RW = 4 / (THREADS / 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    if (a.plan_hdr && blockIdx.x == 0 && b == 0 && tid == 0) a.plan_hdr[0] = a.rows;   // read by post_kernel
    if (a.active && !a.active[b]) return;
    const int L = a.L, S = a.S, rec8 = a.lp.rec8;
    double* sm_rec = smem;
    const TabT* slot = reinterpret_cast<const TabT*>(smem + (size_t)L * rec8) + tid * V;
    const uint32_t stage = smem_u32(slot);
    double* part = a.partials + ((int64_t)b * a.rows + blockIdx.x * 4 + warp * RW) * L * 4;
    const int64_t n_lam = a.n_lam;

    const uint32_t bytes = (uint32_t)((size_t)L * rec8 * 8);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const double* src = a.lp.rec + (int64_t)b * L * rec8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm_rec)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }

    const int64_t j_raw = ((int64_t)blockIdx.x * THREADS + tid) * V;
    const bool live = j_raw < n_lam;
    const int64_t j = live ? j_raw : n_lam - V;
    const TabT* tabj = static_cast<const TabT*>(a.tab) + j;
    const int64_t rowT = (int64_t)a.N_T * n_lam;
    float* Fu = static_cast<float*>(a.F_up) + (int64_t)b * L * n_lam + j;
    float* Fd = static_cast<float*>(a.F_down) + (int64_t)b * L * n_lam + j;
    float* dt_out = DTAUS ? static_cast<float*>(a.dtaus) + (int64_t)b * L * n_lam + j : nullptr;
    LaneF<V, THREADS> t;
    ldv<V>(a.c1 + j, t.c1);
    ldv<V>(a.c2 + j, t.c2);
    ldv<V>(a.sigma + j, t.sg);
    ldv<V>(a.w + j, t.wj);
    const float sscale = a.sigma_scale ? (float)a.sigma_scale[b] : 1.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) { t.sg[v] *= sscale; if (!live) t.wj[v] = 0.0f; }
    if (DTAUS && live) {
        float one[V];
#pragma unroll
        for (int v = 0; v < V; ++v) one[v] = 1.0f;
        stv<V>(dt_out, one);
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
    }

    float F2u[V], F1d[V], dtau[V], oth[V], nxt[V], k[V];
    double red[4];
    auto publish = [&](int i) {
        const double r4 = reduce4(red[0], red[1], red[2], red[3], lane);
        if ((lane & 7) == 0) {
            part[i * 4 + (lane >> 3)] = r4;
#pragma unroll
            for (int x = 1; x < RW; ++x) part[(int64_t)x * L * 4 + i * 4 + (lane >> 3)] = 0.0;
        }
    };
    if (DIR == FREI_EMIT) {
        const double* rec = sm_rec + rec8;
        stage_rows_f<TabT, V, THREADS>(tabj, rec, S, n_lam, rowT, stage);
        ldv<V>(Fu + n_lam, t.Fcar);
        const float invT1 = (float)rec[1];
#pragma unroll
        for (int v = 0; v < V; ++v) t.Bcar[v] = planck_f(t.c1[v], t.c2[v], invT1);
        const float* pFd = Fd + 2 * n_lam;
        float* pFu_out = Fu + 2 * n_lam;
        float* pFd_out = Fd + n_lam;
        float* pdt = DTAUS ? dt_out + n_lam : nullptr;
        if (L > 2) ldv<V>(pFd, nxt);
        cp_async_wait_all();
        gather_f<TabT, V, THREADS>(slot, rec, S, t.sg, k);
        for (int i = 1; i < L - 1; ++i) {
            if (!(reinterpret_cast<const int64_t*>(rec + rec8)[2 + 5 * S] & 1))    // level i + 1: new cell
                stage_rows_f<TabT, V, THREADS>(tabj, rec + rec8, S, n_lam, rowT, stage);
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];
            pFd += n_lam;
            if (i + 1 < L - 1) ldv<V>(pFd, nxt);
            step_f<FREI_EMIT, V, THREADS, false>(t, k, (float)rec[0], oth, (float)rec[rec8 + 1], F2u, F1d, dtau, red);
            if (live) {
                stv<V>(pFu_out, F2u);
                stv<V>(pFd_out, F1d);
                if (DTAUS) stv<V>(pdt, dtau);
            }
            publish(i);
            pFu_out += n_lam; pFd_out += n_lam; rec += rec8;
            if (DTAUS) pdt += n_lam;
            cp_async_wait_all();
            gather_f<TabT, V, THREADS>(slot, rec, S, t.sg, k);
        }
        {
            ldv<V>(a.f_toa + j, oth);
            const float fscale = a.ftoa_scale ? (float)a.ftoa_scale[b] : 1.0f;
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] *= fscale;
            step_f<FREI_EMIT, V, THREADS, true>(t, k, (float)rec[0], oth, 0.0f, F2u, F1d, dtau, red);
            if (live) {
                stv<V>(pFd_out, F1d);
                if (DTAUS) stv<V>(pdt, dtau);
            }
            publish(L - 1);
        }
    } else {
        const double* rec = sm_rec + (size_t)(L - 2) * rec8;
        stage_rows_f<TabT, V, THREADS>(tabj, rec, S, n_lam, rowT, stage);
        ldv<V>(Fd + (int64_t)(L - 1) * n_lam, t.Fcar);
        const float invTt = (float)rec[rec8 + 1];
#pragma unroll
        for (int v = 0; v < V; ++v) t.Bcar[v] = planck_f(t.c1[v], t.c2[v], invTt);
        const float* pFu = Fu + (int64_t)(L - 2) * n_lam;
        float* pFu_out = Fu + (int64_t)(L - 1) * n_lam;
        float* pFd_out = Fd + (int64_t)(L - 2) * n_lam;
        float* pdt = DTAUS ? dt_out + n_lam : nullptr;
        ldv<V>(pFu, nxt);
        cp_async_wait_all();
        gather_f<TabT, V, THREADS>(slot, rec, S, t.sg, k);
        for (int i = L - 2; i >= 0; --i) {
            if (i > 0 && !(reinterpret_cast<const int64_t*>(rec)[2 + 5 * S] & 1))    // level i - 1: new cell
                stage_rows_f<TabT, V, THREADS>(tabj, rec - rec8, S, n_lam, rowT, stage);
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];
            pFu -= n_lam;
            if (i > 0) ldv<V>(pFu, nxt);
            step_f<FREI_ABSORB, V, THREADS, false>(t, k, (float)rec[0], oth, (float)rec[1], F2u, F1d, dtau, red);
            if (live) {
                stv<V>(pFu_out, F2u);
                stv<V>(pFd_out, F1d);
                if (DTAUS) stv<V>(pdt, dtau);
            }
            publish(i);
            pFu_out -= n_lam; pFd_out -= n_lam;
            if (DTAUS) pdt += n_lam;
            if (i > 0) {
                rec -= rec8;
                cp_async_wait_all();
                gather_f<TabT, V, THREADS>(slot, rec, S, t.sg, k);
            }
        }
    }
    if (lane < 4) {
#pragma unroll
        for (int x = 0; x < RW; ++x)
            part[(int64_t)x * L * 4 + ((DIR == FREI_EMIT) ? 0 : (L - 1)) * 4 + lane] = 0.0;
    }
}

template <typename TabT, int DIR, int V, int THREADS, bool DTAUS>
int launch_one(SweepArgs a, size_t smem, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_lam + (int64_t)THREADS * V - 1) / ((int64_t)THREADS * V));
    a.rows = (int)blocks * 4;                    // a CTA fills four rows of the partials
    a.n2 = 0;
    if (cudaFuncSetAttribute(sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS>,
                             cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        return frei_set_err(FREI_E_CUDA, "cudaFuncSetAttribute failed (fp32 sweep)");
    sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS><<<dim3(blocks, a.B), THREADS, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? FREI_OK : frei_set_err(FREI_E_CUDA, "fp32 sweep launch failed");
}

template <typename TabT, int V, int THREADS>
int launch_v(const SweepArgs& a, int direction, cudaStream_t st) {
    const size_t smem = (size_t)a.L * a.lp.rec8 * sizeof(double) + (size_t)4 * a.S * THREADS * V * sizeof(TabT);
    if (smem > 200 * 1024) return frei_set_err(FREI_E_UNSUPPORTED, "level records exceed shared memory");
    if (direction == FREI_EMIT)
        return a.dtaus ? launch_one<TabT, FREI_EMIT, V, THREADS, true>(a, smem, st)
                       : launch_one<TabT, FREI_EMIT, V, THREADS, false>(a, smem, st);
    return a.dtaus ? launch_one<TabT, FREI_ABSORB, V, THREADS, true>(a, smem, st)
                   : launch_one<TabT, FREI_ABSORB, V, THREADS, false>(a, smem, st);
}

}  // namespace

// CTA size in wavelengths: 256 when plan_V >= 2 (V = 4 x 64 threads, or V = 2 x 128 threads),
// else 128 (V = 1: odd counts and very small problems).
int frei_launch_sweep_f32(SweepArgs a, int table_dtype, int direction, int plan_V, cudaStream_t st) {
    const int64_t n = a.n_lam;
    if (table_dtype == FREI_F32) {
        if (plan_V == 1) return launch_v<float, 1, 128>(a, direction, st);
        if (n % 4 == 0) return launch_v<float, 4, 64>(a, direction, st);
        return launch_v<float, 2, 128>(a, direction, st);
    }
    if (plan_V == 1) return launch_v<double, 1, 128>(a, direction, st);
    if (n % 4 == 0) return launch_v<double, 4, 64>(a, direction, st);
    return launch_v<double, 2, 128>(a, direction, st);
}
