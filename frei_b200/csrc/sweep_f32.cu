// sweep_f32.cu — the layer sweep in fp32 arithmetic (flux state and opacity table in fp32,
// wavelength integrals accumulated in fp64).  BASELINE.json allows 1e-4 relative per-wavelength
// flux error in fp32; the kernel is the same algorithm as sweep_kernel in frei_b200.cu
// (frei/twostream.py:351-405, 486-533) with
//   * hardware approximations for 1/x, 1/sqrt(x) and 2^x (MUFU.RCP / RSQ / EX2, ~2^-22),
//   * the cancellation-free grouping of the two-stream expressions (mandatory in fp32: the
//     reference's grouping loses all digits for delta_tau < 1e-3 in single precision),
//   * short series for 1 - exp(-u) and 1 - (1 - exp(-u))/u below u = 1/8,
//   * 4 wavelengths per thread (16-byte loads/stores, 16-byte cp.async of table rows),
//   * [r2] the arithmetic of two wavelengths in one instruction: sm_100's packed FFMA2 / FMUL2 / FADD2
//     (fma.rn.f32x2) take one issue slot for two lanes of work.  The scalar kernel executed 159
//     instructions per evaluation and was bound by the issue rate (60 % of the slots with 2.4 warps
//     per scheduler, 26 % of the HBM bandwidth, ncu capture r2g); the packed form needs about half,
//   * [r2] the E = 1 form of the layer response behind a warp vote (as in the fp64 kernel), level
//     records converted to fp32 once per CTA, programmatic dependent launch.
// Algorithmic traffic 4 S * 4 + 3 * 4 = 60 B per evaluation for S = 3.
#include "common.cuh"

#ifndef SWEEP_F32_PDL
#define SWEEP_F32_PDL 0           // 1: the fp32 sweep is launched with programmatic stream serialization.  Measured
                                  // (profiles/r02_f32_pdl_ab.log): C2 step 0.187 ms with it, 0.157 ms without — this
                                  // kernel leaves room on the SMs, so its CTAs are placed while the 1024-thread CTAs of
                                  // the reduction kernel still hold some of them and the wave ends up unbalanced
#endif
#ifndef SWEEP_F32_TRIGGER
#define SWEEP_F32_TRIGGER 1       // 1: its CTAs release the reduction kernel when their layer loop is done
#endif

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

// ---- two wavelengths per instruction (fma.rn.f32x2 and friends; MUFU and selects stay per lane) ----
typedef float2 f2;
__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 rcp2(f2 a) { return make_float2(rcpf(a.x), rcpf(a.y)); }
__device__ __forceinline__ f2 rsq2(f2 a) { return make_float2(rsqf(a.x), rsqf(a.y)); }
__device__ __forceinline__ f2 ex22(f2 a) { return make_float2(ex2f(a.x), ex2f(a.y)); }
__device__ __forceinline__ f2 sel2(bool px, bool py, f2 a, f2 b) { return make_float2(px ? a.x : b.x, py ? a.y : b.y); }

// T = exp(-u), m = 1 - exp(-u), e1 = 1 - m/u  (u >= 0), same series below u = 1/8 as exp_neg_f
__device__ __forceinline__ void exp_neg_2(f2 u, f2& T, f2& m, f2& e1) {
    T = ex22(mul2(u, bc(-1.4426950408889634f)));
    f2 p = fma2(u, bc(1.0f / 720.0f), bc(-1.0f / 120.0f));
    p = fma2(p, u, bc(1.0f / 24.0f));
    p = fma2(p, u, bc(-1.0f / 6.0f));
    p = fma2(p, u, bc(0.5f));
    const f2 e1s = mul2(p, u);                           // u/2 - u^2/6 + u^3/24 - u^4/120 + u^5/720
    const f2 ms = fma2(mul2(u, bc(-1.0f)), e1s, u);      // m = u (1 - e1)
    const f2 mb = fma2(T, bc(-1.0f), bc(1.0f));
    const f2 e1b = fma2(mul2(mb, rcp2(u)), bc(-1.0f), bc(1.0f));
    const bool sx = u.x < 0.125f, sy = u.y < 0.125f;
    m = sel2(sx, sy, ms, mb);
    e1 = sel2(sx, sy, e1s, e1b);
}

// c2n = -c2 log2(e): B = c1 e^-x / (1 - e^-x), x = c2 / T  (twostream.py:64-67; x >= 0.3 on this grid)
__device__ __forceinline__ f2 planck_2(f2 c1, f2 c2n, float invT) {
    const f2 e = ex22(mul2(c2n, bc(invT)));
    return mul2(mul2(c1, e), rcp2(fma2(e, bc(-1.0f), bc(1.0f))));
}

// two_stream_f for two wavelengths; signs are folded into the products (nchi = -chi, npsi = -psi,
// nA = -A) so that no instruction is spent on a negation.  E_IS_ONE: omega0 <= 0.1 in both lanes
// (E = 1, twostream.py:89-94): one rsqrt gives r = sqrt(1 - omega0) = k y and omega0 = sigma r y
// with y = 1/sqrt(k (sigma + k)), and pi (1 - omega0)/(E - omega0)/chi collapses to pi/chi.
template <bool E_IS_ONE>
__device__ __forceinline__ void two_stream_2(f2 k, f2 sg, float dpg, f2 F1u, f2 F2d, f2 B1, f2 B2,
                                             f2& F2u, f2& F1d, f2& dtau) {
    dtau = mul2(k, bc(dpg));                                             // :371-373
    f2 r, u, w0e, EmW, omw;
    if (E_IS_ONE) {
        const f2 y = rsq2(mul2(k, add2(sg, k)));
        r = mul2(k, y);                                                  // :143
        w0e = mul2(sg, mul2(r, y));                                      // omega0, :376-378
        u = mul2(r, add2(dtau, dtau));                                   // T = exp(-u), :139
    } else {
        const f2 R1 = rcp2(add2(sg, k));
        const f2 w0 = mul2(sg, R1);
        omw = mul2(k, R1);
        const bool hx = w0.x > 0.1f, hy = w0.y > 0.1f;                   // :89-94
        const f2 Ep = fma2(w0, fma2(w0, bc(-0.05582f), bc(-0.1777f)), bc(1.225f));
        const f2 Ew = sel2(hx, hy, Ep, bc(1.0f));
        const f2 invE = sel2(hx, hy, rcp2(Ep), bc(1.0f));
        EmW = fma2(w0, bc(-1.0f), Ew);
        const f2 q = mul2(Ew, EmW);
        const f2 a = mul2(q, rsq2(q));
        r = mul2(a, invE);
        u = mul2(a, add2(dtau, dtau));
        w0e = mul2(w0, invE);
    }
    f2 Tr, m, e1;
    exp_neg_2(u, Tr, m, e1);
    const f2 z = mul2(mul2(w0e, bc(0.5f)), rcp2(add2(r, bc(1.0f))));     // zeta_minus, :145
    const f2 zm = mul2(z, m), omzm = fma2(zm, bc(-1.0f), bc(1.0f));
    const f2 nchi = mul2(add2(r, zm), omzm);                             // -chi, :149
    const f2 xi = mul2(mul2(fma2(z, bc(-1.0f), bc(1.0f)), zm), fma2(m, bc(-1.0f), bc(2.0f)));   // :150
    const f2 npsi = mul2(r, Tr);                                         // -psi, :151
    const f2 nA = fma2(xi, bc(-2.0f), mul2(m, omzm));                    // -(chi + xi - psi)
    const f2 H = mul2(mul2(fma2(B2, bc(-1.0f), B1), r),
                      fma2(zm, fma2(e1, bc(-1.0f), bc(1.0f)), fma2(m, bc(-1.0f), e1)));
    const f2 nH = mul2(H, bc(-1.0f));
    const f2 up = fma2(npsi, F1u, mul2(xi, F2d)), dn = fma2(npsi, F2d, mul2(xi, F1u));
    const f2 su = fma2(B2, nA, nH), sd = fma2(B1, nA, H);
    if (E_IS_ONE) {
        const f2 nic = rcp2(nchi);                                       // -1/chi
        F2u = mul2(nic, fma2(bc(3.14159265358979f), su, up));            // :161-168
        F1d = mul2(nic, fma2(bc(3.14159265358979f), sd, dn));            // :169-176
    } else {
        const f2 nR3 = rcp2(mul2(EmW, nchi));
        const f2 nic = mul2(EmW, nR3);
        const f2 npc = mul2(mul2(omw, bc(3.14159265358979f)), nR3);      // -pi (1 - w0)/(E - w0)/chi, :152
        F2u = fma2(nic, up, mul2(npc, su));
        F1d = fma2(nic, dn, mul2(npc, sd));
    }
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
struct LaneF {
    float c1[V], c2[V], sg[V], wj[V], Fcar[V], Bcar[V];
    float c2n[V], thr[V];       // packed path: -c2 log2(e); 9.0001 sigma (k at or below it may have omega0 > 0.1)
};

template <int V> __device__ __forceinline__ f2 pr(const float* a, int p) { return make_float2(a[2 * p], a[2 * p + 1]); }

template <typename TabT, int V, int THREADS, int S_T>
__device__ __forceinline__ void stage_rows_f(const TabT* __restrict__ tabj, const double* rec, int S_rt,
                                             int64_t n_lam, int64_t rowT, uint32_t stage) {
    const int S = (S_T > 0) ? S_T : S_rt;        // compile-time species count: the loop unrolls, offsets are immediates
    constexpr int kSlot = V * (int)sizeof(TabT);
    constexpr uint32_t kRow = (uint32_t)THREADS * kSlot;
    const int64_t* off = reinterpret_cast<const int64_t*>(rec) + 2 + 4 * S;
#pragma unroll
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

// Packed path (V = 2 or 4).  frow = this level's fp32 record: W[S][4], then dpg, 1/T.
template <typename TabT, int V, int THREADS, int S_T>
__device__ __forceinline__ void gather_f2(const TabT* slot, const float* frow, int S_rt, const float* sg, float* k) {
    constexpr int kRowElems = THREADS * V, P = V / 2;
    const int S = (S_T > 0) ? S_T : S_rt;
    f2 acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = make_float2(sg[2 * p], sg[2 * p + 1]);   // k includes sigma, opacity.py:269
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float4 w = *reinterpret_cast<const float4*>(frow + 4 * s);
        float t0[V], t1[V], t2[V], t3[V];
        ldv<V>(slot + (4 * s + 0) * kRowElems, t0);
        ldv<V>(slot + (4 * s + 1) * kRowElems, t1);
        ldv<V>(slot + (4 * s + 2) * kRowElems, t2);
        ldv<V>(slot + (4 * s + 3) * kRowElems, t3);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            acc[p] = fma2(pr<V>(t0, p), bc(w.x), acc[p]);
            acc[p] = fma2(pr<V>(t1, p), bc(w.y), acc[p]);
            acc[p] = fma2(pr<V>(t2, p), bc(w.z), acc[p]);
            acc[p] = fma2(pr<V>(t3, p), bc(w.w), acc[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) { k[2 * p] = acc[p].x; k[2 * p + 1] = acc[p].y; }
}

template <int DIR, int V, int THREADS, bool SAME_T, bool E_IS_ONE>
__device__ __forceinline__ void tail_f2(LaneF<V, THREADS>& t, const float* k, float dpg, const float* other,
                                        float invTn, float* F2u, float* F1d, float* dtau, double* red) {
    constexpr int P = V / 2;
    f2 r0 = bc(0.f), r1 = bc(0.f), r2 = bc(0.f), r3 = bc(0.f);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const f2 Fc = pr<V>(t.Fcar, p), Bc = pr<V>(t.Bcar, p), ot = pr<V>(other, p), wj = pr<V>(t.wj, p);
        const f2 Bn = SAME_T ? Bc : planck_2(pr<V>(t.c1, p), pr<V>(t.c2n, p), invTn);
        f2 a, d, dt;
        if (DIR == FREI_EMIT) {
            two_stream_2<E_IS_ONE>(pr<V>(k, p), pr<V>(t.sg, p), dpg, Fc, ot, Bc, Bn, a, d, dt);
            r0 = fma2(wj, a, r0); r1 = fma2(wj, ot, r1); r2 = fma2(wj, Fc, r2); r3 = fma2(wj, d, r3);
            t.Fcar[2 * p] = a.x; t.Fcar[2 * p + 1] = a.y;
        } else {
            two_stream_2<E_IS_ONE>(pr<V>(k, p), pr<V>(t.sg, p), dpg, ot, Fc, Bn, Bc, a, d, dt);
            r0 = fma2(wj, a, r0); r1 = fma2(wj, Fc, r1); r2 = fma2(wj, ot, r2); r3 = fma2(wj, d, r3);
            t.Fcar[2 * p] = d.x; t.Fcar[2 * p + 1] = d.y;
        }
        t.Bcar[2 * p] = Bn.x; t.Bcar[2 * p + 1] = Bn.y;
        F2u[2 * p] = a.x; F2u[2 * p + 1] = a.y;
        F1d[2 * p] = d.x; F1d[2 * p + 1] = d.y;
        dtau[2 * p] = dt.x; dtau[2 * p + 1] = dt.y;
    }
    red[0] = (double)(r0.x + r0.y); red[1] = (double)(r1.x + r1.y);
    red[2] = (double)(r2.x + r2.y); red[3] = (double)(r3.x + r3.y);
}

// The warp votes on omega0 > 0.1 before any division (k <= 9.0001 sigma, a superset: the general
// form is right for every omega0) and takes the E = 1 form when no lane needs Deitrick's E(omega0).
template <int DIR, int V, int THREADS, bool SAME_T>
__device__ __forceinline__ void step_f2(LaneF<V, THREADS>& t, const float* k, float dpg, const float* other,
                                        float invTn, float* F2u, float* F1d, float* dtau, double* red) {
    bool hi = false;
#pragma unroll
    for (int v = 0; v < V; ++v) hi = hi || (k[v] <= t.thr[v]);
    if (__any_sync(0xffffffffu, hi))
        tail_f2<DIR, V, THREADS, SAME_T, false>(t, k, dpg, other, invTn, F2u, F1d, dtau, red);
    else
        tail_f2<DIR, V, THREADS, SAME_T, true>(t, k, dpg, other, invTn, F2u, F1d, dtau, red);
}

template <int DIR, int V, int THREADS, bool SAME_T>
__device__ __forceinline__ void step_any(LaneF<V, THREADS>& t, const float* k, float dpg, const float* other,
                                         float invTn, float* F2u, float* F1d, float* dtau, double* red) {
    if constexpr (V >= 2) step_f2<DIR, V, THREADS, SAME_T>(t, k, dpg, other, invTn, F2u, F1d, dtau, red);
    else step_f<DIR, V, THREADS, SAME_T>(t, k, dpg, other, invTn, F2u, F1d, dtau, red);
}

// One CTA = THREADS * V = 256 consecutive wavelengths (128 when V == 1); every warp writes one
// [L][4] row of partials (the reduction reads the row count from the plan header).  S_T = species
// count at compile time (0 = any): the staging and gather loops unroll.
template <typename TabT, int DIR, int V, int THREADS, bool DTAUS, int S_T>
__global__ void __launch_bounds__(THREADS) sweep_f32_kernel(SweepArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    constexpr int kW = THREADS / 32;             // one [L][4] row of partials per warp
    constexpr bool PACKED = (V >= 2);            // two wavelengths per instruction (FFMA2), fp32 records, E = 1 vote
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    asm volatile("griddepcontrol.wait;" ::: "memory");                  // records, T, flags: previous kernel
    if (a.plan_hdr && blockIdx.x == 0 && b == 0 && tid == 0) a.plan_hdr[0] = a.rows;   // read by the post kernel
    if (a.active && !a.active[b]) return;
    const int L = a.L, S = a.S, rec8 = a.lp.rec8;
    double* sm_rec = smem;
    const int RF = 4 * S + 4;                    // fp32 record: W[S][4], dpg, 1/T, 2 pad (16-byte rows)
    float* sm_frec = reinterpret_cast<float*>(smem + (size_t)L * rec8) ;
    const TabT* slot = reinterpret_cast<const TabT*>(sm_frec + (PACKED ? (size_t)L * RF : 0)) + tid * V;
    const uint32_t stage = smem_u32(slot);
    double* part = part_base(a.partials, b, a.rows, L, blockIdx.x * kW + warp);
    const int64_t pstride = part_level_stride(a.rows);
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
    for (int v = 0; v < V; ++v) {
        t.sg[v] *= sscale; if (!live) t.wj[v] = 0.0f;
        t.c2n[v] = -1.4426950408889634f * t.c2[v];
        t.thr[v] = 9.0001f * t.sg[v];
    }
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
#if POST_PER_LEVEL
    for (int lev = tid; lev < L; lev += THREADS) {           // same-rows flags from the staged offsets (see sweep_kernel)
        int64_t* r1 = reinterpret_cast<int64_t*>(sm_rec + (size_t)lev * rec8);
        int64_t same = lev > 0;
        if (lev > 0)
            for (int s = 0; s < S; ++s) if (r1[2 + 4 * S + s] != (r1 - rec8)[2 + 4 * S + s]) same = 0;
        r1[2 + 5 * S] = same;
    }
    __syncthreads();
#endif
    if (PACKED) {                                // the records once more, in fp32
        const int per = 4 * S + 2;
        for (int e = tid; e < L * per; e += THREADS) {
            const int i = e / per, c = e - i * per;
            const double* rec = sm_rec + (size_t)i * rec8;
            sm_frec[i * RF + c] = (float)(c < 4 * S ? rec[2 + c] : rec[c - 4 * S]);
        }
        __syncthreads();
    }

    float F2u[V], F1d[V], dtau[V], oth[V], nxt[V], k[V];
    // one layer-step / one gather, packed or scalar
    // Table rows: copied (cp.async) into the thread's slots one level ahead of their use, only when the
    // (P, T) cell changes.  (Staging the next cell's rows a whole cell ahead into a second buffer was
    // measured in round 2: no difference — the kernel waits on MUFU / shared-memory / shuffle latencies
    // with 2.4 warps per scheduler, not on HBM: ncu r2h, long_scoreboard 0.44 of 5.9 stall cycles.)
    auto flag_same = [&](int lev) -> bool {      // level lev uses the same table rows as level lev - 1
        return reinterpret_cast<const int64_t*>(sm_rec + (size_t)lev * rec8)[2 + 5 * S] & 1;
    };
    auto stage_level = [&](int lev) {
        stage_rows_f<TabT, V, THREADS, S_T>(tabj, sm_rec + (size_t)lev * rec8, S, n_lam, rowT, stage);
    };
    auto gather = [&](const double* rec, int lev) {
        if constexpr (PACKED) gather_f2<TabT, V, THREADS, S_T>(slot, sm_frec + (size_t)lev * RF, S, t.sg, k);
        else gather_f<TabT, V, THREADS>(slot, rec, S, t.sg, k);
    };
    auto dpg_of = [&](int lev) -> float {
        return PACKED ? sm_frec[lev * RF + 4 * S] : (float)sm_rec[(size_t)lev * rec8];
    };
    auto invT_of = [&](int lev) -> float {
        return PACKED ? sm_frec[lev * RF + 4 * S + 1] : (float)sm_rec[(size_t)lev * rec8 + 1];
    };
    double red[4];
    auto publish = [&](int i) {
        const double r4 = reduce4(red[0], red[1], red[2], red[3], lane);
        if ((lane & 7) == 0) part[i * pstride + (lane >> 3)] = r4;
    };
    if (DIR == FREI_EMIT) {
        const double* rec = sm_rec + rec8;
        stage_level(1);
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
        gather(rec, 1);
        for (int i = 1; i < L - 1; ++i) {
            if (!flag_same(i + 1)) stage_level(i + 1);          // level i + 1: new cell
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];
            pFd += n_lam;
            if (i + 1 < L - 1) ldv<V>(pFd, nxt);
            step_any<FREI_EMIT, V, THREADS, false>(t, k, dpg_of(i), oth, invT_of(i + 1), F2u, F1d, dtau, red);
            if (live) {
                stv<V>(pFu_out, F2u);
                stv<V>(pFd_out, F1d);
                if (DTAUS) stv<V>(pdt, dtau);
            }
            publish(i);
            pFu_out += n_lam; pFd_out += n_lam; rec += rec8;
            if (DTAUS) pdt += n_lam;
            cp_async_wait_all();
            gather(rec, i + 1);
        }
        {
            ldv<V>(a.f_toa + j, oth);
            const float fscale = a.ftoa_scale ? (float)a.ftoa_scale[b] : 1.0f;
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] *= fscale;
            step_any<FREI_EMIT, V, THREADS, true>(t, k, dpg_of(L - 1), oth, 0.0f, F2u, F1d, dtau, red);
            if (live) {
                stv<V>(pFd_out, F1d);
                if (DTAUS) stv<V>(pdt, dtau);
            }
            publish(L - 1);
        }
    } else {
        const double* rec = sm_rec + (size_t)(L - 2) * rec8;
        stage_level(L - 2);
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
        gather(rec, L - 2);
        for (int i = L - 2; i >= 0; --i) {
            if (i > 0 && !flag_same(i)) stage_level(i - 1);      // level i - 1: new cell
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];
            pFu -= n_lam;
            if (i > 0) ldv<V>(pFu, nxt);
            step_any<FREI_ABSORB, V, THREADS, false>(t, k, dpg_of(i), oth, invT_of(i), F2u, F1d, dtau, red);
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
                gather(rec, i - 1);
            }
        }
    }
    if (lane < 4) part[((DIR == FREI_EMIT) ? 0 : (L - 1)) * pstride + lane] = 0.0;     // the level the sweep does not visit
    // The reduction kernel may become resident now.  (Not earlier: this kernel leaves room on every SM,
    // so CTAs of 1024 threads parked in griddepcontrol.wait from the start displace sweep CTAs into a
    // second wave — measured: +9 us per sweep.)
    if (SWEEP_F32_TRIGGER && gridDim.y == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename TabT, int DIR, int V, int THREADS, bool DTAUS, int S_T>
int launch_one(SweepArgs a, size_t smem, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n_lam + (int64_t)THREADS * V - 1) / ((int64_t)THREADS * V));
    a.rows = (int)blocks * (THREADS / 32);       // one row of partials per warp
    a.n2 = 0;
    if (cudaFuncSetAttribute(sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS, S_T>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS, S_T>,
                             cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        return frei_set_err(FREI_E_CUDA, "cudaFuncSetAttribute failed (fp32 sweep)");
    // optional programmatic dependent launch (SWEEP_F32_PDL, off: see there); griddepcontrol.wait in the
    // kernel returns at once for a normally serialised launch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, a.B); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = SWEEP_F32_PDL;
    return cudaLaunchKernelEx(&cfg, sweep_f32_kernel<TabT, DIR, V, THREADS, DTAUS, S_T>, a) == cudaSuccess
               ? FREI_OK : frei_set_err(FREI_E_CUDA, "fp32 sweep launch failed");
}

template <typename TabT, int V, int THREADS, int S_T>
int launch_s(const SweepArgs& a, int direction, size_t smem, cudaStream_t st) {
    if (direction == FREI_EMIT)
        return a.dtaus ? launch_one<TabT, FREI_EMIT, V, THREADS, true, S_T>(a, smem, st)
                       : launch_one<TabT, FREI_EMIT, V, THREADS, false, S_T>(a, smem, st);
    return a.dtaus ? launch_one<TabT, FREI_ABSORB, V, THREADS, true, S_T>(a, smem, st)
                   : launch_one<TabT, FREI_ABSORB, V, THREADS, false, S_T>(a, smem, st);
}

template <typename TabT, int V, int THREADS>
int launch_v(const SweepArgs& a, int direction, cudaStream_t st) {
    // level records, staging slots of 4 S rows, fp32 records of the packed path
    const size_t smem = (size_t)a.L * a.lp.rec8 * sizeof(double) + (size_t)4 * a.S * THREADS * V * sizeof(TabT) +
                        (V >= 2 ? (size_t)a.L * (4 * a.S + 4) * sizeof(float) : 0);
    if (smem > 200 * 1024) return frei_set_err(FREI_E_UNSUPPORTED, "level records exceed shared memory");
    // species counts of the BASELINE configs get their own instantiation of the main (fp32 table, 4 per thread) kernel
    if (sizeof(TabT) == 4 && V == 4) {
        if (a.S == 3) return launch_s<TabT, V, THREADS, 3>(a, direction, smem, st);
        if (a.S == 8) return launch_s<TabT, V, THREADS, 8>(a, direction, smem, st);
    }
    return launch_s<TabT, V, THREADS, 0>(a, direction, smem, st);
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
