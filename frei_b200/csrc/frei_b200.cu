// frei_b200.cu — sm_100a kernels and the C ABI (include/frei_b200.h) of the
// radiative-equilibrium hot path: (P,T) bracket + weights (K0), opacity gather
// (K1), two-stream layer response (K2), the fused layer sweep with
// wavelength-integral partials (K2+K3), the fixed-order reduction and the
// per-layer temperature update (K4).
//
// Reference behaviour restated here (paths into the reference checkout):
//   frei/opacity.py:173-269   kappa(), Rayleigh
//   frei/twostream.py:16-287  bolometric_flux, BB, E, propagate_fluxes, layer thermodynamics
//   frei/twostream.py:351-416 emit loop body,  :486-545 absorb loop body
// No tensor cores: nothing here is a dense contraction.  The sweep is a
// one-thread-per-wavelength serial recurrence over layers held in registers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "common.cuh"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
// used by the other translation units of the library
int frei_set_err(int code, const char* msg) { return set_err(code, "%s%s", msg); }

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess)                                                     \
            return set_err(FREI_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define ARG_TRY(cond)                                                               \
    do {                                                                            \
        if (!(cond)) return set_err(FREI_E_ARG, "bad argument: %s%s", #cond);       \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// sweep_kernel and the post kernel (post_level_kernel) alternate on one stream, each consuming what
// the other wrote.  Both are launched with programmatic stream serialisation: the next kernel's CTAs
// may become resident while the previous kernel drains and block in griddepcontrol.wait — which returns when the previous grid has completed and
// its writes are visible — before they touch any global data.  This hides the launch gap between
// the four kernels of an RE iteration.  -DFREI_PDL=0 restores plain stream order.
#ifndef FREI_PDL
#define FREI_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if FREI_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if FREI_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
#if FREI_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
#else
    kern<<<grid, block, smem, st>>>(args...);
    return cudaGetLastError();
#endif
}

static inline int64_t layer_params_bytes(int B, int L, int S) {
    return (int64_t)B * L * rec_words(S) * 8;
}
static inline LayerParams layer_params_view(void* p, int S) {
    LayerParams v;
    v.rec = (double*)p;
    v.rec8 = rec_words(S);
    v.S = S;
    return v;
}
static inline int64_t round16(int64_t x) { return (x + 15) & ~int64_t(15); }
static inline int64_t sweep_rows_max(int64_t n_lam) { return (n_lam + 31) / 32 + 2 * kWarps; }   // bound on the plan's rows
// ws->partials: [B][rows_max][L][4] chunk rows | [B][kPostChunks][L][4] stage-1 sums | [B] tickets | header | [rows_max] relay flags
//               | [B][L] new temperatures | [B] converged-level counts
static inline double* ws_chunk_sums(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return ws->partials + (int64_t)B * sweep_rows_max(n_lam) * L * 4;
}
static inline unsigned int* ws_counters(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return reinterpret_cast<unsigned int*>(ws_chunk_sums(ws, B, L, n_lam) + (int64_t)B * kPostChunks * L * 4);
}
static inline int32_t* ws_plan_hdr(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws_counters(ws, B, L, n_lam)) + round16((int64_t)B * 4));
}
// relay flags, one per chunk row of a single atmosphere (zero between launches)
static inline unsigned int* ws_relay_flags(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return reinterpret_cast<unsigned int*>(ws_plan_hdr(ws, B, L, n_lam) + 4);
}
// per-level reduction (post_level_kernel): the new temperatures until the last CTA of an atmosphere
// publishes them, and the count of converged levels
static inline double* ws_T_next(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(ws_relay_flags(ws, B, L, n_lam)) +
                                     round16(sweep_rows_max(n_lam) * 4));
}
static inline unsigned int* ws_conv_count(const frei_workspace* ws, int B, int L, int64_t n_lam) {
    return reinterpret_cast<unsigned int*>(ws_T_next(ws, B, L, n_lam) + (int64_t)B * L);
}

// ---------------------------------------------------------------------------
// PROBE build only: time stamps (globaltimer, ns) of the sweep -> post -> sweep hand-over
#ifdef POST_STAMPS
__device__ unsigned long long g_stamps[32];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define STAMP_MAX(i) atomicMax(&g_stamps[i], gtime())
#define STAMP_MIN(i) atomicMin(&g_stamps[i], gtime())
extern "C" int frei_b200_debug_stamps(unsigned long long* out, int reset) {
    if (reset) { unsigned long long z[32]; for (int i = 0; i < 32; ++i) z[i] = (i & 1) ? 0ull : ~0ull; return (int)cudaMemcpyToSymbol(g_stamps, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(out, g_stamps, sizeof(g_stamps));
}
#else
#define STAMP_MAX(i)
#define STAMP_MIN(i)
#endif
// K0: brackets, weights, per-layer scalars
// ---------------------------------------------------------------------------
// scipy find_indices rule: below grid -> 0; >= last node -> n-2; else x[i] <= v < x[i+1].
__device__ __forceinline__ int bracket_index(const double* __restrict__ x, int n, double v) {
    if (v < x[0]) return 0;
    if (v >= x[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;              // invariant x[lo] <= v < x[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

struct PrepArgs {
    const double* axis_P; const double* axis_T; const int32_t* has_T;
    const double* T; const double* P; const double* mmr; const double* g;
    LayerParams lp;
    int32_t* iP; int32_t* iT; double* wP; double* wT; uint8_t* oob;
    int B, L, S, N_P, N_T;
    int64_t n_lam;
};

// Where a block finds the per-level inputs of atmosphere b.  Either views of global memory or a
// staged copy in shared memory (post_kernel copies them before it waits for the sweep, see there).
struct LevelView {
    const double* T;        // [L]
    const double* P;        // [L]
    const double* mmr;      // [L][S]
    int64_t* off;           // [L][S] shared-memory scratch for the row offsets, or null
};
__device__ __forceinline__ LevelView global_levels(const double* T, const double* P, const double* mmr,
                                                   int b, int L, int S) {
    LevelView v;
    v.T = T + (int64_t)b * L; v.P = P + (int64_t)b * L; v.mmr = mmr ? mmr + (int64_t)b * L * S : nullptr;
    v.off = nullptr;
    return v;
}

// Copy the axes of K0 into shared memory (S (N_P + N_T) doubles); the caller synchronises.
__device__ __forceinline__ void stage_axes(const double* axis_P, const double* axis_T, int S, int N_P, int N_T,
                                           double* sm_axes) {
    for (int e = threadIdx.x; e < S * N_P; e += blockDim.x) sm_axes[e] = axis_P[e];
    for (int e = threadIdx.x; e < S * N_T; e += blockDim.x) sm_axes[S * N_P + e] = axis_T[e];
}

// Block-cooperative K0 for atmosphere b: the level records of all L levels.
// The bracket search is a chain of dependent loads, ~10 L2 round trips per (level, species) when
// the nodes are read from global memory — it was half of the 26 us of the kernel that follows
// every sweep — so the axes are searched in shared memory (`sm_axes`, filled by stage_axes and
// synchronised by the caller; null = search in global memory).  One thread per (level, species)
// pair, then one per level for the scalars and the same-cell flag.  Callers that have just
// written T synchronise first; this function ends with all records written (no trailing barrier).
__device__ __forceinline__ double fast_rcp(double x);     // defined with the other fp64 building blocks below

// K0 for one (level, species) pair: bracket of (vp, vt) in the species' axes, the four mmr-premultiplied
// corner weights and the row offset, written into the level's record; returns the offset.
__device__ __forceinline__ int64_t prep_pair(const PrepArgs& a, int b, int i, int s, const double* axP,
                                             const double* axT, const int32_t* has_T, double vp, double vt,
                                             double m) {
    const int S = a.S;
    const int64_t li = (int64_t)b * a.L + i;
    double* rec = a.lp.rec + li * a.lp.rec8;
    const double* xp = axP + (int64_t)s * a.N_P;
    const double* xt = axT + (int64_t)s * a.N_T;
    const int ip = bracket_index(xp, a.N_P, vp);
    const double wp = (vp - xp[ip]) / (xp[ip + 1] - xp[ip]);
    bool out = (vp < xp[0]) || (vp > xp[a.N_P - 1]);
    int it = 0; double wt = 0.0;
    if (has_T[s]) {
        it = bracket_index(xt, a.N_T, vt);
        wt = (vt - xt[it]) / (xt[it + 1] - xt[it]);
        out = out || (vt < xt[0]) || (vt > xt[a.N_T - 1]);
    }
    double w00 = (1.0 - wp) * (1.0 - wt), w01 = (1.0 - wp) * wt;
    double w10 = wp * (1.0 - wt), w11 = wp * wt;
    if (out) { w00 = w01 = w10 = w11 = 0.0; }       // fill_value=0, opacity.py:243
    double* W = rec + 2 + 4 * s;
    W[0] = m * w00; W[1] = m * w01; W[2] = m * w10; W[3] = m * w11;
    const int64_t off = (int64_t)((s * a.N_P + ip) * a.N_T + it) * a.n_lam;
    reinterpret_cast<int64_t*>(rec)[2 + 4 * S + s] = off;
    if (a.iP) a.iP[li * S + s] = ip;
    if (a.iT) a.iT[li * S + s] = it;
    if (a.wP) a.wP[li * S + s] = wp;
    if (a.wT) a.wT[li * S + s] = wt;
    if (a.oob) a.oob[li * S + s] = out ? 1 : 0;
    return off;
}

// `has_T` = a.has_T or a copy of it in shared memory, `g` = a.g[b] (the post kernel loads both before it
// waits for the sweep: a load from global memory here is a full L2 round trip on the critical path).
// `dpg_mine` (nullable): (p1 - p2)/g of level threadIdx.x, computed ahead of time by the caller.
__device__ __forceinline__ void prep_block(const PrepArgs& a, int b, const double* sm_axes, const LevelView& lv,
                                           const int32_t* has_T, double g, const double* dpg_mine = nullptr) {
    const int L = a.L, S = a.S;
    const double* axP = sm_axes ? sm_axes : a.axis_P;
    const double* axT = sm_axes ? sm_axes + S * a.N_P : a.axis_T;
    const double* T = lv.T;
    const double* P = lv.P;
    for (int idx = threadIdx.x; idx < L * S; idx += blockDim.x) {
        const int i = idx / S, s = idx - i * S;
        const int64_t off = prep_pair(a, b, i, s, axP, axT, has_T, P[i], T[i], lv.mmr[idx]);
        if (lv.off) lv.off[idx] = off;
    }
    if (threadIdx.x == 0) { STAMP_MAX(31); }                         // probe: brackets of (level 0, species 0) written
    __syncthreads();                                     // offsets of the neighbouring level
    if (threadIdx.x == 0) { STAMP_MAX(15); }                         // probe: all brackets written
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const int64_t li = (int64_t)b * L + i;
        double* rec = a.lp.rec + li * a.lp.rec8;
        const double p1 = P[i] * FREI_BAR;
        double p2;
        if (i == L - 1) p2 = p1 * (P[L - 2] * FREI_BAR) / (P[L - 3] * FREI_BAR);   // twostream.py:359
        else p2 = P[i + 1] * FREI_BAR;
        rec[0] = (dpg_mine && i == (int)threadIdx.x) ? *dpg_mine : (p1 - p2) / g;   // twostream.py:231
        const double Ti = T[i];                          // 1/T: the kernels' reciprocal (<= 0.6 ulp) for ordinary
        rec[1] = (Ti > 1e-3 && Ti < 1e9) ? fast_rcp(Ti) : 1.0 / Ti;   // temperatures, IEEE division otherwise
        // same (P, T) cell as level i - 1 for every species?
        int64_t same = 0;
        if (i > 0) {
            same = 1;
            const int64_t* o1 = lv.off ? lv.off + (int64_t)i * S
                                       : reinterpret_cast<const int64_t*>(rec) + 2 + 4 * S;
            const int64_t* o0 = lv.off ? o1 - S : o1 - a.lp.rec8;
            for (int s = 0; s < S; ++s) if (o0[s] != o1[s]) same = 0;
        }
        reinterpret_cast<int64_t*>(rec)[2 + 5 * S] = same;
    }
}

// bytes of shared memory prep_block wants for the axes (0 = too large, search in global memory)
static inline size_t prep_axes_bytes(int S, int N_P, int N_T) {
    const size_t n = (size_t)S * ((size_t)N_P + N_T) * sizeof(double);
    return n <= 24 * 1024 ? n : 0;
}

__global__ void prep_kernel(PrepArgs a, int use_smem) {
    extern __shared__ double sm_prep[];
    if (use_smem) {
        stage_axes(a.axis_P, a.axis_T, a.S, a.N_P, a.N_T, sm_prep);
        __syncthreads();
    }
    prep_block(a, blockIdx.x, use_smem ? sm_prep : nullptr,
               global_levels(a.T, a.P, a.mmr, blockIdx.x, a.L, a.S), a.has_T, a.g[blockIdx.x]);
}

// ---------------------------------------------------------------------------
// per-wavelength constants: Planck prefactors, Rayleigh sigma, trapezoid weights, F_TOA
// ---------------------------------------------------------------------------
__global__ void spectral_kernel(const double* __restrict__ lam_um, int64_t n_global, int64_t off,
                                int64_t n_local, double m_bar, double T_star, double a_rstar, double f,
                                double* __restrict__ c1, double* __restrict__ c2,
                                double* __restrict__ sigma, double* __restrict__ w,
                                double* __restrict__ f_toa) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_local) return;
    const int64_t jg = off + j;
    const double lu = lam_um[jg], l = lu * 1e-4;
    const double k1 = 2.0 * FREI_H * FREI_C * FREI_C / pow(l, 5.0);
    c1[j] = k1;
    c2[j] = FREI_H * FREI_C / (l * FREI_KB);
    // Rayleigh, opacity.py:173-200
    const double nh2 = 13.58e-5 * (1.0 + 7.52e-11 / (l * l)) + 1.0;
    const double nhe = 1e-8 * (2283.0 + (1.8102e13 / (1.5342e10 - 1.0 / (lu * lu)))) + 1.0;
    const double pi3 = FREI_PI * FREI_PI * FREI_PI, l4 = (l * l) * (l * l);
    const double rh = (nh2 * nh2 - 1.0) / (nh2 * nh2 + 2.0), re = (nhe * nhe - 1.0) / (nhe * nhe + 2.0);
    const double s_h2 = (24.0 * pi3 / (2.68678e19 * 2.68678e19) / l4 * (rh * rh)) / m_bar;
    const double s_he = (24.0 * pi3 / (2.546899e19 * 2.546899e19) / l4 * (re * re)) / m_bar;
    sigma[j] = s_h2 + s_he;
    // trapezoid weights of the global grid (in cm)
    double wj;
    if (n_global == 1) wj = 0.0;
    else if (jg == 0) wj = 0.5 * (lam_um[1] * 1e-4 - l);
    else if (jg == n_global - 1) wj = 0.5 * (l - lam_um[jg - 1] * 1e-4);
    else wj = 0.5 * (lam_um[jg + 1] * 1e-4 - lam_um[jg - 1] * 1e-4);
    w[j] = wj;
    // F_TOA, core.py:48-62
    const double Bs = k1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T_star));
    f_toa[j] = f * (1.0 / (a_rstar * a_rstar)) * 1.0 / (2.0 * FREI_PI) * (FREI_PI * Bs);
}

// ---------------------------------------------------------------------------
// two-stream layer response, g_0 = 0   (twostream.py:139-176)
// ---------------------------------------------------------------------------
// ---- branch-free fp64 building blocks ---------------------------------------------------
// The CUDA library's '/', sqrt(), exp() and expm1() each carry special-case branches and
// call-outs that dominated the issue slots of the sweep (ncu, profiles/).  Arguments on this
// path are finite, positive and far from the denormal range, so plain Newton refinement of the
// hardware seeds is enough: results are within ~1 ulp (checked by tests through
// frei_b200_debug_math).
// The hardware seeds (MUFU.RCP64H / RSQ64H) are good to ~1e-6 (measured on B200,
// scripts/seed_probe.cu).  One third-order step — y0 (1 + e + e^2) for the reciprocal,
// y0 (1 + e/2 + 3 e^2/8) for the reciprocal square root, e the residual of the seed — leaves a
// truncation error of e^3 ~ 1e-18, so the result is the correctly rounded value up to ~0.6 ulp with
// 3 (5) fp64 instructions instead of the 4 (7) of two Newton steps.  -DFREI_SEED_REFINE=2 selects
// the two-step Newton form, =1 a single Newton step (1e-12, fails the parity tests; timing only).
#ifndef FREI_SEED_REFINE
#define FREI_SEED_REFINE 3
#endif
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
#if FREI_SEED_REFINE == 3
    return fma(y, fma(e, e, e), y);
#else
    y = fma(y, e, y);
#if FREI_SEED_REFINE == 2
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
#endif
    return y;
#endif
}

// sqrt(x) and 1/sqrt(x) for normal positive x
__device__ __forceinline__ double fast_sqrt(double x, double& rs) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if FREI_SEED_REFINE == 3
    const double e = fma(-(x * y), y, 1.0);
    y = fma(y, fma(0.375, e, 0.5) * e, y);
#else
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
#if FREI_SEED_REFINE == 2
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
#endif
#endif
    rs = y;
    return x * y;                   // ~1.5 ulp; a correction step is not worth 3 more fp64 slots
}

// 2^(j/32), j = 0..31 (correctly rounded); copied to shared memory by the kernels that use it
__constant__ double kExp2Tab[32] = {
    1, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.6457554781539649,
    1.681792830507429, 1.7186192981224779, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};

// For u >= 0: T = exp(-u) and m = 1 - exp(-u).  -u = n ln2/32 + r with |r| <= ln2/64, so
// exp(-u) = 2^(n >> 5) * 2^((n & 31)/32) * (1 + p(r)) with a degree-6 p: 14 fp64 instructions with
// mostly constant operands (the fp64 pipe issues 3-register DFMAs at 2/3 rate).
// T is good to ~1.5 ulp; m = -p when n == 0 (u <= 0.0108, no cancellation), else 1 - T.
// uq = u Q(r) = 1 - m/u is valid when `small` (n == 0) and lets the caller form (T - 1)/u + 1
// without cancellation.  exp(-u) is flushed to exp(-708) ~ 3e-308 beyond u = 708.
__device__ __forceinline__ void exp_neg(double u, const double* tab, double& T, double& m, double& uq,
                                        bool& small) {
    const double MAGIC = 6755399441055744.0;                   // 1.5 * 2^52
    // clamp at 708 on the high word (u >= 0: the integer order of the high words is the order of
    // the doubles; 0x40862000 = high word of 708.0, so u < 708.0003): one integer min instead of
    // the compare + selects of fmin(); beyond it 2^(n >> 5) underflows to 0 or ~1e-308
    u = __hiloint2double(min(__double2hiint(u), 0x40862000), __double2loint(u));
    const double tn = fma(u, -46.166241308446828, MAGIC);      // -u * 32 / ln 2
    const int n = __double2loint(tn);                          // rint, <= 0
    const double fn = tn - MAGIC;
    double r = fma(fn, -0.021660849392446835, -u);             // ln2/32 hi (36 bits: n * hi exact)
    r = fma(fn, -5.1456092446553382e-14, r);                   // ln2/32 lo
#ifndef FREI_EXP_DEG
#define FREI_EXP_DEG 6            // degree of the expm1 polynomial: 6 (truncation 3e-18) or 5 (2e-15; experiment)
#endif
#if FREI_EXP_DEG == 6
    double q = fma(1.3888888888888889e-03, r, 8.3333333333333332e-03);   // 1/6!, 1/5!
    q = fma(q, r, 4.1666666666666664e-02);                     // 1/4!
#else
    double q = fma(8.3333333333333332e-03, r, 4.1666666666666664e-02);   // 1/5!, 1/4!
#endif
    q = fma(q, r, 1.6666666666666666e-01);                     // 1/3!
    q = fma(q, r, 0.5);                                        // 1/2!
    const double p = fma(r * r, q, r);                         // expm1(r)
    const double t0 = tab[n & 31];
    const double sc = __hiloint2double(((n >> 5) + 1023) << 20, 0);      // 2^(n >> 5), n >> 5 >= -1022
    T = fma(t0, p, t0) * sc;
    small = (n == 0);
    m = small ? -p : 1.0 - T;
    uq = u * q;
}
__device__ __forceinline__ void exp_neg(double u, const double* tab, double& T, double& m) {
    double uq; bool small;
    exp_neg(u, tab, T, m, uq, small);
}

__device__ __forceinline__ double planck(double c1, double c2, double invT, const double* tab) {
    double e, m;                                         // c1 / expm1(x) = c1 e^-x / (1 - e^-x)
    exp_neg(c2 * invT, tab, e, m);                       // twostream.py:64-67
    return c1 * e * fast_rcp(m);
}

// Algebraically identical to the reference's expressions (twostream.py:139-176), regrouped so
// that no step subtracts nearly equal numbers.  With z = zeta_minus, m = 1 - T, u = -ln T:
//   zeta_minus = (omega0 / E) / (2 (1 + r))          r = sqrt((E - omega0) / E)
//   chi = -(r + z m)(1 - z m)     xi = (1 - z) z m (2 - m)     psi = -r T
//   chi - psi - xi = -m (1 - z m) = -(xi + psi - chi)         B'/(2E) = (B1 - B2) r / u
// The reference's own grouping loses up to ~4e-6 relative accuracy where delta_tau < 1e-6
// ("conditioning"); this one tracks the exact value of the same formulas.
// k = total opacity (includes sigma, opacity.py:269), sg = sigma, dpg = (p1 - p2)/g.
// two_stream_front gives delta_tau, omega0 and 1 - omega0; two_stream_tail<E_IS_ONE> the rest
// (E_IS_ONE = true drops the quadratic of Deitrick's E(omega0), twostream.py:89-94, its reciprocal
// and the multiplications by 1 for callers that know omega0 <= 0.1).  The sweep uses this general
// form only for warps in which some lane has omega0 > 0.1; all others take two_stream_E1 below.
__device__ __forceinline__ void two_stream_front(double k, double sg, double dpg, double& dtau,
                                                 double& w0, double& omw) {
    dtau = dpg * k;                                                     // :371-373
    const double R1 = fast_rcp(sg + k);
    w0 = sg * R1; omw = k * R1;                                         // omega0 (:376-378), 1 - omega0
}

template <bool E_IS_ONE>
__device__ __forceinline__ void two_stream_tail(double dtau, double w0, double omw, double F1u, double F2d,
                                                double B1, double B2, const double* tab, double& F2u,
                                                double& F1d) {
    // Deitrick (2020) Eqn 19 with g_0 = 0 (:89-94), select instead of branch so the whole
    // layer step stays one basic block for the instruction scheduler
    double Ew = 1.0, invE = 1.0;
    if (!E_IS_ONE) {
        const bool hi = w0 > 0.1;
        const double Ep = 1.225 - 0.1777 * w0 - 0.05582 * (w0 * w0);
        Ew = hi ? Ep : 1.0;
        invE = hi ? fast_rcp(Ep) : 1.0;
    }
    const double EmW = Ew - w0;
    double rs;
    const double a = fast_sqrt(E_IS_ONE ? EmW : Ew * EmW, rs);
    const double r = E_IS_ONE ? a : a * invE;                           // sqrt((E - w0)/E), :143
    const double u = 2.0 * a * dtau;                                    // T = exp(-u), :139
    double Tr, m, uq;
    bool small;
    exp_neg(u, tab, Tr, m, uq, small);
    const double opr = 1.0 + r;
    const double R2 = fast_rcp(opr * u);
    const double z = 0.5 * (E_IS_ONE ? w0 : w0 * invE) * (u * R2);      // zeta_minus, :145
    const double e1 = small ? uq : fma(-(m * opr), R2, 1.0);            // 1 - m/u = (T - 1)/u + 1
    const double zm = z * m, omzm = 1.0 - zm;
    const double chi = -(r + zm) * omzm;                                // :149
    const double xi = (1.0 - z) * zm * (2.0 - m);                       // :150
    const double psi = -r * Tr;                                         // :151
    const double A = fma(2.0, xi, -m * omzm);                           // chi + xi - psi
    // T + ((T - 1)/u)(1 - z m) = (e1 - m) + z m (1 - e1): no O(1) terms cancel for small u
    const double H = (B1 - B2) * r * fma(zm, 1.0 - e1, e1 - m);         // psi D + B'/(2E)(chi-psi-xi)
    const double R3 = fast_rcp(EmW * chi);
    const double ic = EmW * R3;                                         // 1 / chi
    const double pc = (FREI_PI * omw) * R3;                             // pi (1 - w0)/(E - w0)/chi, :152
    F2u = fma(ic, fma(psi, F1u, -xi * F2d), pc * fma(B2, A, H));        // :161-168
    F1d = fma(ic, fma(psi, F2d, -xi * F1u), pc * fma(B1, A, -H));       // :169-176
}

// Layer response for omega0 < 0.1 (E = 1, twostream.py:89-94), taken when the warp vote finds no
// lane that needs Deitrick's correction.  With E = 1 the reference's expressions collapse further:
//   s = sigma + k,  1 - omega0 = k / s,  r = sqrt(1 - omega0) = k y  with  y = 1 / sqrt(k s)
//   omega0 = sigma / s = sigma (r y)            (one rsqrt replaces the reciprocal AND the sqrt)
//   pi (1 - omega0) / (E - omega0) / chi = pi / chi     (no second factor to carry, :152)
// which is 8 fp64 instructions and one MUFU fewer per evaluation than the general form.  The
// results agree with the general form to rounding (~1e-16 relative), not bit for bit.
// dpg2 = 2 (p1 - p2)/g.
__device__ __forceinline__ void two_stream_E1(double k, double sg, double dpg2, double F1u, double F2d,
                                              double B1, double B2, const double* tab, double& F2u,
                                              double& F1d) {
    const double s = sg + k;
    double y;
    (void)fast_sqrt(k * s, y);                                          // y = 1/sqrt(k s)
    const double r = k * y;                                             // sqrt((E - w0)/E), :143
    const double w0 = sg * (r * y);                                     // omega0, :376-378
    const double u = r * (dpg2 * k);                                    // T = exp(-u), :139
    double Tr, m, uq;
    bool small;
    exp_neg(u, tab, Tr, m, uq, small);
    const double opr = 1.0 + r;
    const double R2 = fast_rcp(opr * u);
    const double z = 0.5 * w0 * (u * R2);                               // zeta_minus, :145
    const double e1 = small ? uq : fma(-(m * opr), R2, 1.0);            // 1 - m/u
    const double zm = z * m, omzm = 1.0 - zm;
    const double chi = -(r + zm) * omzm;                                // :149
    const double xi = (1.0 - z) * zm * (2.0 - m);                       // :150
    const double psi = -r * Tr;                                         // :151
    const double A = fma(2.0, xi, -m * omzm);                           // chi + xi - psi
    const double H = (B1 - B2) * r * fma(zm, 1.0 - e1, e1 - m);         // psi D + B'/(2E)(chi-psi-xi)
    const double ic = fast_rcp(chi);
    F2u = ic * fma(FREI_PI, fma(B2, A, H), fma(psi, F1u, -xi * F2d));   // :161-168
    F1d = ic * fma(FREI_PI, fma(B1, A, -H), fma(psi, F2d, -xi * F1u));  // :169-176
}

__device__ __forceinline__ void two_stream_k(double k, double sg, double dpg, double F1u, double F2d,
                                             double B1, double B2, const double* tab, double& F2u,
                                             double& F1d, double& dtau) {
    double w0, omw;
    two_stream_front(k, sg, dpg, dtau, w0, omw);
    two_stream_tail<false>(dtau, w0, omw, F1u, F2d, B1, B2, tab, F2u, F1d);
}

// propagate_fluxes() signature: delta_tau and omega_0 given (twostream.py:97-99).
__device__ __forceinline__ void two_stream(double dtau, double w0, double F1u, double F2d,
                                           double B1, double B2, const double* tab, double& F2u,
                                           double& F1d) {
    // any (k, sigma, dpg) with sigma/(sigma+k) = w0 and dpg*k = dtau reproduces the inputs
    double unused;
    two_stream_k(1.0 - w0, w0, dtau / (1.0 - w0), F1u, F2d, B1, B2, tab, F2u, F1d, unused);
}

// ---------------------------------------------------------------------------
// debug: expose the fp64 building blocks for accuracy tests
// ---------------------------------------------------------------------------
__global__ void debug_math_kernel(const double* __restrict__ x, double* __restrict__ out, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();
    const double v = x[j < n ? j : n - 1];
    double rs, T, m;
    if (j >= n) return;
    out[j] = fast_rcp(v);
    out[n + j] = fast_sqrt(v, rs);
    out[2 * n + j] = rs;
    exp_neg(v, tab, T, m);
    out[3 * n + j] = T;
    out[4 * n + j] = m;
}

// ---------------------------------------------------------------------------
// K1 standalone: k and sigma for every level
// ---------------------------------------------------------------------------
template <typename TabT>
__global__ void kappa_kernel(const TabT* __restrict__ tab, const double* __restrict__ sigma,
                             const double* __restrict__ sigma_scale, LayerParams lp,
                             double* __restrict__ k_out, double* __restrict__ sigma_out,
                             int L, int S, int N_T, int64_t n_lam) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= n_lam) return;
    const int64_t li = (int64_t)b * L + i;
    const double sg = sigma[j] * (sigma_scale ? sigma_scale[b] : 1.0);
    double acc = 0.0;
    const double* rec = lp.rec + li * lp.rec8;
    for (int s = 0; s < S; ++s) {
        const double* W = rec + 2 + 4 * s;
        const TabT* r0 = tab + reinterpret_cast<const int64_t*>(rec)[2 + 4 * S + s] + j;
        double v = 0.0;
        v += (double)r0[0] * W[0];
        v += (double)r0[n_lam] * W[1];
        v += (double)r0[(int64_t)N_T * n_lam] * W[2];
        v += (double)r0[(int64_t)(N_T + 1) * n_lam] * W[3];
        acc += v;
    }
    k_out[li * n_lam + j] = acc + sg;                    // opacity.py:269 (k includes sigma)
    if (i == 0) sigma_out[(int64_t)b * n_lam + j] = sg;
}

// ---------------------------------------------------------------------------
// K2 standalone: propagate_fluxes elementwise
// ---------------------------------------------------------------------------
__global__ void propagate_kernel(const double* __restrict__ lam, const double* __restrict__ F1u,
                                 const double* __restrict__ F2d, double T1, double T2,
                                 const double* __restrict__ dtau, const double* __restrict__ w0,
                                 double* __restrict__ F2u, double* __restrict__ F1d, int64_t n) {
    __shared__ double tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double l = lam[j];
    const double c1 = 2.0 * FREI_H * FREI_C * FREI_C / pow(l, 5.0);
    const double B1 = c1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T1));
    const double B2 = c1 / expm1(FREI_H * FREI_C / (l * FREI_KB * T2));
    double a, d;
    two_stream(dtau[j], w0[j], F1u[j], F2d[j], B1, B2, tab, a, d);
    F2u[j] = a; F1d[j] = d;
}

// ---------------------------------------------------------------------------
// K2+K3: the layer sweep
// ---------------------------------------------------------------------------
// V consecutive wavelengths with one (vectorised when V == 2) load/store
template <int V> struct Vec;
template <> struct Vec<1> {
    static __device__ __forceinline__ void ld(const double* p, double* o) { o[0] = *p; }
    static __device__ __forceinline__ void ldg(const double* p, double* o) { o[0] = __ldg(p); }
    static __device__ __forceinline__ void ldg(const float* p, double* o) { o[0] = (double)__ldg(p); }
    static __device__ __forceinline__ void st(double* p, const double* v) { *p = v[0]; }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void ld(const double* p, double* o) {
        const double2 t = *reinterpret_cast<const double2*>(p); o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void ldg(const double* p, double* o) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p)); o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void ldg(const float* p, double* o) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p)); o[0] = (double)t.x; o[1] = (double)t.y;
    }
    static __device__ __forceinline__ void st(double* p, const double* v) {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    }
};

// slot of (species s, corner c) for this thread: stage + (4 s + c) * kRow (bytes)
template <typename TabT, int S_T, int V>
__device__ __forceinline__ void stage_rows(const TabT* __restrict__ tabj, const double* rec, int S,
                                           int64_t n_lam, int64_t rowT, uint32_t stage) {
    const int SS = (S_T > 0) ? S_T : S;
    constexpr int kSlot = V * (int)sizeof(TabT);
    constexpr uint32_t kRow = 32u * kSlot;       // one warp's row of the staging block
    const int64_t* off = reinterpret_cast<const int64_t*>(rec) + 2 + 4 * SS;
#pragma unroll
    for (int s = 0; s < SS; ++s) {
        const TabT* r0 = tabj + off[s];
        cp_async<kSlot>(stage + (4 * s + 0) * kRow, r0);
        cp_async<kSlot>(stage + (4 * s + 1) * kRow, r0 + n_lam);
        cp_async<kSlot>(stage + (4 * s + 2) * kRow, r0 + rowT);
        cp_async<kSlot>(stage + (4 * s + 3) * kRow, r0 + rowT + n_lam);
    }
    cp_async_commit();
}

template <int V> struct SVec;
template <> struct SVec<1> {
    static __device__ __forceinline__ void ld(const double* p, double* o) { o[0] = *p; }
    static __device__ __forceinline__ void ld(const float* p, double* o) { o[0] = (double)*p; }
};
template <> struct SVec<2> {
    static __device__ __forceinline__ void ld(const double* p, double* o) {
        const double2 t = *reinterpret_cast<const double2*>(p); o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void ld(const float* p, double* o) {
        const float2 t = *reinterpret_cast<const float2*>(p); o[0] = (double)t.x; o[1] = (double)t.y;
    }
};

// k[v] = sigma + sum_s sum_corners W[s][c] * staged[s][c][v]   (opacity.py:261-269)
template <typename TabT, int S_T, int V>
__device__ __forceinline__ void gather_smem(const TabT* slot, const double* rec, int S,
                                            const double* sg, double* k) {
    const int SS = (S_T > 0) ? S_T : S;
    constexpr int kRowElems = 32 * V;
    // two accumulation chains (corners 0, 1 starting from sigma; corners 2, 3), joined at the end:
    // 4 S + 1 fp64 instructions with a dependent depth of 2 S + 1 (the reference adds the species
    // left to right, opacity.py:265-269; the difference is rounding in the last place)
    double ka[V], kb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { ka[v] = sg[v]; kb[v] = 0.0; }          // k includes sigma, opacity.py:269
#if SWEEP_GATHER4
    if (S_T >= SWEEP_WIDE_S && sizeof(TabT) == 8) {      // the instantiations compiled for SWEEP_MINB_WIDE CTAs/SM
        // many species: one chain per corner (depth S + 2) and all loads of the level in flight
        double kc[V], kd[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { kc[v] = 0.0; kd[v] = 0.0; }
#pragma unroll
        for (int s = 0; s < SS; ++s) {
            const double2 wa = *reinterpret_cast<const double2*>(rec + 2 + 4 * s);
            const double2 wb = *reinterpret_cast<const double2*>(rec + 4 + 4 * s);
            double t0[V], t1[V], t2[V], t3[V];
            SVec<V>::ld(slot + (4 * s + 0) * kRowElems, t0);
            SVec<V>::ld(slot + (4 * s + 1) * kRowElems, t1);
            SVec<V>::ld(slot + (4 * s + 2) * kRowElems, t2);
            SVec<V>::ld(slot + (4 * s + 3) * kRowElems, t3);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                ka[v] = fma(t0[v], wa.x, ka[v]);
                kb[v] = fma(t1[v], wa.y, kb[v]);
                kc[v] = fma(t2[v], wb.x, kc[v]);
                kd[v] = fma(t3[v], wb.y, kd[v]);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) k[v] = (ka[v] + kb[v]) + (kc[v] + kd[v]);
        return;
    }
#endif
#pragma unroll 4
    for (int s = 0; s < SS; ++s) {
        const double2 wa = *reinterpret_cast<const double2*>(rec + 2 + 4 * s);
        const double2 wb = *reinterpret_cast<const double2*>(rec + 4 + 4 * s);
        double t0[V], t1[V], t2[V], t3[V];
        SVec<V>::ld(slot + (4 * s + 0) * kRowElems, t0);
        SVec<V>::ld(slot + (4 * s + 1) * kRowElems, t1);
        SVec<V>::ld(slot + (4 * s + 2) * kRowElems, t2);
        SVec<V>::ld(slot + (4 * s + 3) * kRowElems, t3);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            ka[v] = fma(t0[v], wa.x, ka[v]);
            ka[v] = fma(t1[v], wa.y, ka[v]);
            kb[v] = fma(t2[v], wb.x, kb[v]);
            kb[v] = fma(t3[v], wb.y, kb[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) k[v] = ka[v] + kb[v];
}

// Per-thread state of the sweep: V wavelengths.
template <int V>
struct Lane {
    double c1[V], c2[V], sg[V], wj[V];
    double Fcar[V], Bcar[V];     // carried stream (F_up for emit, F_down for absorb) and Planck term
    int thr[V];                  // high word of 9 sigma (1 + 2^-18): k above it has omega0 < 0.1 for certain
};

// One layer-step for V wavelengths after the warp vote: Planck term of the new level, the
// two-stream response (general form or the E = 1 form), the four wavelength-integral contributions.
template <int DIR, int V, bool SAME_T, bool E_IS_ONE>
__device__ __forceinline__ void layer_tail(Lane<V>& t, const double* k, double dpg, const double* other,
                                           double invTn, const double* tab, double* F2u, double* F1d,
                                           double* dtau, double* red) {
    red[0] = red[1] = red[2] = red[3] = 0.0;
    const double dpg2 = dpg + dpg;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const double Bn = SAME_T ? t.Bcar[v] : planck(t.c1[v], t.c2[v], invTn, tab);
        dtau[v] = dpg * k[v];                                            // :371-373 (dead unless DTAUS)
        // emit: carried = F_1_up, B_1; other = F_2_down; new B = B_2
        // absorb: carried = F_2_down, B_2; other = F_1_up; new B = B_1
        const double F1u = (DIR == FREI_EMIT) ? t.Fcar[v] : other[v];
        const double F2d = (DIR == FREI_EMIT) ? other[v] : t.Fcar[v];
        const double B1 = (DIR == FREI_EMIT) ? t.Bcar[v] : Bn;
        const double B2 = (DIR == FREI_EMIT) ? Bn : t.Bcar[v];
        if (E_IS_ONE) {
            two_stream_E1(k[v], t.sg[v], dpg2, F1u, F2d, B1, B2, tab, F2u[v], F1d[v]);
        } else {
            double w0, omw, dt;
            two_stream_front(k[v], t.sg[v], dpg, dt, w0, omw);
            two_stream_tail<false>(dt, w0, omw, F1u, F2d, B1, B2, tab, F2u[v], F1d[v]);
        }
        red[0] = fma(t.wj[v], F2u[v], red[0]);
        red[1] = fma(t.wj[v], F2d, red[1]);
        red[2] = fma(t.wj[v], F1u, red[2]);
        red[3] = fma(t.wj[v], F1d[v], red[3]);
        t.Fcar[v] = (DIR == FREI_EMIT) ? F2u[v] : F1d[v];
        t.Bcar[v] = Bn;
    }
}

// One layer-step for V wavelengths.  `other` = the stale stream entering the layer, `invTn` =
// 1/T of the level whose Planck term is new this step (ignored when SAME_T: emit's top
// pseudo-layer has T_2 = T_1, twostream.py:358-363).  Writes the two outgoing streams, the four
// wavelength-integral contributions of this thread, and delta_tau.  The warp votes on
// omega0 > 0.1 before any division — omega0 = sigma / (sigma + k) > 0.1  <=>  k < 9 sigma, tested
// on the high words against the per-wavelength threshold t.thr (a superset: the general form is
// correct for every omega0, so a false positive only costs time) — and takes the E = 1 form when
// no lane needs the general one.
template <int DIR, int V, bool SAME_T>
__device__ __forceinline__ void layer_step(Lane<V>& t, const double* k, double dpg, const double* other,
                                           double invTn, const double* tab, double* F2u, double* F1d,
                                           double* dtau, double* red) {
#if SWEEP_E_VOTE
    bool hi = false;
#pragma unroll
    for (int v = 0; v < V; ++v) hi = hi || (__double2hiint(k[v]) <= t.thr[v]);
    if (__any_sync(0xffffffffu, hi))
        layer_tail<DIR, V, SAME_T, false>(t, k, dpg, other, invTn, tab, F2u, F1d, dtau, red);
    else
        layer_tail<DIR, V, SAME_T, true>(t, k, dpg, other, invTn, tab, F2u, F1d, dtau, red);
#else
    layer_tail<DIR, V, SAME_T, false>(t, k, dpg, other, invTn, tab, F2u, F1d, dtau, red);
#endif
}

// A warp-chunk = 32 * V consecutive wavelengths of one atmosphere, all layers.  Each thread carries
// the running stream (F_up for emit, F_down for absorb) and the Planck term of the shared level in
// registers; the other stream is read stale from HBM one layer ahead of its use and both are
// written back.  The layer loop body is a single basic block (no data-dependent or uniform
// branches) and there is no CTA barrier: warps run their chunks independently.
// `jbase` = first wavelength of the chunk, `part` = its [L][4] row of wavelength-integral partials.
// A chunk cut by the relay plan is handed from the warp that ran its first layer-steps to the warp that
// runs the rest through the flux arrays themselves (the carried stream of step s is the row step
// s - 1 stored) and one flag per chunk: stores -> warp barrier -> st.release by lane 0 on one side,
// ld.acquire spin by lane 0 -> warp barrier on the other.  The receiver puts the flag back to 0, so a
// launch leaves the flags as it found them (CUDA graphs replay the same arguments).  The sender is
// always a warp with a lower index in the same or the previous CTA and runs its piece first; the
// receiver runs its piece last, at least relay_quota - (L - 1) layer-steps later.  A wait that runs
// out (a lost sender: never seen) poisons the chunk's integrals so that the solve fails loudly.
__device__ __forceinline__ bool relay_wait(unsigned int* flag, int lane) {
    unsigned int ok = 1;
    if (lane == 0) {
        unsigned int v = 0;
#pragma unroll 1
        for (int spin = 0; spin < (1 << 22); ++spin) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v) break;
            __nanosleep(64);
        }
        ok = v;
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void relay_signal(unsigned int* flag, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
}

// PIECES = false: all L - 1 layer-steps of the chunk (s0, s1, flag ignored); true: steps [s0, s1) in
// visiting order (emit: level 1 + s, absorb: level L - 2 - s), `flag` = the chunk's relay flag.
template <typename TabT, int S_T, int DIR, int V, bool DTAUS, bool PIECES>
__device__ __forceinline__ void sweep_chunk(const SweepArgs& a, int b, int64_t jbase, double* part,
                                            const double* sm_rec, const void* sm_rows, const double* tab,
                                            double sscale, double fscale, int s0, int s1,
                                            unsigned int* flag) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int L = a.L, S = a.S, rec8 = a.lp.rec8;
    if (!PIECES) { s0 = 0; s1 = L - 1; }
    const int SS = (S_T > 0) ? S_T : S;
    const int64_t n_lam = a.n_lam;
    const int64_t rowT = (int64_t)a.N_T * n_lam;
    // staged table elements: a private [4 S][32][V] block per warp (the warps of a CTA may run
    // chunks of different widths at the same time), sized for V = 2
    const TabT* slot = reinterpret_cast<const TabT*>(sm_rows) + (size_t)(tid >> 5) * (4 * SS * 64) + lane * V;
    const uint32_t stage = smem_u32(slot);
    // ---- per-wavelength constants ----
    const int64_t j_raw = jbase + (int64_t)lane * V;
    const bool live = j_raw < n_lam;             // V == 2 only for even n_lam, so all V lanes are in range
    const int64_t j = live ? j_raw : n_lam - V;
    const TabT* tabj = static_cast<const TabT*>(a.tab) + j;
    double* Fu = static_cast<double*>(a.F_up) + (int64_t)b * L * n_lam + j;
    double* Fd = static_cast<double*>(a.F_down) + (int64_t)b * L * n_lam + j;
    double* dt_out = DTAUS ? static_cast<double*>(a.dtaus) + (int64_t)b * L * n_lam + j : nullptr;
    Lane<V> t;
    Vec<V>::ldg(a.c1 + j, t.c1);
    Vec<V>::ldg(a.c2 + j, t.c2);
    Vec<V>::ldg(a.sigma + j, t.sg);
    Vec<V>::ldg(a.w + j, t.wj);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        t.sg[v] *= sscale;
        if (!live) t.wj[v] = 0.0;
        t.thr[v] = __double2hiint(t.sg[v] * 9.00003433227539062) + 1;    // 9 (1 + 2^-18)
    }
    if (DTAUS && live && s0 == 0) {              // leading row of ones, twostream.py:352/:487
        double one[V];
#pragma unroll
        for (int v = 0; v < V; ++v) one[v] = 1.0;
        Vec<V>::st(dt_out, one);
    }

    double F2u[V], F1d[V], dtau[V], red[4], oth[V], nxt[V], k[V];
    const int64_t pstride = part_level_stride(a.rows);       // from one level of this chunk's partials to the next
    auto publish = [&](const double* r, int row) {
        const double r4 = warp_reduce4(r[0], r[1], r[2], r[3], lane);
        if ((lane & 7) == 0) part[row * pstride + (lane >> 3)] = r4;
    };
    if (DIR == FREI_EMIT) {
        // i = 1 .. L-2 regular (other = fluxes_down[i+1], stale), i = L-1 top pseudo-layer
        const int i0 = 1 + s0, iE = min(1 + s1, L - 1);                  // regular layers [i0, iE) of this piece
        const double* rec = sm_rec + (size_t)i0 * rec8;
        const double* pFd = Fd + (int64_t)(i0 + 1) * n_lam;              // fluxes_down[i + 1]
        if (i0 < L - 1) {
            Vec<V>::ld(pFd, nxt);
        }
        stage_rows<TabT, S_T, V>(tabj, rec, S, n_lam, rowT, stage);      // commits the group
        if (PIECES && s0 > 0 && !relay_wait(flag, lane)) {               // lost sender: NaN integrals, the solve fails loudly
#pragma unroll
            for (int v = 0; v < V; ++v) t.wj[v] = __longlong_as_double(0x7ff8000000000000LL);
        }
        Vec<V>::ld(Fu + (int64_t)i0 * n_lam, t.Fcar);                    // fluxes_up[i0]: stale (i0 = 1) or handed over
        const double invT1 = rec[1];
#pragma unroll
        for (int v = 0; v < V; ++v) t.Bcar[v] = planck(t.c1[v], t.c2[v], invT1, tab);
        double* pFu_out = Fu + (int64_t)(i0 + 1) * n_lam;                // fluxes_up[i + 1]
        double* pFd_out = Fd + (int64_t)i0 * n_lam;                      // fluxes_down[i]
        double* pdt = DTAUS ? dt_out + (int64_t)i0 * n_lam : nullptr;
        cp_async_wait_all();
        gather_smem<TabT, S_T, V>(slot, rec, S, t.sg, k);
        for (int i = i0; i < iE; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];                 // fluxes_down[i + 1]
            pFd += n_lam;
            if (i + 1 < L - 1) Vec<V>::ld(pFd, nxt);                     // one layer ahead
            if (!(reinterpret_cast<const int64_t*>(rec + rec8)[2 + 5 * SS] & 1))  // level i + 1: new cell
                stage_rows<TabT, S_T, V>(tabj, rec + rec8, S, n_lam, rowT, stage);
            layer_step<FREI_EMIT, V, false>(t, k, rec[0], oth, rec[rec8 + 1], tab, F2u, F1d, dtau, red);
            if (live) {
                Vec<V>::st(pFu_out, F2u);                                // :392-394
                Vec<V>::st(pFd_out, F1d);
                if (DTAUS) Vec<V>::st(pdt, dtau);
            }
            publish(red, i);
            pFu_out += n_lam; pFd_out += n_lam; rec += rec8;
            if (DTAUS) pdt += n_lam;
            cp_async_wait_all();
            gather_smem<TabT, S_T, V>(slot, rec, S, t.sg, k);
        }
        if (!PIECES || s1 == L - 1) {
            // top: p_2 extrapolated (in the record), T_2 = T_1, F_2_down = F_TOA, F_2_up discarded
            Vec<V>::ldg(a.f_toa + j, oth);                               // :379-382
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] *= fscale;
            layer_step<FREI_EMIT, V, true>(t, k, rec[0], oth, 0.0, tab, F2u, F1d, dtau, red);
            if (live) {
                Vec<V>::st(pFd_out, F1d);
                if (DTAUS) Vec<V>::st(pdt, dtau);
            }
            publish(red, L - 1);
        }
        if (s0 == 0 && lane < 4) part[lane] = 0.0;         // level 0 is not visited
    } else {
        const int i0 = L - 2 - s0, iE = L - 1 - s1;                      // layers i0 down to iE of this piece
        const double* rec = sm_rec + (size_t)i0 * rec8;
        const double* pFu = Fu + (int64_t)i0 * n_lam;                    // fluxes_up[i], stale
        Vec<V>::ld(pFu, nxt);
        stage_rows<TabT, S_T, V>(tabj, rec, S, n_lam, rowT, stage);      // commits the group
        if (PIECES && s0 > 0 && !relay_wait(flag, lane)) {
#pragma unroll
            for (int v = 0; v < V; ++v) t.wj[v] = __longlong_as_double(0x7ff8000000000000LL);
        }
        Vec<V>::ld(Fd + (int64_t)(i0 + 1) * n_lam, t.Fcar);              // fluxes_down[i0 + 1]: stale (top) or handed over
        const double invTt = rec[rec8 + 1];
#pragma unroll
        for (int v = 0; v < V; ++v) t.Bcar[v] = planck(t.c1[v], t.c2[v], invTt, tab);
        double* pFu_out = Fu + (int64_t)(i0 + 1) * n_lam;                // fluxes_up[i + 1]
        double* pFd_out = Fd + (int64_t)i0 * n_lam;                      // fluxes_down[i]
        double* pdt = DTAUS ? dt_out + (int64_t)(1 + s0) * n_lam : nullptr;   // visiting order
        cp_async_wait_all();
        gather_smem<TabT, S_T, V>(slot, rec, S, t.sg, k);
        for (int i = i0; i >= iE; --i) {
#pragma unroll
            for (int v = 0; v < V; ++v) oth[v] = nxt[v];                 // fluxes_up[i], :512
            pFu -= n_lam;
            if (i > 0) Vec<V>::ld(pFu, nxt);                             // one layer ahead
            if (i > 0 && !(reinterpret_cast<const int64_t*>(rec)[2 + 5 * SS] & 1))   // level i - 1: new cell
                stage_rows<TabT, S_T, V>(tabj, rec - rec8, S, n_lam, rowT, stage);
            layer_step<FREI_ABSORB, V, false>(t, k, rec[0], oth, rec[1], tab, F2u, F1d, dtau, red);
            if (live) {
                Vec<V>::st(pFu_out, F2u);                                // :521-522
                Vec<V>::st(pFd_out, F1d);
                if (DTAUS) Vec<V>::st(pdt, dtau);
            }
            publish(red, i);
            pFu_out -= n_lam; pFd_out -= n_lam;
            if (DTAUS) pdt += n_lam;
            if (i > 0) {
                rec -= rec8;
                cp_async_wait_all();
                gather_smem<TabT, S_T, V>(slot, rec, S, t.sg, k);
            }
        }
        if (s0 == 0 && lane < 4) part[(L - 1) * pstride + lane] = 0.0;    // level L-1 is not visited
    }
    if (PIECES && s1 < L - 1) relay_signal(flag, lane);    // the rest of the chunk belongs to another warp
}

// The sweep kernel.  The wavelength axis is cut into a.n2 chunks of 64 wavelengths (two per thread)
// followed by a.rows - a.n2 chunks of 32 (one per thread); chunk q owns row q of the partials.  For
// a single atmosphere the grid holds at most one resident wave of CTAs (host: occupancy x SM
// count) and warp w of CTA c takes the chunks q = w G + c, + kWarps G, ... (G = gridDim.x), so
// every round of chunks is spread evenly over the SMs.  A chunk is a serial recurrence over the
// layers: a launch proceeds in rounds of (resident warps) chunks and a partly filled last round
// costs a whole chunk latency.  The host plan (sweep_plan) therefore fills the complete rounds with
// 64-wide chunks and cuts what is left into 32-wide ones — twice as many warps, each with half the
// instructions per layer — which shortens the tail of the launch (C2: 1.32 rounds of 64-wide chunks
// cost 1.59 round times, one round + a 64 % full round of 32-wide chunks 1.38).
// The level records are staged into shared memory by one TMA bulk copy per CTA.
template <typename TabT, int S_T, int DIR, bool DTAUS, bool RELAY>
__global__ void __launch_bounds__(kThreads, (S_T >= SWEEP_WIDE_S && sizeof(TabT) == 8) ? SWEEP_MINB_WIDE :
                                            (DIR == FREI_EMIT) ? SWEEP_MINB_EMIT : SWEEP_MINB)
sweep_kernel(SweepArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double tab[32];                   // 2^(j/32) for exp_neg
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y;
    constexpr int SB = (DIR == FREI_EMIT) ? 0 : 8;                   // probe: absorb sweeps stamp 8 slots higher
    if (tid == 0) { STAMP_MIN(14 + SB); }                            // sweep CTA resident (before the wait)
    pdl_wait();                                  // records, T, active flags come from the previous kernel
    if (tid == 0) { STAMP_MIN(16 + SB); STAMP_MAX(17 + SB); }        // sweep released
    // the row count for the post kernel, which reads it BEFORE its griddepcontrol.wait: written and fenced
    // before this CTA lets the dependent kernel be scheduled (it starts when every CTA has done so)
    if (a.plan_hdr && blockIdx.x == 0 && b == 0 && tid == 0) { a.plan_hdr[0] = a.rows; __threadfence(); }
    if (gridDim.y == 1) pdl_launch_dependents(); // one resident wave: the post kernel may queue up behind it
    // converged atmosphere of a batch: nothing to do.  A single tracked atmosphere (Grid.emission_spectrum)
    // consumes the flag only after the records have arrived, so that its load overlaps theirs
    const unsigned act = a.active ? a.active[b] : 1u;
    if (gridDim.y > 1 && !act) return;
    const int L = a.L, rec8 = a.lp.rec8;
    if (tid < 32) tab[tid] = kExp2Tab[tid];
    const double sscale = a.sigma_scale ? a.sigma_scale[b] : 1.0;
    const double fscale = a.ftoa_scale ? a.ftoa_scale[b] : 1.0;
    double* sm_rec = smem;                       // [L][rec8]
    const void* sm_rows = smem + (size_t)L * rec8;
    // ---- stage the level records (TMA bulk copy global -> shared, mbarrier completion) ----
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
    {   // ---- wait for the records ----
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
    }
    if (!act) return;
#if POST_PER_LEVEL
    // "same table rows as the level below" (bit 0 of the record's flag word): the kernel that wrote
    // the records works one level per CTA and cannot see the neighbour's new bracket, so the flags
    // are made here, from the row offsets of the staged records
    for (int lev = tid; lev < L; lev += kThreads) {
        int64_t* r1 = reinterpret_cast<int64_t*>(sm_rec + (size_t)lev * rec8);
        int64_t same = lev > 0;
        if (lev > 0)
            for (int s = 0; s < a.S; ++s) if (r1[2 + 4 * a.S + s] != (r1 - rec8)[2 + 4 * a.S + s]) same = 0;
        r1[2 + 5 * a.S] = same;
    }
    __syncthreads();
#endif
    if (tid == 0) { STAMP_MIN(18 + SB); STAMP_MAX(19 + SB); }        // records in shared memory

    const int G = gridDim.x, NS = L - 1;
    if (!RELAY) {
        // whole chunks q = w G + c, + kWarps G, ...
        for (int q = warp * G + blockIdx.x; q < a.rows; q += kWarps * G) {
            double* part = part_base(a.partials, b, a.rows, L, q);
            if (q < a.n2)
                sweep_chunk<TabT, S_T, DIR, 2, DTAUS, false>(a, b, (int64_t)q * 64, part, sm_rec, sm_rows, tab, sscale,
                                                             fscale, 0, NS, nullptr);
            else
                sweep_chunk<TabT, S_T, DIR, 1, DTAUS, false>(a, b, (int64_t)a.n2 * 64 + (int64_t)(q - a.n2) * 32, part,
                                                             sm_rec, sm_rows, tab, sscale, fscale, 0, NS, nullptr);
        }
        return;
    }
    // Relay plan (all chunks 64 wide, more chunks than resident warps): the rows * (L - 1) layer-steps,
    // chunk after chunk, are cut into runs of relay_quota steps; warp m = 4 c + w owns run m and works
    // through it BACKWARDS: first the leading steps of the chunk that the end of its run cuts (nothing to
    // wait for), then whole chunks, last the trailing steps of the chunk that the start of its run cuts —
    // handed over by warp m - 1, which ran the leading steps of that chunk first of all.  Every warp is
    // busy for the same number of steps, so the launch has no partly filled last round; every (chunk,
    // layer) is computed by the same instructions as in a whole chunk, so the results are bit-identical.
    const int W = a.rows * NS;                   // < 2^30 (checked by the launcher)
    const int m = blockIdx.x * kWarps + warp;
    const int p0 = min(m, W / a.relay_quota + 1) * a.relay_quota;      // my run: steps [p0, p) of the step line
    for (int p = min(p0 + a.relay_quota, W); p > p0;) {
        const int q = (p - 1) / NS, base = q * NS;
        const int lo = max(p0, base) - base, hi = p - base;
        p = base + lo;
        double* part = part_base(a.partials, 0, a.rows, L, q);
        if (lo == 0 && hi == NS)
            sweep_chunk<TabT, S_T, DIR, 2, DTAUS, false>(a, b, (int64_t)q * 64, part, sm_rec, sm_rows, tab, sscale,
                                                         fscale, 0, NS, nullptr);
        else
            sweep_chunk<TabT, S_T, DIR, 2, DTAUS, true>(a, b, (int64_t)q * 64, part, sm_rec, sm_rows, tab, sscale,
                                                        fscale, lo, hi, a.relay_flags + q);
    }
    if ((tid & 31) == 0) { STAMP_MIN(20 + SB); STAMP_MAX(21 + SB); } // warp done
}

// ---------------------------------------------------------------------------
// K4: per-layer thermodynamics and the temperature update
// ---------------------------------------------------------------------------
__device__ __forceinline__ double cp_of(double m_bar) { return (2.0 + 5.0) / (2.0 * m_bar) * FREI_KB; }   // :220-224

struct UpdateArgs {
    double* T; const double* P; const double* g; const double* m_bar; const double* alpha;
    double* dT; double* T_hist;
    int L, direction;
    double alpha_override;
    // batch convergence (core.py:301-318), all nullable
    uint8_t* active;            // [B] in/out
    double* trk_T;              // [B][L] temperature in the previous history column
    int32_t* trk_state;         // [B][L][2] sign of the previous difference (2 = none yet), sign flips
    int32_t* trk_ncol;          // [B] history columns so far
    int32_t* trk_iters;         // [B] emit+absorb iterations done when the atmosphere converged
    int n_zero_crossings;
    double convergence_dT;
};

// dT of level i of atmosphere b from its four wavelength integrals s[0..3]
// (div_bol_net_flux, convective_flux, delta_t_i, delta_temperature; twostream.py:23-43, 190-287)
// Pressure-only part of delta_T_level: the same in every sweep of a solve.  The post kernel computes it
// before it waits for the sweep (thread i = level i), so that the loads of g, m_bar and alpha, one of
// the two logarithms and half of the divisions are off the critical path between two sweeps.
struct LevelPre {
    double dpg;      // (p1 - p2) / g                                   twostream.py:238
    double kz;       // k_B / (m_bar g) log(p1 / p2):  dz = kz T_1      :186-187
    double kz_def;   // the same with the default mean mass 2.4 m_p     :403-405
    double gcp;      // g / c_p                                         :241-266
    double lmix_c;   // alpha k_B / (m_bar g):  mixing length = lmix_c T_1   :270
    double dtr_c;    // c_p p1 / sigma_SB / g:  dt_rad = dtr_c / T_1^3  :37
    double g, cp;
    double rdpg;     // g / (p1 - p2)
};
__device__ __forceinline__ LevelPre level_pre(const UpdateArgs& u, int b, int i, const double* Pb) {
    const int L = u.L;
    LevelPre q;
    const double g = u.g[b], m_bar = u.m_bar[b];
    const double alpha = (u.alpha_override >= 0.0) ? u.alpha_override : u.alpha[b];
    const double p1 = Pb[i] * FREI_BAR;
    double p2;
    if (i == L - 1) p2 = p1 * (Pb[L - 2] * FREI_BAR) / (Pb[L - 3] * FREI_BAR);   // :358-363
    else p2 = Pb[i + 1] * FREI_BAR;
    const double lg = log(p1 / p2);
    q.g = g; q.cp = cp_of(m_bar);
    q.dpg = (p1 - p2) / g;
    q.kz = FREI_KB / (m_bar * g) * lg;
    q.kz_def = FREI_KB / (2.4 * FREI_MP * g) * lg;
    q.gcp = g / q.cp;
    q.lmix_c = alpha * FREI_KB / (m_bar * g);
    q.dtr_c = q.cp * p1 / FREI_SIGSB / g;
    q.rdpg = g / (p1 - p2);
    return q;
}

// dT of a level from its temperatures and dF_rad with the IEEE division, sqrt, log and exp of the
// CUDA math library: the path of every input the short forms below do not cover (non-positive or
// non-finite temperatures of an atmosphere whose explicit update is diverging, |X| outside
// 1e-290 .. 1e290) — NaN and infinity propagate exactly as in the reference's numpy expressions.
__device__ __noinline__ double delta_T_general(const LevelPre& q, double T1, double T2, double dF_rad) {
    // This chain runs in one warp after the last ticket of the reduction, i.e. it is pure latency
    // on the critical path of every sweep: x^1.5 = x sqrt(x) and |X|^0.9 = exp(0.9 log|X|) replace
    // the two pow() calls (~250 dependent instructions each); the results differ from pow() by
    // < 1e-14 relative, against a parity tolerance of 1e-7 on dT.
    const double dz = q.kz * T1;                                              // :186-187
    const double rho = q.dpg / dz;                                            // :238
    const double dgam = (T1 - T2) / dz - q.gcp;                               // :241-266
    const double lmix = q.lmix_c * T1;                                        // :270
    double F_conv = 0.0;
    if (dgam > 0.0) F_conv = rho * q.cp * (lmix * lmix) * sqrt(q.g / T1) * (dgam * sqrt(dgam));   // :285-287
    const double div = (dF_rad + F_conv) / dz;                                // :205
    const double X = div * dz;
    const double f_pre = (X != 0.0) ? 1e5 * exp(-0.9 * log(fabs(X))) : 1.0;   // :32-35
    const double dt_rad = q.dtr_c / (T1 * T1 * T1);                           // :37
    double dt = f_pre * dt_rad;
    if (dgam > 0.0) dt = f_pre * fmin(dt_rad, sqrt(T1 / q.g / dgam));         // :39-43
    // delta_temperature is called without m_bar: defaults 2.4 m_p, n_dof 5 (:403-405)
    const double rho_def = q.dpg / (q.kz_def * T1);
    return 1.0 / rho_def / cp_of(2.4 * FREI_MP) * div * dt;                   // :216-217
}

// log(x) for normal positive finite x: x = m 2^e with m in [sqrt(1/2), sqrt 2), log m = 2 atanh(s),
// s = (m - 1)/(m + 1), |s| < 0.172; series to s^17 (truncation 4e-15 relative), ~25 instructions inline.
__device__ __forceinline__ double fast_log(double x) {
    int hi = __double2hiint(x);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));      // [1, 2)
    const bool big = m > 1.4142135623730951;
    m = big ? 0.5 * m : m;
    e += big ? 1 : 0;
    const double sn = (m - 1.0) * fast_rcp(m + 1.0), s2 = sn * sn;
    double p = fma(s2, 2.0 / 17.0, 2.0 / 15.0);
    p = fma(p, s2, 2.0 / 13.0);
    p = fma(p, s2, 2.0 / 11.0);
    p = fma(p, s2, 2.0 / 9.0);
    p = fma(p, s2, 2.0 / 7.0);
    p = fma(p, s2, 2.0 / 5.0);
    p = fma(p, s2, 2.0 / 3.0);
    const double lm = fma(sn * s2, p, sn + sn);
    const double de = (double)e;
    return fma(de, 0.6931471803691238, fma(de, 1.9082149292705877e-10, lm));    // ln 2 in two parts
}

// dT of level i of atmosphere b from its four wavelength integrals s[0..3]
// (div_bol_net_flux, convective_flux, delta_t_i, delta_temperature; twostream.py:23-43, 190-287).
// This chain runs in one warp after the last ticket of the reduction — pure latency between two
// sweeps, and cold straight-line code: with the library's division / sqrt / log / exp it took 3.8 us
// (time stamps, scripts/stamp_probe.py), mostly instruction fetch.  The short forms use the
// kernels' own reciprocal, square root and exponential (<= 1.5 ulp each) and a series logarithm;
// the result differs from the library path by ~1e-15 relative, against a parity tolerance of 1e-7.
// T1 = T of the level, T2 = T of the level above (T1 itself for the top level, :358-363)
__device__ __forceinline__ double delta_T_level(const UpdateArgs& u, const LevelPre& q, int i, const double* s,
                                                double T1, double T2) {
    const int L = u.L;
    const bool active = (u.direction == FREI_EMIT) ? (i >= 1) : (i <= L - 2);
    if (!active) return 0.0;                                  // dT[0] = 0 (emit), dT[L-1] = 0 (absorb)
    const double dF_rad = (s[0] - s[1]) - (s[2] - s[3]);                      // :199
    const double dz = q.kz * T1;                                              // :186-187
    const double rdz = fast_rcp(dz), rT1 = fast_rcp(T1);
    const double rho = q.dpg * rdz;                                           // :238
    const double dgam = (T1 - T2) * rdz - q.gcp;                              // :241-266
    const double lmix = q.lmix_c * T1;                                        // :270
    const bool conv = dgam > 0.0;
    double rs, F_conv = 0.0, dt_conv = 0.0;
    const double dg = conv ? dgam : 1.0;                                      // keeps the unused branch finite
    const double sq_dg = fast_sqrt(dg, rs);
    if (conv) {
        double rs2;
        F_conv = rho * q.cp * (lmix * lmix) * fast_sqrt(q.g * rT1, rs2) * (dgam * sq_dg);    // :285-287
        dt_conv = fast_sqrt(T1 * fast_rcp(q.g), rs2) * rs;                    // sqrt(T1 / g / dgam), :41
    }
    const double div = (dF_rad + F_conv) * rdz;                               // :205
    const double X = div * dz, aX = fabs(X);
    double f_pre = 1.0;                                                       // :32-35
    if (X != 0.0) {
        const double w = 0.9 * fast_log(aX);                                  // |X|^-0.9 = exp(-0.9 log |X|)
        double ew, mw;
        exp_neg(fabs(w), kExp2Tab, ew, mw);
        f_pre = 1e5 * (w >= 0.0 ? ew : fast_rcp(ew));
    }
    const double dt_rad = q.dtr_c * (rT1 * rT1 * rT1);                        // :37
    const double dt = f_pre * (conv ? fmin(dt_rad, dt_conv) : dt_rad);        // :39-43
    // delta_temperature is called without m_bar: defaults 2.4 m_p, n_dof 5 (:403-405)
    const double res = (q.kz_def * T1 * q.rdpg) * (1.0 / cp_of(2.4 * FREI_MP)) * div * dt;   // :216-217
    const bool plain = T1 > 1e-3 && T1 < 1e9 && T2 > 1e-3 && T2 < 1e9 && (X == 0.0 || (aX > 1e-290 && aX < 1e290)) &&
                       res == res && fabs(res) < 1e300;
    return plain ? res : delta_T_general(q, T1, T2, dF_rad);
}

// Block-wide: thread i = level i.  All reads of T precede the barrier, all writes follow it;
// then (optionally) the records of the new T are rebuilt for the next sweep.  `lv` = where T, P
// and the mixing ratios of this atmosphere are read; `sm_T` (nullable) = a writable shared-memory
// copy of T that lv.T points to: the new T is stored there as well, so K0 does not wait for a
// global-memory round trip of what this block has just computed.
// `pre` (nullable) = the pressure-only terms of this thread's level, `has_T` (nullable) = a shared
// memory copy of the species' has_T flags: both prepared by the caller ahead of time.
__device__ __forceinline__ void update_and_prep(const UpdateArgs& u, const PrepArgs& pa, int do_prep,
                                                int b, const double* sums_b, const double* sm_axes,
                                                const LevelView& lv, double* sm_T, const LevelPre* pre,
                                                const int32_t* has_T) {
    const int i = threadIdx.x, L = u.L;
    double dT = 0.0, T1 = 0.0;
    // tracker state of this level, loaded ahead of the thermodynamics it does not depend on
    double trk_T = 0.0;
    int trk_sgn = 2, trk_flips = 0, trk_ncol = 0;
    if (u.trk_T) {
        trk_ncol = u.trk_ncol[b];
        if (i < L) {
            const int64_t li = (int64_t)b * L + i;
            trk_T = u.trk_T[li];
            trk_sgn = u.trk_state[li * 2]; trk_flips = u.trk_state[li * 2 + 1];
        }
    }
    if (i < L) {
        T1 = lv.T[i];
        const LevelPre q = pre ? *pre : level_pre(u, b, i, lv.P);
        dT = delta_T_level(u, q, i, sums_b + i * 4, T1, (i == L - 1) ? T1 : lv.T[i + 1]);
    }
    if (threadIdx.x == 0) { STAMP_MAX(23); }                         // probe: dT of level 0 computed (thread 0: inactive level for emit)
    if (threadIdx.x == 1) { STAMP_MAX(25); }                         // probe: dT of level 1 computed
    __syncthreads();
    if (threadIdx.x == 0) { STAMP_MAX(27); }                         // probe: all levels' dT computed
    const double Tn = T1 - dT;                                                // :407, :536
    if (i < L) {
        u.dT[(int64_t)b * L + i] = dT;
        u.T[(int64_t)b * L + i] = Tn;
        if (sm_T) sm_T[i] = Tn;
        if (u.T_hist) u.T_hist[(int64_t)b * L + i] = Tn;
    }
    if (u.trk_T) {
        // Per-layer convergence of Grid.emission_spectrum (core.py:306-311), kept incrementally:
        // every sweep appends a history column; a layer is converged when the successive column
        // differences changed sign more than n_zero_crossings times, or the absorb step is below
        // convergence_dT; the atmosphere stops when all layers are (tested after absorb, :317).
        bool conv = true;
        if (i < L) {
            const int64_t li = (int64_t)b * L + i;
            const int ncol = trk_ncol;
            int sgn_prev = trk_sgn, flips = trk_flips;
            if (ncol > 0) {
                // np.sign(diffs[1:]) != np.sign(diffs[:-1]) (core.py:308): a NaN difference compares
                // unequal to everything, itself included, so an atmosphere whose explicit update has
                // diverged collects a "sign change" per column and is stopped by the rule (code 3)
                const double d = Tn - trk_T;
                const int sgn = (d != d) ? 3 : (d > 0.0) - (d < 0.0);
                if (sgn_prev != 2 && (sgn != sgn_prev || sgn == 3)) ++flips;
                sgn_prev = sgn;
            }
            u.trk_T[li] = Tn;
            u.trk_state[li * 2] = sgn_prev;
            u.trk_state[li * 2 + 1] = flips;
            conv = (flips > u.n_zero_crossings) || (fabs(dT) < u.convergence_dT);
        }
        const int all_conv = __syncthreads_and(conv ? 1 : 0);
        if (i == 0) {
            const int ncol = trk_ncol + 1;
            u.trk_ncol[b] = ncol;
            if (u.direction == FREI_ABSORB && all_conv && u.active) {
                u.active[b] = 0;
                if (u.trk_iters) u.trk_iters[b] = ncol / 2;
            }
        }
    }
    if (threadIdx.x == 0) { STAMP_MAX(11); }                         // temperature update done
    if (!do_prep) return;
    __syncthreads();
    prep_block(pa, b, sm_axes, lv, has_T ? has_T : pa.has_T, pre ? pre->g : pa.g[b], pre ? &pre->dpg : nullptr);
}

__global__ void update_prep_kernel(UpdateArgs u, PrepArgs pa, int do_prep, const double* __restrict__ sums,
                                   int use_smem) {
    extern __shared__ double sm_upd[];
    if (u.active && !u.active[blockIdx.x]) return;
    if (do_prep && use_smem) stage_axes(pa.axis_P, pa.axis_T, pa.S, pa.N_P, pa.N_T, sm_upd);   // synchronised inside
    update_and_prep(u, pa, do_prep, blockIdx.x, sums + (int64_t)blockIdx.x * u.L * 4,
                    (do_prep && use_smem) ? sm_upd : nullptr,
                    global_levels(u.T, u.P, do_prep ? pa.mmr : nullptr, blockIdx.x, u.L, pa.S), nullptr, nullptr, nullptr);
}

// ---------------------------------------------------------------------------
// fixed-order reduction of the per-warp partials (+ optional fused T update and re-bracketing):
// the two-stage kernel of round 1, built with -DPOST_PER_LEVEL=0 (default: post_level_kernel below)
// ---------------------------------------------------------------------------
// grid (nchunks, B).  Stage 1: every CTA sums its chunk of partial rows.  The CTA that finishes
// last for an atmosphere (atomic ticket) sums the chunk results in fixed chunk order — the
// result does not depend on which CTA that is — and, on a single device, goes on to update T
// and rebuild the level records, so one launch follows each sweep.
struct PostArgs {
    const double* partials; double* chunk_sums; unsigned int* counters; double* sums;
    const uint8_t* active;
    // fused cross-GPU sum over NVLink peer memory (null peer_bufs = single device)
    unsigned long long* const* peer_bufs;    // [world] exchange buffers, 8-byte words [2][world][B][L*4][2]
    int* p2p_error;
    unsigned long long epoch;
    int rank, world, B;
    const int32_t* plan_hdr;             // [0] = partial rows written by the sweep before this launch
    int nchunks;
    int do_update, do_prep;
    int axes_smem;                       // the axes of K0 are searched in shared memory
    int stage_levels;                    // T, P, mmr (and the row offsets) of the atmosphere are staged in shared memory
    double* T_next;                      // post_level_kernel: [B][L] new temperatures of this launch
    unsigned int* conv_count;            // post_level_kernel: [B] levels that met the convergence rule
};

// Sum `count` rows of n doubles (row stride n) element-wise with all threads of the CTA:
// thread (g, e) adds rows g, g + G, ... of element e, then the G group results are added in
// group order through shared memory.  Fixed order -> deterministic.  The loads of a thread are
// issued kBatch at a time before the first addition (the additions keep their order): the second
// stage of the reduction reads ~150 rows with 5 groups, and as a chain of dependent
// load-then-add steps it cost ~5 us of L2 latency on the critical path of every sweep.
__device__ __forceinline__ double cta_column_sum(const double* __restrict__ rows, int count, int n,
                                                 double* scratch /* [G][n] */, int G) {
    constexpr int kBatch = 16;
    const int e = threadIdx.x % n, g = threadIdx.x / n;
    double s = 0.0;
    if (g < G) {
        int r = g;
        for (; r + (kBatch - 1) * G < count; r += kBatch * G) {
            double v[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) v[k] = rows[(int64_t)(r + k * G) * n + e];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) s += v[k];
        }
        {
            double v[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) v[k] = (r + k * G < count) ? rows[(int64_t)(r + k * G) * n + e] : 0.0;
#pragma unroll
            for (int k = 0; k < kBatch; ++k) if (r + k * G < count) s += v[k];
        }
        scratch[g * n + e] = s;
    }
    __syncthreads();
    double tot = 0.0;
    if (g == 0)
        for (int k = 0; k < G; ++k) tot += scratch[k * n + e];
    __syncthreads();
    return tot;              // valid in threads with g == 0 (threadIdx.x < n)
}

// Shared memory of post_kernel, in doubles: [G n] scratch, [n] sums, then (do_update) T[L], P[L],
// then (do_prep) mmr[L S], off[L S], axes[S (N_P + N_T)].
__global__ void __launch_bounds__(1024) post_kernel(PostArgs q, UpdateArgs u, PrepArgs pa) {
    extern __shared__ double sm_post[];
    __shared__ int is_last;
    const int b = blockIdx.y, chunk = blockIdx.x, n = u.L * 4, L = u.L;
    const int G = blockDim.x / n;                // row groups per CTA (>= 1, host guarantees)
    double* sm_sums = sm_post + (size_t)G * n;
    const int Lst = q.stage_levels ? L : 0;      // level areas are empty when they do not fit
    double* sm_T = sm_sums + n;
    double* sm_P = sm_T + Lst;
    double* sm_mmr = sm_P + Lst;
    int64_t* sm_off = reinterpret_cast<int64_t*>(sm_mmr + (size_t)Lst * pa.S);
    double* sm_axes = reinterpret_cast<double*>(sm_off + (size_t)Lst * pa.S);
    // Everything the serial tail needs that the sweep does not write — T (last written by the
    // previous post kernel, which completed before the sweep in front of us got past its own
    // griddepcontrol.wait), P, the mixing ratios, the table axes — is copied to shared memory
    // BEFORE waiting for the sweep: with programmatic dependent launch these CTAs are resident
    // while the sweep drains, so the loads overlap its tail instead of following it.  Any CTA may
    // turn out to be the last one of its atmosphere, so all of them do it (a few KB each).
    if (q.stage_levels) {
        if (q.do_update)
            for (int e = threadIdx.x; e < L; e += blockDim.x) {
                sm_T[e] = u.T[(int64_t)b * L + e];
                sm_P[e] = u.P[(int64_t)b * L + e];
            }
        if (q.do_prep)
            for (int e = threadIdx.x; e < L * pa.S; e += blockDim.x) sm_mmr[e] = pa.mmr[(int64_t)b * L * pa.S + e];
    }
    if (threadIdx.x == 0) { STAMP_MIN(0); STAMP_MAX(1); }           // post CTA entry
    if (q.do_prep && q.axes_smem) stage_axes(pa.axis_P, pa.axis_T, pa.S, pa.N_P, pa.N_T, sm_axes);
    // also before the wait: the species' has_T flags, the pressure-only terms of the temperature
    // update (thread i = level i; loads of g, m_bar, alpha, a logarithm and four divisions) and the
    // number of partial rows — the sweep writes it, fenced, before it lets this kernel be scheduled
    __shared__ int32_t sm_hasT[kMaxS];
    if (q.do_prep && threadIdx.x < pa.S) sm_hasT[threadIdx.x] = pa.has_T[threadIdx.x];
    LevelPre pre;
    // (single atmosphere only: in a batch every CTA of every atmosphere would repeat what one CTA per
    // atmosphere needs — C4 lost 3.6 % to it)
    const bool have_pre = q.do_update && q.stage_levels && gridDim.y == 1;
    if (have_pre && threadIdx.x < L) pre = level_pre(u, b, threadIdx.x, u.P + (int64_t)b * L);
    const int rows = __ldcg(q.plan_hdr);
    pdl_wait();                                  // partials come from the sweep before
    if (threadIdx.x == 0) { STAMP_MIN(2); STAMP_MAX(3); }           // after griddepcontrol.wait
    pdl_launch_dependents();                     // the next sweep's CTAs may line up behind the serial tail
    if (q.active && !q.active[b]) return;        // converged atmosphere of a batch
    const int rows_per_chunk = (rows + q.nchunks - 1) / q.nchunks;
    const int r0 = min(rows, chunk * rows_per_chunk), r1 = min(rows, r0 + rows_per_chunk);
    const double* p = q.partials + ((int64_t)b * rows + r0) * n;
    const double s1 = cta_column_sum(p, r1 - r0, n, sm_post, G);
    if (threadIdx.x < n) q.chunk_sums[((int64_t)b * q.nchunks + chunk) * n + threadIdx.x] = s1;
    if (threadIdx.x == 0) { STAMP_MIN(4); STAMP_MAX(5); }           // stage 1 done
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(q.counters + b, 1u) == (unsigned)(q.nchunks - 1));
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) { STAMP_MAX(7); }                          // last ticket taken
    __threadfence();
    double s2 = cta_column_sum(q.chunk_sums + (int64_t)b * q.nchunks * n, q.nchunks, n, sm_post, G);
    if (threadIdx.x == 0) { STAMP_MAX(9); }                          // stage 2 done
    if (q.peer_bufs) {
        // One-shot all-reduce fused into this kernel, flag-in-data ("LL") protocol: every double of
        // this rank's [L][4] integrals is sent to every rank as two 8-byte words, each holding 32
        // payload bits and the 32-bit epoch.  An aligned 8-byte store is single-copy atomic, so a
        // word that shows the current epoch carries valid payload: no fence, no separate flag, no
        // second trip over NVLink — the receiver polls the words themselves.  (A 16-byte vector
        // store is two such words; nothing relies on the pair arriving together.)  The
        // contributions are added in rank order, the same on every rank, so all ranks get
        // bit-identical sums and apply the identical temperature update.  Slots are
        // double-buffered by the parity of the epoch: a rank can be at most one sweep ahead of the
        // slowest one, because its next exchange needs that rank's contribution.
        const int par = (int)(q.epoch & 1ull);
        const unsigned long long fl = (q.epoch & 0xffffffffull) << 32;
        if (threadIdx.x < n) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(s2);
            const unsigned long long w0 = (bits & 0xffffffffull) | fl, w1 = (bits >> 32) | fl;
            const int64_t slot = ((((int64_t)par * q.world + q.rank) * q.B + b) * n + threadIdx.x) * 2;
            for (int r = 0; r < q.world; ++r) {
                const int dst = (q.rank + r) % q.world;      // spread the first stores over the links
                asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};"
                             ::"l"(q.peer_bufs[dst] + slot), "l"(w0), "l"(w1) : "memory");
            }
            double tot = 0.0;
            bool timed_out = false;
            for (int r = 0; r < q.world; ++r) {
                const unsigned long long* src =
                    q.peer_bufs[q.rank] + ((((int64_t)par * q.world + r) * q.B + b) * n + threadIdx.x) * 2;
                unsigned long long a0 = 0, a1 = 0;
                for (long long spins = 0;; ++spins) {
                    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];"
                                 : "=l"(a0), "=l"(a1) : "l"(src) : "memory");
                    if ((a0 & 0xffffffff00000000ull) == fl && (a1 & 0xffffffff00000000ull) == fl) break;
                    if (spins > (1ll << 27)) { timed_out = true; break; }     // seconds: a peer is gone
                }
                tot += __longlong_as_double((long long)((a0 & 0xffffffffull) | (a1 << 32)));
            }
            if (timed_out && q.p2p_error) *q.p2p_error = 1;   // the host raises at its next poll
            s2 = tot;
        }
    }
    if (threadIdx.x < n) {
        q.sums[(int64_t)b * n + threadIdx.x] = s2;
        sm_sums[threadIdx.x] = s2;
    }
    if (threadIdx.x == 0) q.counters[b] = 0u;    // self-cleaning for the next launch
    if (!q.do_update) return;
    __syncthreads();
    LevelView lv;
    if (q.stage_levels) { lv.T = sm_T; lv.P = sm_P; lv.mmr = sm_mmr; lv.off = sm_off; }
    else lv = global_levels(u.T, u.P, q.do_prep ? pa.mmr : nullptr, b, L, pa.S);
    update_and_prep(u, pa, q.do_prep, b, sm_sums, (q.do_prep && q.axes_smem) ? sm_axes : nullptr, lv,
                    q.stage_levels ? sm_T : nullptr, have_pre ? &pre : nullptr, q.do_prep ? sm_hasT : nullptr);
    __syncthreads();
    if (threadIdx.x == 0) { STAMP_MAX(13); }                         // records written
}

// ---------------------------------------------------------------------------
// one CTA per level: reduction, temperature update and re-bracketing without a serial tail
// ---------------------------------------------------------------------------
// grid (L, B).  Everything the sweep leaves for a level — the four wavelength integrals of its
// partial rows, the temperature step (needs only T of the level and of the level above), the new
// brackets and weights of that level — is independent of the other levels, so each level gets a CTA:
// no second reduction stage, no ticket between stages, and the chain "sums -> dT -> brackets ->
// records" runs once per level in parallel instead of for all levels in one CTA (time stamps of the
// round-1 structure, DESIGN.md 3.2: 18 us from the end of the sweep to the records; here the same
// chain is one pass).  The partial rows of a level are contiguous (POST_PER_LEVEL layout), read with
// 32-byte loads and summed in a fixed order: thread-strided over the rows, the warp butterfly of
// warp_reduce4, then the warps in order — bitwise deterministic.  The new T of a level goes to
// T_next: other CTAs still read the old T (as "the level above"); the CTA that takes the last ticket
// of the atmosphere copies T_next to T and closes the convergence tracker.  The same-rows flag of a
// record needs the neighbour's new bracket and is made by the sweep from the staged records.
__global__ void __launch_bounds__(1024) post_level_kernel(PostArgs q, UpdateArgs u, PrepArgs pa) {
    extern __shared__ double sm_pl[];            // [warps][4] | axes of K0
    __shared__ double sm_bc[8];                  // [0..3] sums of the level, [4] new T
    __shared__ int32_t sm_hasT[kMaxS];
    __shared__ int is_last;
    const int i = blockIdx.x, b = blockIdx.y, L = u.L, S = pa.S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int64_t li = (int64_t)b * L + i;
    double* sm_red = sm_pl;
    double* sm_axes = sm_pl + nw * 4;
    // ---- before the sweep is waited for: inputs that it does not write
    if (q.do_prep) {
        if (q.axes_smem) stage_axes(pa.axis_P, pa.axis_T, S, pa.N_P, pa.N_T, sm_axes);
        if (tid < S) sm_hasT[tid] = pa.has_T[tid];
    }
    __shared__ LevelPre sm_pre;                  // thread 0's pressure-only terms (kept out of its registers)
    double T1 = 0.0, T2 = 0.0, Pi = 0.0, mmr_s = 0.0;
    double trk_T = 0.0;
    int trk_sgn = 2, trk_flips = 0, trk_ncol = 0;
    if (q.do_update && tid == 0) {
        sm_pre = level_pre(u, b, i, u.P + (int64_t)b * L);
        T1 = u.T[li];
        T2 = (i == L - 1) ? T1 : u.T[li + 1];                            // :358-363
        if (u.trk_T) {
            trk_ncol = u.trk_ncol[b];
            trk_T = u.trk_T[li];
            trk_sgn = u.trk_state[li * 2]; trk_flips = u.trk_state[li * 2 + 1];
        }
    }
    if (q.do_prep && tid < S) { Pi = pa.P[li]; mmr_s = pa.mmr[li * S + tid]; }
    const int rows = __ldcg(q.plan_hdr);
    if (tid == 0) { STAMP_MIN(0); STAMP_MAX(1); }
    pdl_wait();                                  // partials come from the sweep before
    if (tid == 0) { STAMP_MIN(2); STAMP_MAX(3); }
    pdl_launch_dependents();                     // the next sweep's CTAs may line up behind this kernel
    if (q.active && !q.active[b]) return;        // converged atmosphere of a batch
    // ---- the four integrals of this level: rows x 4 contiguous doubles
    const double2* p = reinterpret_cast<const double2*>(q.partials + li * (int64_t)rows * 4);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    constexpr int kB = 4;                        // rows per thread in flight
    // the same number of load batches for every thread of the CTA (C2: 3125 rows = one batch of up to
    // four rows per thread); a thread adds its rows in ascending order
    for (int base = tid; base < rows + tid; base += kB * (int)blockDim.x) {
        double2 x[kB], y[kB];
#pragma unroll
        for (int k = 0; k < kB; ++k) {
            const int rr = base + k * (int)blockDim.x;
            const bool in = rr < rows;
            x[k] = in ? __ldcg(p + 2 * (int64_t)rr) : make_double2(0.0, 0.0);
            y[k] = in ? __ldcg(p + 2 * (int64_t)rr + 1) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int k = 0; k < kB; ++k)
            if (base + k * (int)blockDim.x < rows) { a0 += x[k].x; a1 += x[k].y; a2 += y[k].x; a3 += y[k].y; }
    }
    const double r4 = warp_reduce4(a0, a1, a2, a3, lane);
    if ((lane & 7) == 0) sm_red[warp * 4 + (lane >> 3)] = r4;
    __syncthreads();
    double s4 = 0.0;
    if (tid < 4) {
        for (int w = 0; w < nw; ++w) s4 += sm_red[w * 4 + tid];          // warps in order
        if (q.peer_bufs) {
            // fused one-shot all-reduce over the wavelength shards (flag-in-data words, see post_kernel):
            // this CTA exchanges the four integrals of its own level
            const int n = L * 4, e = i * 4 + tid;
            const int par = (int)(q.epoch & 1ull);
            const unsigned long long fl = (q.epoch & 0xffffffffull) << 32;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(s4);
            const unsigned long long w0 = (bits & 0xffffffffull) | fl, w1 = (bits >> 32) | fl;
            const int64_t slot = ((((int64_t)par * q.world + q.rank) * q.B + b) * n + e) * 2;
            for (int rk = 0; rk < q.world; ++rk) {
                const int dst = (q.rank + rk) % q.world;
                asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};"
                             ::"l"(q.peer_bufs[dst] + slot), "l"(w0), "l"(w1) : "memory");
            }
            double tot = 0.0;
            bool timed_out = false;
            for (int rk = 0; rk < q.world; ++rk) {
                const unsigned long long* src =
                    q.peer_bufs[q.rank] + ((((int64_t)par * q.world + rk) * q.B + b) * n + e) * 2;
                unsigned long long x0 = 0, x1 = 0;
                for (long long spins = 0;; ++spins) {
                    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];"
                                 : "=l"(x0), "=l"(x1) : "l"(src) : "memory");
                    if ((x0 & 0xffffffff00000000ull) == fl && (x1 & 0xffffffff00000000ull) == fl) break;
                    if (spins > (1ll << 27)) { timed_out = true; break; }
                }
                tot += __longlong_as_double((long long)((x0 & 0xffffffffull) | (x1 << 32)));
            }
            if (timed_out && q.p2p_error) *q.p2p_error = 1;
            s4 = tot;
        }
        q.sums[li * 4 + tid] = s4;
        sm_bc[tid] = s4;
    }
    if (tid == 0) { STAMP_MIN(4); STAMP_MAX(5); }
    if (!q.do_update) return;                    // reduction only (NCCL path)
    __syncthreads();
    // ---- temperature step of this level (thread 0)
    int conv = 1;
    if (tid == 0) {
        const double dT = delta_T_level(u, sm_pre, i, sm_bc, T1, T2);
        const double Tn = T1 - dT;                                       // :407, :536
        u.dT[li] = dT;
        q.T_next[li] = Tn;
        if (u.T_hist) u.T_hist[li] = Tn;
        sm_bc[4] = Tn;
        if (u.trk_T) {                                                    // per-level convergence, core.py:306-311
            int sgn_prev = trk_sgn, flips = trk_flips;
            if (trk_ncol > 0) {
                const double d = Tn - trk_T;
                const int sgn = (d != d) ? 3 : (d > 0.0) - (d < 0.0);
                if (sgn_prev != 2 && (sgn != sgn_prev || sgn == 3)) ++flips;
                sgn_prev = sgn;
            }
            u.trk_T[li] = Tn;
            u.trk_state[li * 2] = sgn_prev;
            u.trk_state[li * 2 + 1] = flips;
            conv = (flips > u.n_zero_crossings) || (fabs(dT) < u.convergence_dT);
        }
        STAMP_MAX(11);
    }
    __syncthreads();
    // ---- K0 of this level for the next sweep
    if (q.do_prep) {
        const double Tn = sm_bc[4];
        const double* axP = q.axes_smem ? sm_axes : pa.axis_P;
        const double* axT = q.axes_smem ? sm_axes + S * pa.N_P : pa.axis_T;
        if (tid < S) (void)prep_pair(pa, b, i, tid, axP, axT, sm_hasT, Pi, Tn, mmr_s);
        if (tid == 32) {                                                  // another warp: the two scalars of the record
            double* rec = pa.lp.rec + li * pa.lp.rec8;
            const double* Pb = pa.P + (int64_t)b * L;
            const double p1 = Pb[i] * FREI_BAR;
            const double p2 = (i == L - 1) ? p1 * (Pb[L - 2] * FREI_BAR) / (Pb[L - 3] * FREI_BAR) : Pb[i + 1] * FREI_BAR;
            rec[0] = (p1 - p2) / pa.g[b];                                 // twostream.py:231 (identical every sweep)
            rec[1] = (Tn > 1e-3 && Tn < 1e9) ? fast_rcp(Tn) : 1.0 / Tn;
            reinterpret_cast<int64_t*>(rec)[2 + 5 * S] = 0;               // same-rows flag: made by the sweep
        }
    }
    if (tid == 0) { STAMP_MAX(13); }
    // ---- last CTA of the atmosphere: publish the new temperatures, close the tracker
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        if (u.trk_T && conv) atomicAdd(q.conv_count + b, 1u);
        __threadfence();
        is_last = (atomicAdd(q.counters + b, 1u) == (unsigned)(L - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int e = tid; e < L; e += blockDim.x) u.T[(int64_t)b * L + e] = __ldcg(q.T_next + (int64_t)b * L + e);
    if (tid == 0) {
        if (u.trk_T) {
            const int ncol = u.trk_ncol[b] + 1;
            u.trk_ncol[b] = ncol;
            const unsigned int nconv = __ldcg(q.conv_count + b);
            if (u.direction == FREI_ABSORB && nconv == (unsigned)L && u.active) {   // core.py:317
                u.active[b] = 0;
                if (u.trk_iters) u.trk_iters[b] = ncol / 2;
            }
            q.conv_count[b] = 0u;
        }
        q.counters[b] = 0u;                      // self-cleaning for the next launch
    }
}

// ---------------------------------------------------------------------------
// host side of the ABI
// ---------------------------------------------------------------------------
// ---- launch plan -------------------------------------------------------------------------------
// The wavelength axis is cut into n2 warp-chunks of 64 wavelengths (two per thread) followed by n1
// chunks of 32 (one per thread); the partials have one [L][4] row per chunk, and the sweep kernel
// leaves the row count in the workspace header for the reduction that follows it.  For a single
// atmosphere the grid is capped at one resident wave of CTAs (occupancy x SMs) and the kernel
// strides over the chunks, which balances every round over the SMs: a column is a serial recurrence
// over the layers, so with the hardware's dynamic CTA placement the SMs that drained first took
// whole extra CTAs while the others idled (C2: 1.32 waves cost 1.7x one wave).  Complete rounds
// run 64-wide chunks; a last round that is at most half full is cut into 32-wide chunks (see
// sweep_kernel).  Batches keep one CTA per kWarps chunks: converged atmospheres exit at once and
// must not strand the work of the others.
constexpr int kMaxDevices = 64;
static int num_sms() {
    static int cache[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cache[dev];
}

// test hook (frei_b200_debug_plan): 0 = automatic, 1 = 32-wide chunks only, 2 = 64-wide chunks only,
// 3 = half of the 64-wide chunks replaced by 32-wide ones (exercises the mixed kernel at test sizes)
static int g_force_V = 0;

struct SweepPlan { int n2, n1, relay_quota, relay_warps; };
// slots = warps of one resident wave (0 = not capped: batches), NS = layer-steps of a chunk
static SweepPlan sweep_plan(int64_t n_lam, int B, int64_t slots, int NS, bool may_relay) {
    SweepPlan p;
    p.relay_quota = 0; p.relay_warps = 0;
    const int64_t c1 = (n_lam + 31) / 32, c2 = (n_lam + 63) / 64;
#if SWEEP_RELAY
    // Relay plan: more 64-wide chunks than resident warps -> equal runs of layer-steps per warp
    // (sweep_kernel).  quota >= NS keeps every chunk in at most two pieces and gives the receiver of
    // a piece quota - NS steps of slack.  Test hook 4 builds one at any size: 2 warps per 3 chunks.
    if (may_relay && n_lam % 2 == 0 && B == 1 && c2 * NS < (int64_t)1 << 30 && (g_force_V == 0 || g_force_V == 4)) {
        int64_t warps = (g_force_V == 4) ? (2 * c2 + 2) / 3 : slots;
        if (warps > 0 && c2 > warps && (g_force_V == 4 || c2 >= SWEEP_V1_CHUNKS_PER_SM * (int64_t)num_sms())) {
            const int64_t quota = (c2 * NS + warps - 1) / warps;
            p.n2 = (int)c2; p.n1 = 0;
            p.relay_quota = (int)quota;
            p.relay_warps = (int)((c2 * NS + quota - 1) / quota);
            return p;
        }
    }
#endif
    if (n_lam % 2 != 0 || g_force_V == 1) { p.n2 = 0; p.n1 = (int)c1; return p; }
    if (g_force_V == 3) { p.n2 = (int)(c2 / 2); p.n1 = (int)((n_lam - 64 * (int64_t)p.n2 + 31) / 32); return p; }
    p.n2 = (int)c2; p.n1 = 0;
    if (g_force_V == 2 || g_force_V == 4) return p;
    // so small that 64-wide chunks would leave SMs without a warp (C1's 5k bins = 79 chunks on 148
    // SMs): a chunk is a serial recurrence, so halving it is the only parallelism left
    if ((int64_t)B * c2 < SWEEP_V1_CHUNKS_PER_SM * (int64_t)num_sms()) { p.n2 = 0; p.n1 = (int)c1; return p; }
    if (B > 1 || slots <= 0 || !SWEEP_MIXED_TAIL) return p;
    const int64_t full = n_lam / (64 * slots);            // complete rounds of 64-wide chunks
    const int64_t rem = n_lam - full * 64 * slots;        // wavelengths left for the last round
    if (full >= 1 && rem > 0 && rem <= 32 * slots) {
        p.n2 = (int)(full * slots);
        p.n1 = (int)((rem + 31) / 32);
    }
    return p;
}

#ifndef SWEEP_PERSISTENT
#define SWEEP_PERSISTENT 1        // experiment knob: 0 = one CTA per kWarps chunks for every launch
#endif

// per device and kernel: CTAs per SM at the largest shared-memory size seen
struct SweepOcc {
    int resident[kMaxDevices];
    size_t smem_set[kMaxDevices];
    bool init = false;
};
template <typename K>
static int sweep_occupancy(K kern, SweepOcc& o, size_t smem, int* per_sm) {
    if (!o.init) { for (int d = 0; d < kMaxDevices; ++d) { o.resident[d] = -1; o.smem_set[d] = 0; } o.init = true; }
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return set_err(FREI_E_UNSUPPORTED, "device ordinal out of range%s%s");
    if (o.resident[dev] < 0 || smem > o.smem_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      (int)cudaSharedmemCarveoutMaxShared));
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem));
        o.resident[dev] = nb > 0 ? nb : 1;
        o.smem_set[dev] = smem;
    }
    *per_sm = o.resident[dev];
    return FREI_OK;
}

template <typename TabT, int S_T, int DIR, bool DTAUS>
static int launch_sweep_one(SweepArgs a, size_t smem, cudaStream_t st) {
    static SweepOcc occ, occ_relay;
    auto kern = sweep_kernel<TabT, S_T, DIR, DTAUS, false>;
    int per_sm = 1;
    int rc = sweep_occupancy(kern, occ, smem, &per_sm);
    if (rc) return rc;
    // One resident wave for a single atmosphere, with as many CTAs per SM as fit: launching fewer
    // to trade a nearly empty last round for fuller ones was measured (2, 3, 4 CTAs/SM at 100k ...
    // 600k wavelengths, scripts/ctas_scan.sh) and never won — 4 CTAs/SM are 0 ... 18 % faster than 3.
    const bool persistent = SWEEP_PERSISTENT && a.B == 1;
    int64_t cap = (int64_t)per_sm * num_sms();
    // the relay kernel exists for sweeps without the dtaus output (every sweep of a solve but the last)
    const bool may_relay = SWEEP_RELAY && !DTAUS && persistent && a.relay_flags;
    SweepPlan p = sweep_plan(a.n_lam, a.B, persistent ? cap * kWarps : 0, a.L - 1, may_relay);
    if (!DTAUS && p.relay_quota > 0) {
        auto kern_r = sweep_kernel<TabT, S_T, DIR, false, true>;
        int per_sm_r = 1;
        rc = sweep_occupancy(kern_r, occ_relay, smem, &per_sm_r);
        if (rc) return rc;
        if (per_sm_r == per_sm || g_force_V == 4) {     // same residency as the plan assumed: every warp of the grid is resident
            a.n2 = p.n2;
            a.rows = p.n2;
            a.relay_quota = p.relay_quota;
            const unsigned blocks = (unsigned)((p.relay_warps + kWarps - 1) / kWarps);   // <= cap
            CUDA_TRY(launch_pdl(kern_r, dim3(blocks, 1), dim3(kThreads), smem, st, a));
            return FREI_OK;
        }
        p = sweep_plan(a.n_lam, a.B, cap * kWarps, a.L - 1, false);
    }
    a.n2 = p.n2;
    a.rows = p.n2 + p.n1;
    a.relay_quota = 0;
    unsigned blocks = (unsigned)((a.rows + kWarps - 1) / kWarps);
    if (persistent && blocks > cap) blocks = (unsigned)cap;
    CUDA_TRY(launch_pdl(kern, dim3(blocks, a.B), dim3(kThreads), smem, st, a));
    return FREI_OK;
}

template <typename TabT, int S_T>
static int launch_sweep_dir(const SweepArgs& a, int direction, size_t smem, cudaStream_t st) {
    if (direction == FREI_EMIT)
        return a.dtaus ? launch_sweep_one<TabT, S_T, FREI_EMIT, true>(a, smem, st)
                       : launch_sweep_one<TabT, S_T, FREI_EMIT, false>(a, smem, st);
    return a.dtaus ? launch_sweep_one<TabT, S_T, FREI_ABSORB, true>(a, smem, st)
                   : launch_sweep_one<TabT, S_T, FREI_ABSORB, false>(a, smem, st);
}

template <typename TabT>
static int launch_sweep(const SweepArgs& a, int direction, size_t smem, cudaStream_t st) {
    switch (a.S) {
        case 1: return launch_sweep_dir<TabT, 1>(a, direction, smem, st);
        case 3: return launch_sweep_dir<TabT, 3>(a, direction, smem, st);
        case 8: return launch_sweep_dir<TabT, 8>(a, direction, smem, st);
        default: return launch_sweep_dir<TabT, 0>(a, direction, smem, st);
    }
}

extern "C" {

const char* frei_b200_last_error(void) { return g_err; }
int frei_b200_abi_version(void) { return FREI_B200_ABI_VERSION; }

int frei_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int frei_b200_workspace_bytes(int32_t B, int32_t L, int32_t S, int64_t n_lam,
                              int64_t* layer_params, int64_t* partials, int64_t* sums, int64_t* dT) {
    ARG_TRY(B > 0 && L >= 3 && S > 0 && S <= kMaxS && n_lam > 0);
    if (layer_params) *layer_params = layer_params_bytes(B, L, S);
    if (partials)       // per-warp rows + chunk sums + per-atmosphere tickets
        *partials = ((int64_t)B * sweep_rows_max(n_lam) + (int64_t)B * kPostChunks) * L * 4 * 8 + round16((int64_t)B * 4) + 16 +
                    round16(sweep_rows_max(n_lam) * 4) + (int64_t)B * L * 8 + round16((int64_t)B * 4);
    if (sums) *sums = (int64_t)B * L * 4 * 8;
    if (dT) *dT = (int64_t)B * L * 8;
    return FREI_OK;
}

static int check_common(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws) {
    ARG_TRY(tab && atm && ws);
    ARG_TRY(tab->values && tab->axis_P && tab->axis_T && tab->has_T);
    ARG_TRY(tab->S > 0 && tab->S <= kMaxS && tab->N_P >= 2 && tab->N_T >= 2 && tab->n_lam > 0);
    ARG_TRY(tab->dtype == FREI_F32 || tab->dtype == FREI_F64);
    ARG_TRY(atm->T && atm->P && atm->mmr && atm->g && atm->m_bar && atm->alpha);
    ARG_TRY(atm->B > 0 && atm->L >= 3);
    ARG_TRY(ws->layer_params);
    return FREI_OK;
}

int frei_b200_spectral_setup(const double* d_lam_um, int64_t n_global, int64_t offset, int64_t n_local,
                             double m_bar, double T_star, double a_rstar, double f,
                             double* d_c1, double* d_c2, double* d_sigma, double* d_w, double* d_f_toa,
                             void* stream) {
    ARG_TRY(d_lam_um && d_c1 && d_c2 && d_sigma && d_w && d_f_toa);
    ARG_TRY(n_global > 0 && n_local > 0 && offset >= 0 && offset + n_local <= n_global);
    ARG_TRY(m_bar > 0 && T_star > 0 && a_rstar > 0);
    spectral_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_lam_um, n_global, offset, n_local, m_bar, T_star, a_rstar, f,
        d_c1, d_c2, d_sigma, d_w, d_f_toa);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_debug_plan(int32_t force_V) {
    ARG_TRY(force_V >= 0 && force_V <= 4);
    g_force_V = force_V;
    return FREI_OK;
}

int frei_b200_debug_plan_query(int64_t n_lam, int32_t B, int32_t L, int64_t resident_warps, int32_t may_relay,
                               int32_t* out4) {
    ARG_TRY(n_lam > 0 && B > 0 && L >= 3 && resident_warps >= 0 && out4);
    const SweepPlan p = sweep_plan(n_lam, B, resident_warps, L - 1, may_relay != 0);
    out4[0] = p.n2; out4[1] = p.n1; out4[2] = p.relay_quota; out4[3] = p.relay_warps;
    return FREI_OK;
}

int frei_b200_debug_math(const double* d_x, double* d_out, int64_t n, void* stream) {
    ARG_TRY(d_x && d_out && n > 0);
    debug_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_layer_prep(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                         int32_t* d_iP, int32_t* d_iT, double* d_wP, double* d_wT, uint8_t* d_oob,
                         void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    PrepArgs a;
    a.axis_P = tab->axis_P; a.axis_T = tab->axis_T; a.has_T = tab->has_T;
    a.T = atm->T; a.P = atm->P; a.mmr = atm->mmr; a.g = atm->g;
    a.lp = layer_params_view(ws->layer_params, tab->S);
    a.iP = d_iP; a.iT = d_iT; a.wP = d_wP; a.wT = d_wT; a.oob = d_oob;
    a.B = atm->B; a.L = atm->L; a.S = tab->S; a.N_P = tab->N_P; a.N_T = tab->N_T;
    a.n_lam = tab->n_lam;
    const size_t axes = prep_axes_bytes(tab->S, tab->N_P, tab->N_T);
    prep_kernel<<<atm->B, 128, axes, (cudaStream_t)stream>>>(a, axes > 0);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_kappa(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                    const frei_workspace* ws, double* d_k, double* d_sigma, void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    ARG_TRY(spec && spec->sigma && d_k && d_sigma && spec->n_lam == tab->n_lam);
    ARG_TRY(atm->L <= 65535 && atm->B <= 65535);
    LayerParams lp = layer_params_view(ws->layer_params, tab->S);
    dim3 grid((unsigned)((tab->n_lam + 255) / 256), atm->L, atm->B);
    if (tab->dtype == FREI_F32)
        kappa_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float*)tab->values, spec->sigma, atm->sigma_scale, lp, d_k, d_sigma,
            atm->L, tab->S, tab->N_T, tab->n_lam);
    else
        kappa_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const double*)tab->values, spec->sigma, atm->sigma_scale, lp, d_k, d_sigma,
            atm->L, tab->S, tab->N_T, tab->n_lam);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_propagate(const double* d_lam_cm, const double* d_F1_up, const double* d_F2_down,
                        double T1, double T2, const double* d_delta_tau, const double* d_omega0,
                        double* d_F2_up, double* d_F1_down, int64_t n, void* stream) {
    ARG_TRY(d_lam_cm && d_F1_up && d_F2_down && d_delta_tau && d_omega0 && d_F2_up && d_F1_down);
    ARG_TRY(n > 0);
    propagate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_lam_cm, d_F1_up, d_F2_down, T1, T2, d_delta_tau, d_omega0, d_F2_up, d_F1_down, n);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_sweep(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                    const frei_flux* flux, int32_t direction, const frei_workspace* ws, void* stream) {
    int rc = check_common(tab, atm, ws);
    if (rc) return rc;
    ARG_TRY(spec && spec->c1 && spec->c2 && spec->sigma && spec->w && spec->f_toa);
    ARG_TRY(spec->n_lam == tab->n_lam);
    ARG_TRY(flux && flux->F_up && flux->F_down && ws->partials);
    ARG_TRY(direction == FREI_EMIT || direction == FREI_ABSORB);
    ARG_TRY(flux->dtype == FREI_F64 || flux->dtype == FREI_F32);
    ARG_TRY(atm->B <= 65535);
    SweepArgs a;
    a.tab = tab->values;
    a.c1 = spec->c1; a.c2 = spec->c2; a.sigma = spec->sigma; a.w = spec->w; a.f_toa = spec->f_toa;
    a.sigma_scale = atm->sigma_scale; a.ftoa_scale = atm->ftoa_scale; a.active = atm->active;
    a.lp = layer_params_view(ws->layer_params, tab->S);
    a.F_up = flux->F_up; a.F_down = flux->F_down; a.dtaus = flux->dtaus;
    a.partials = ws->partials;
    a.n_lam = tab->n_lam; a.B = atm->B; a.L = atm->L; a.S = tab->S; a.N_T = tab->N_T;
    a.rows = 0; a.n2 = 0;                          // set by the launcher from its plan
    a.plan_hdr = ws_plan_hdr(ws, atm->B, atm->L, tab->n_lam);
    a.relay_quota = 0;
    a.relay_flags = ws_relay_flags(ws, atm->B, atm->L, tab->n_lam);
    if (flux->dtype == FREI_F32) {               // fp32 arithmetic: sweep_f32.cu, same partials layout
        const SweepPlan p = sweep_plan(tab->n_lam, atm->B, 0, atm->L - 1, false);
        return frei_launch_sweep_f32(a, tab->dtype, direction, p.n2 > 0 ? 2 : 1, (cudaStream_t)stream);
    }
#ifndef SWEEP_SMEM_PAD
#define SWEEP_SMEM_PAD 0          // experiment knob: extra dynamic shared memory to cap CTAs/SM
#endif
    // level records + thread-private slots for 4 S table elements of two wavelengths
    const size_t smem = (size_t)atm->L * a.lp.rec8 * sizeof(double) + SWEEP_SMEM_PAD +
                        (size_t)4 * tab->S * kThreads * 2 * (tab->dtype == FREI_F32 ? 4 : 8);
    if (smem > 200 * 1024)
        return set_err(FREI_E_UNSUPPORTED, "L * (species + layers) state exceeds shared memory%s%s");
    return (tab->dtype == FREI_F32) ? launch_sweep<float>(a, direction, smem, (cudaStream_t)stream)
                                    : launch_sweep<double>(a, direction, smem, (cudaStream_t)stream);
}

static void fill_prep(PrepArgs& a, const frei_table* tab, const frei_atmosphere* atm,
                      const frei_workspace* ws) {
    a.axis_P = tab->axis_P; a.axis_T = tab->axis_T; a.has_T = tab->has_T;
    a.T = atm->T; a.P = atm->P; a.mmr = atm->mmr; a.g = atm->g;
    a.lp = layer_params_view(ws->layer_params, tab->S);
    a.iP = nullptr; a.iT = nullptr; a.wP = nullptr; a.wT = nullptr; a.oob = nullptr;
    a.B = atm->B; a.L = atm->L; a.S = tab->S; a.N_P = tab->N_P; a.N_T = tab->N_T;
    a.n_lam = tab->n_lam;
}

static void fill_update(UpdateArgs& u, const frei_atmosphere* atm, const frei_workspace* ws,
                        int direction, double alpha_override, double* d_T_hist) {
    u.T = atm->T; u.P = atm->P; u.g = atm->g; u.m_bar = atm->m_bar; u.alpha = atm->alpha;
    u.dT = ws->dT; u.T_hist = d_T_hist;
    u.L = atm->L; u.direction = direction; u.alpha_override = alpha_override;
    u.active = atm->active;
    const frei_tracker* t = atm->tracker;
    u.trk_T = t ? t->last_T : nullptr;
    u.trk_state = t ? t->state : nullptr;
    u.trk_ncol = t ? t->n_columns : nullptr;
    u.trk_iters = t ? t->iterations : nullptr;
    u.n_zero_crossings = t ? t->n_zero_crossings : 0;
    u.convergence_dT = t ? t->convergence_dT : 0.0;
}

// reduce (+ update T (+ rebuild records)) in one launch
static int launch_post(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                       int64_t n_lam, int do_update, int do_prep, int direction,
                       double alpha_override, double* d_T_hist, cudaStream_t st,
                       const frei_p2p* p2p = nullptr) {
    ARG_TRY(atm && ws && ws->partials && ws->sums && n_lam > 0);
    ARG_TRY(atm->L >= 3 && atm->L <= 256 && atm->B <= 65535);
    PostArgs q;
    // The row count of the sweep's plan is read on the device from the workspace header; the grid
    // only needs a bound.  Two reduction stages of about sqrt(rows) rows each keep the dependent
    // load batches of both short (C1: 9 chunks, C2: 56, C3 on one GPU: 125); at most kPostChunks,
    // the workspace layout's bound.
    const int64_t rows_est = (n_lam + 63) / 64;
    q.nchunks = (int)ceil(sqrt((double)rows_est));
    if (q.nchunks > kPostChunks) q.nchunks = kPostChunks;
    if (q.nchunks < 1) q.nchunks = 1;
    const int64_t n = (int64_t)atm->L * 4;
    q.partials = ws->partials;
    q.chunk_sums = ws_chunk_sums(ws, atm->B, atm->L, n_lam);
    q.counters = ws_counters(ws, atm->B, atm->L, n_lam);
    q.plan_hdr = ws_plan_hdr(ws, atm->B, atm->L, n_lam);
    q.sums = ws->sums;
    q.active = atm->active;
    q.peer_bufs = nullptr; q.p2p_error = nullptr;
    q.epoch = 0; q.rank = 0; q.world = 1; q.B = atm->B;
    if (p2p) {
        ARG_TRY(p2p->peer_bufs && p2p->world >= 1 && p2p->world <= 64);
        ARG_TRY(p2p->rank >= 0 && p2p->rank < p2p->world && (p2p->epoch & 0xffffffffull) != 0);
        q.peer_bufs = (unsigned long long* const*)p2p->peer_bufs;
        q.p2p_error = p2p->error; q.epoch = p2p->epoch; q.rank = p2p->rank; q.world = p2p->world;
    }
    q.do_update = do_update; q.do_prep = do_prep;
    UpdateArgs u{};
    PrepArgs pa{};
    if (do_update) {
        ARG_TRY(ws->dT && atm->T && atm->P && atm->g && atm->m_bar && atm->alpha);
        ARG_TRY(direction == FREI_EMIT || direction == FREI_ABSORB);
        fill_update(u, atm, ws, direction, alpha_override, d_T_hist);
    } else {
        u.L = atm->L;
    }
    if (do_prep) {
        int rc = check_common(tab, atm, ws);
        if (rc) return rc;
        fill_prep(pa, tab, atm, ws);
    }
#if POST_PER_LEVEL
    {   // one CTA per level: threads ~ rows (a thread sums its rows 4 at a time), shared memory = warp sums + axes
        q.T_next = ws_T_next(ws, atm->B, atm->L, n_lam);
        q.conv_count = ws_conv_count(ws, atm->B, atm->L, n_lam);
        int threads = (int)((rows_est + 31) / 32 * 32);
        if (threads < 128) threads = 128;
        if (threads > 1024) threads = 1024;
        const size_t axes = do_prep ? prep_axes_bytes(tab->S, tab->N_P, tab->N_T) : 0;
        q.axes_smem = axes > 0;
        q.stage_levels = 0;
        const size_t smem = (size_t)(threads / 32) * 4 * sizeof(double) + axes;
        (void)n;
        CUDA_TRY(launch_pdl(post_level_kernel, dim3(atm->L, atm->B), dim3(threads), smem, st, q, u, pa));
        return FREI_OK;
    }
#endif
    // threads = G row groups x n columns, at least L (update) and n = 4 L (columns)
    int G = 1024 / (int)n;
    if (G < 1) return set_err(FREI_E_UNSUPPORTED, "more than 256 levels%s%s");
    if (G > 8) G = 8;
    const int threads = (G * (int)n + 31) / 32 * 32;
    const size_t axes = do_prep ? prep_axes_bytes(tab->S, tab->N_P, tab->N_T) : 0;
    q.axes_smem = axes > 0;
    // T, P (update) and mmr, row offsets (K0) of the atmosphere, staged before the wait on the sweep
    const size_t levels = do_update ? (size_t)atm->L * (2 + (do_prep ? 2 * tab->S : 0)) * sizeof(double) : 0;
    size_t smem = (size_t)(G + 1) * n * sizeof(double) + axes;
    q.stage_levels = (levels > 0 && smem + levels <= 48 * 1024) ? 1 : 0;
    if (q.stage_levels) smem += levels;
    CUDA_TRY(launch_pdl(post_kernel, dim3(q.nchunks, atm->B), dim3(threads), smem, st, q, u, pa));
    return FREI_OK;
}

int frei_b200_reduce(const frei_atmosphere* atm, const frei_workspace* ws, int64_t n_lam, void* stream) {
    return launch_post(nullptr, atm, ws, n_lam, 0, 0, FREI_EMIT, -1.0, nullptr, (cudaStream_t)stream);
}

int frei_b200_update_T(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                       int32_t direction, double alpha_override, double* d_T_hist, void* stream) {
    ARG_TRY(atm && ws && ws->sums && ws->dT && atm->T && atm->P && atm->g && atm->m_bar && atm->alpha);
    ARG_TRY(atm->L >= 3 && atm->L <= 1024);
    ARG_TRY(direction == FREI_EMIT || direction == FREI_ABSORB);
    UpdateArgs u{};
    PrepArgs pa{};
    fill_update(u, atm, ws, direction, alpha_override, d_T_hist);
    if (tab) {
        int rc = check_common(tab, atm, ws);
        if (rc) return rc;
        fill_prep(pa, tab, atm, ws);
    }
    int threads = ((atm->L + 31) / 32) * 32;
    if (tab) {                                   // K0 runs one thread per (level, species)
        const int want = ((atm->L * tab->S + 31) / 32) * 32;
        threads = want > 256 ? (threads > 256 ? threads : 256) : (want > threads ? want : threads);
    }
    const size_t axes = tab ? prep_axes_bytes(tab->S, tab->N_P, tab->N_T) : 0;
    update_prep_kernel<<<atm->B, threads, axes, (cudaStream_t)stream>>>(u, pa, tab ? 1 : 0, ws->sums, axes > 0);
    CUDA_TRY(cudaGetLastError());
    return FREI_OK;
}

int frei_b200_post(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                   int64_t n_lam, int32_t direction, double alpha_override, double* d_T_hist,
                   int32_t prep_next, void* stream) {
    return launch_post(tab, atm, ws, n_lam, 1, prep_next ? 1 : 0, direction, alpha_override, d_T_hist,
                       (cudaStream_t)stream);
}

int frei_b200_post_p2p(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                       int64_t n_lam, int32_t direction, double alpha_override, double* d_T_hist,
                       int32_t prep_next, const frei_p2p* p2p, void* stream) {
    ARG_TRY(p2p);
    return launch_post(tab, atm, ws, n_lam, 1, prep_next ? 1 : 0, direction, alpha_override, d_T_hist,
                       (cudaStream_t)stream, p2p);
}

int frei_b200_sweep_step(const frei_table* tab, const frei_spectral* spec, const frei_atmosphere* atm,
                         const frei_flux* flux, int32_t direction, double alpha_override,
                         const frei_workspace* ws, double* d_T_hist, int32_t prep_first,
                         int32_t prep_next, void* stream) {
    int rc;
    if (prep_first) {
        rc = frei_b200_layer_prep(tab, atm, ws, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    rc = frei_b200_sweep(tab, spec, atm, flux, direction, ws, stream);
    if (rc) return rc;
    return frei_b200_post(tab, atm, ws, tab->n_lam, direction, alpha_override, d_T_hist, prep_next, stream);
}

}  // extern "C"
