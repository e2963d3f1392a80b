// common.cuh — declarations shared by the translation units of libfrei_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/frei_b200.h"

// constants (CGS, CODATA 2018 as shipped by astropy >= 4.0)
#define FREI_KB      1.380649e-16
#define FREI_MP      1.67262192369e-24
#define FREI_H       6.62607015e-27
#define FREI_C       2.99792458e10
#define FREI_SIGSB   5.6703744191844314e-5
#define FREI_BAR     1e6
#define FREI_PI      3.141592653589793

#ifndef SWEEP_THREADS
#define SWEEP_THREADS 128
#endif
#ifndef SWEEP_MINB
#define SWEEP_MINB 4              // resident CTAs per SM the sweep kernel is compiled for (<= 128 registers,
                                  // no spills; the host may launch fewer per SM, see choose_ctas_per_sm)
#endif
#ifndef SWEEP_MINB_EMIT
#define SWEEP_MINB_EMIT SWEEP_MINB  // experiment knob: a separate register budget for the emit instantiations (under ncu
                                  // the emit kernel is 6 % slower at 128 registers than at 158, absorb 2 % faster)
#endif
#ifndef SWEEP_WIDE_S
#define SWEEP_WIDE_S 8            // species counts from here on: the staged table rows (4 S x 16 B per thread) cap the
#endif                            // resident CTAs below SWEEP_MINB anyway, so the kernel may use the registers
#ifndef SWEEP_MINB_WIDE
#define SWEEP_MINB_WIDE 2         // resident CTAs per SM those instantiations are compiled for (S = 8, 100 levels:
                                  // 99 KB of shared memory per CTA -> 2 CTAs; 206 registers; +9 % over the
                                  // 128-register build, profiles/r02_wide_species_ab.log)
#endif
#ifndef SWEEP_GATHER4
#define SWEEP_GATHER4 1           // 1: four accumulation chains in the opacity gather of those instantiations
#endif
#ifndef SWEEP_V1_CHUNKS_PER_SM
#define SWEEP_V1_CHUNKS_PER_SM 2  // below this many 64-wavelength chunks per SM the plan uses one wavelength per thread
#endif
#ifndef SWEEP_MIXED_TAIL
#define SWEEP_MIXED_TAIL 0        // 1: a last round of chunks that is at most half full runs 32-wide chunks (round-2
                                  // experiment: no gain at any size; the relay plan replaces it)
#endif
#ifndef SWEEP_RELAY
#define SWEEP_RELAY 1             // 1: relay plan for single atmospheres with more 64-wide chunks than resident warps
#endif
#ifndef SWEEP_E_VOTE
#define SWEEP_E_VOTE 1            // 1: warp vote selects the E = 1 specialisation of the layer step
#endif
constexpr int kThreads = SWEEP_THREADS;  // threads per sweep CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMaxS = 32;
constexpr int kPostChunks = 148;        // stage-1 CTAs of the partials reduction (one per SM)

// ---------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------
// ws->layer_params holds one record of `rec8` 8-byte words per (atmosphere, level):
//   [0] dpg = (p1 - p2) / g      [1] invT = 1 / T_i
//   [2 + 4 s + c]  W[s][c]       mmr-premultiplied weight of corner c of species s
//   [2 + 4 S + s]  off[s]        int64 element offset of table row (iP, iT) of species s
//   [2 + 5 S]      flags         int64, bit 0: every off[s] equals that of level i - 1 (the sweep then
//                                keeps the table rows it already staged instead of copying them again)
// rec8 is even, so records and the W quadruples are 16-byte aligned: the sweep stages the
// L records of its atmosphere into shared memory with one TMA bulk copy.
struct LayerParams {
    double* rec;                // [B][L][rec8]
    int rec8;
    int S;
};
__host__ __device__ static inline int rec_words(int S) { return (3 + 5 * S + 1) & ~1; }

// ---------------------------------------------------------------------------
// K2+K3: arguments of the layer sweep (fp64 and fp32 arithmetic)
// ---------------------------------------------------------------------------
struct SweepArgs {
    const void* tab;
    const double* c1; const double* c2; const double* sigma; const double* w; const double* f_toa;
    const double* sigma_scale; const double* ftoa_scale;
    const uint8_t* active;      // [B] or null: atmospheres with 0 are skipped (batch convergence)
    LayerParams lp;
    void* F_up; void* F_down; void* dtaus;
    double* partials;           // [B][rows][L][4], one row per sweep warp
    int64_t n_lam;
    int rows;                   // partial rows of the sweep = warp-chunks (set by the launcher from its plan)
    int n2;                     // fp64 kernel: chunks [0, n2) are 64 wavelengths wide, [n2, rows) 32 wide
    int32_t* plan_hdr;          // workspace header: [0] = rows, written by the sweep, read by the post kernel
    int B, L, S, N_T;
    // relay plan (fp64 kernel, single atmosphere, more chunks than resident warps): the (chunk, layer-step)
    // pairs are dealt out in equal runs of relay_quota steps per resident warp, 0 = every warp keeps
    // whole chunks; relay_flags[chunk] hands a chunk cut by a run boundary from one warp to the next
    int relay_quota;
    unsigned int* relay_flags;  // [rows], zero between launches
};

// Partial sums of the wavelength integrals: element (atmosphere b, chunk row q, level i, integral k).
// POST_PER_LEVEL = 1: [B][L][rows][4] — the rows of a level are contiguous, because the kernel that
// follows a sweep gives every level its own CTA; 0: [B][rows][L][4] (round 1 layout, one CTA per
// block of rows).  `rows` = the row count of the launch (SweepArgs::rows).
#ifndef POST_PER_LEVEL
#define POST_PER_LEVEL 1
#endif
__device__ __forceinline__ double* part_base(double* partials, int b, int rows, int L, int q) {
    return POST_PER_LEVEL ? partials + ((int64_t)b * L * rows + q) * 4 : partials + ((int64_t)b * rows + q) * L * 4;
}
__device__ __forceinline__ int64_t part_level_stride(int rows) { return POST_PER_LEVEL ? (int64_t)rows * 4 : 4; }

// Sum four per-lane values across the warp; on return lanes 0, 8, 16, 24 hold the
// totals of v0, v1, v2, v3 respectively.  Fixed butterfly -> deterministic.
__device__ __forceinline__ double warp_reduce4(double v0, double v1, double v2, double v3, int lane) {
    const unsigned full = 0xffffffffu;
    const bool up16 = lane & 16;
    double s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3;     // what I send
    double k0 = up16 ? v2 : v0, k1 = up16 ? v3 : v1;     // what I keep
    k0 += __shfl_xor_sync(full, s0, 16);
    k1 += __shfl_xor_sync(full, s1, 16);
    const bool up8 = lane & 8;
    double s = up8 ? k0 : k1, k = up8 ? k1 : k0;
    k += __shfl_xor_sync(full, s, 8);
    k += __shfl_xor_sync(full, k, 4);
    k += __shfl_xor_sync(full, k, 2);
    k += __shfl_xor_sync(full, k, 1);
    return k;     // lane 0: v0, lane 8: v1, lane 16: v2, lane 24: v3
}

// ---- opacity rows: asynchronous staging into shared memory (cp.async) ------------------------
// Every thread copies the 4 S table elements (x V wavelengths) of the NEXT level into its own
// slots while it computes the current level, then folds them into k.  The slots of a thread are
// private to it, so no CTA barrier is involved — only cp.async.wait_group.
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
    if (BYTES == 32) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                     ::"r"(dst + 16), "l"((const char*)src + 16) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;"
                     ::"r"(dst), "l"(src), "n"(BYTES == 32 ? 16 : BYTES) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}


int frei_set_err(int code, const char* msg);
// fp32-arithmetic sweep (sweep_f32.cu): flux state and table in fp32, wavelength integrals in fp64
int frei_launch_sweep_f32(SweepArgs a, int table_dtype, int direction, int plan_V, cudaStream_t st);
