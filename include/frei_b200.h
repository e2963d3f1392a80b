/*
 * frei_b200.h — C ABI of the B200-native radiative-equilibrium hot path.
 *
 * The reference (bmorris3/frei) is pure Python and defines no FFI; the
 * functions below are what a ctypes binding in the reference would call in
 * place of the Python bodies cited next to each entry point (paths are into
 * the reference checkout, `frei/...`).  INTEGRATION.md shows that binding.
 *
 * Rules of the boundary
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every pointer named d_* / documented "device" is device memory owned by
 *     the caller (e.g. a torch allocation); the library only borrows it for
 *     the duration of the (asynchronous) call.  `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).
 *   - All functions return 0 on success and a negative FREI_E_* code on
 *     failure; frei_b200_last_error() returns the message of the last failure
 *     on the calling thread.  Nothing is thrown across the boundary.
 *   - There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with FREI_E_CUDA.
 *
 * Units: T [K]; P [bar] (frei/tp.py:32, table axis frei/opacity.py:253);
 * g [cm s^-2]; m_bar [g]; wavelength-derived constants in CGS; fluxes
 * [erg s^-1 cm^-3] (frei/twostream.py:13); opacities [cm^2 g^-1].
 * Layer index 0 = bottom of the atmosphere (highest pressure).
 */
#ifndef FREI_B200_H
#define FREI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FREI_B200_ABI_VERSION 2

enum {
    FREI_OK = 0,
    FREI_E_ARG = -1,      /* bad argument (null pointer, size, dtype, L < 3, ...) */
    FREI_E_CUDA = -2,     /* CUDA runtime error; see frei_b200_last_error()      */
    FREI_E_UNSUPPORTED = -3
};

enum { FREI_EMIT = 0, FREI_ABSORB = 1 };     /* sweep direction */
enum { FREI_F32 = 32, FREI_F64 = 64 };       /* storage dtypes  */

/* Opacity tables of all S species, one dense block, wavelength contiguous:
 * values[s][iP][iT][j].  Replaces the dict of xarray.DataArray handed to
 * kappa() (frei/opacity.py:203-209, 250-263).  Axes are ascending (xarray
 * sorts before interpolating).  has_T[s] == 0 reproduces the
 * "single unique temperature -> interpolate in pressure only" branch
 * (frei/opacity.py:256); such a species still occupies N_T (>= 2) rows. */
typedef struct {
    const void*    values;     /* device, S*N_P*N_T*n_lam elements of `dtype`      */
    const double*  axis_P;     /* device [S][N_P], bar                             */
    const double*  axis_T;     /* device [S][N_T], K                               */
    const int32_t* has_T;      /* device [S]                                       */
    int32_t        S, N_P, N_T;
    int32_t        dtype;      /* FREI_F32 | FREI_F64                              */
    int64_t        n_lam;      /* wavelength bins held by this device (row length) */
} frei_table;

/* Per-wavelength constants of this device's wavelength range.
 * c1 = 2 h c^2 / lam^5 and c2 = h c / (lam k_B) give Planck
 * B = c1 / expm1(c2 / T) (frei/twostream.py:64-67); sigma = Rayleigh H2+He
 * (frei/opacity.py:187-200, 233); w = trapezoid weights of the GLOBAL grid so
 * sum_j w_j F_j == np.trapz(F, lam) (frei/twostream.py:16-20) even when the
 * grid is sharded; f_toa = stellar flux (frei/core.py:48-55). */
typedef struct {
    const double* c1;
    const double* c2;
    const double* sigma;
    const double* w;
    const double* f_toa;
    int64_t       n_lam;
} frei_spectral;

/* Device-side convergence bookkeeping of a batch (optional): the rule of
 * Grid.emission_spectrum (frei/core.py:301-318) evaluated incrementally after every sweep,
 * so that a batch of atmospheres never synchronises with the host per atmosphere.
 * All arrays are device memory, zero-initialised by the caller except state[..][0] = 2. */
typedef struct {
    double*  last_T;            /* [B][L] temperature in the previous history column        */
    int32_t* state;             /* [B][L][2]: sign of the previous column difference (2 =
                                   none yet), number of sign changes so far                  */
    int32_t* n_columns;         /* [B] history columns so far (2 per iteration)              */
    int32_t* iterations;        /* [B] or NULL: iterations done when the atmosphere converged */
    int32_t  n_zero_crossings;  /* default 2, frei/core.py:233                               */
    double   convergence_dT;    /* default 3 K                                               */
} frei_tracker;

/* B independent atmospheres of L levels each. */
typedef struct {
    double*       T;            /* device [B][L], in/out                               */
    const double* P;            /* device [B][L], bar, bottom -> top                   */
    const double* mmr;          /* device [B][L][S] mass-mixing ratios = the output of
                                   chemistry() (frei/opacity.py:246-248)               */
    const double* g;            /* device [B]                                          */
    const double* m_bar;        /* device [B]                                          */
    const double* alpha;        /* device [B], mixing-length parameter                 */
    const double* sigma_scale;  /* device [B] or NULL: sigma_b = sigma * scale_b       */
    const double* ftoa_scale;   /* device [B] or NULL: F_TOA_b = f_toa * scale_b       */
    int32_t       B, L;
    uint8_t*      active;       /* device [B] or NULL: atmospheres whose flag is 0 are
                                   skipped by every kernel; cleared by the tracker when
                                   an atmosphere converges                              */
    const frei_tracker* tracker;/* host pointer or NULL                                */
} frei_atmosphere;

/* Flux state, mutated in place like the reference's fluxes_up / fluxes_down
 * (frei/twostream.py:392-394, 521-522). */
typedef struct {
    void*   F_up;      /* device [B][L][n_lam]                                         */
    void*   F_down;    /* device [B][L][n_lam]                                         */
    void*   dtaus;     /* device [B][L][n_lam] or NULL; row order as the reference
                          returns it (frei/twostream.py:352, 374, 487, 505)            */
    int32_t dtype;     /* FREI_F32 | FREI_F64 (also selects the arithmetic type)       */
} frei_flux;

/* Scratch owned by the caller; sizes from frei_b200_workspace_bytes(). */
typedef struct {
    void*   layer_params;  /* written by layer_prep, read by kappa / sweep */
    double* partials;      /* written by sweep, read by reduce; must be zero-filled once
                              before first use (the library leaves its ticket counters
                              and the sweep's relay flags at zero after every call) */
    double* sums;          /* [B][L][4]: wavelength integrals of F2_up, F2_down, F1_up,
                              F1_down per layer-step (the four bolometric_flux calls,
                              frei/twostream.py:396-398, 524-527)          */
    double* dT;            /* [B][L] temperature change of the last sweep  */
} frei_workspace;

const char* frei_b200_last_error(void);
int  frei_b200_abi_version(void);
int  frei_b200_device_count(void);

/* Bytes needed for each workspace member (any out pointer may be NULL). */
int frei_b200_workspace_bytes(int32_t B, int32_t L, int32_t S, int64_t n_lam,
                              int64_t* layer_params, int64_t* partials,
                              int64_t* sums, int64_t* dT);

/* Per-wavelength constants of the slice [offset, offset + n_local) of the global
 * grid d_lam_um[n_global] (micron, ascending): fills a frei_spectral's arrays.
 * Covers BB prefactors (frei/twostream.py:64-67), rayleigh_H2 + rayleigh_He
 * (frei/opacity.py:173-200, 233, pure functions of wavelength the reference
 * recomputes every layer-step), np.trapz weights (frei/twostream.py:16-20) and
 * F_TOA (frei/core.py:48-62) with f = 2/3. */
int frei_b200_spectral_setup(const double* d_lam_um, int64_t n_global, int64_t offset,
                             int64_t n_local, double m_bar, double T_star, double a_rstar,
                             double f, double* d_c1, double* d_c2, double* d_sigma,
                             double* d_w, double* d_f_toa, void* stream);

/* Test hook: for n positive inputs x writes out[0..5n) = 1/x, sqrt(x), 1/sqrt(x),
 * exp(-x), 1 - exp(-x) as computed by the kernels' branch-free fp64 routines. */
int frei_b200_debug_math(const double* d_x, double* d_out, int64_t n, void* stream);

/* Test hook: shape of the sweep plan.  The sweep cuts the wavelength axis into warp-chunks of 64
 * wavelengths (two per thread) followed by chunks of 32 (one per thread).  0 (default) = automatic:
 * complete rounds of resident warps take 64-wide chunks and a last round that is at most half full
 * is cut into 32-wide ones; 32-wide only for odd wavelength counts and for problems too small to
 * give every SM two warps.  1 / 2 force 32-wide / 64-wide chunks only, 3 forces a mixed plan (half
 * of the 64-wide chunks, the rest 32-wide), 4 forces the relay plan (what the automatic plan runs
 * for a single atmosphere with more 64-wide chunks than resident warps: the layer-steps of all chunks
 * are dealt out in equal runs per warp and a chunk cut by a run boundary is handed from one warp to
 * the next; forced: two warps per three chunks); 2, 3 and 4 only take effect for even counts, 4 only
 * for a single atmosphere.
 * Process-wide; not meant for production use: it exists so that small parity cases can exercise
 * every chunk shape of the production-size kernels. */
int frei_b200_debug_plan(int32_t force_V);

/* Test hook (host only, no device needed): the plan the sweep launcher would choose for n_lam
 * wavelengths, B atmospheres of L levels and `resident_warps` warps in one resident wave (0 = not
 * capped: batches).  out4 = {64-wide chunks, 32-wide chunks, relay quota (layer-steps per warp, 0 = whole
 * chunks), warps of the relay grid}.  With a relay quota q the layer-steps of the 64-wide chunks,
 * chunk after chunk, are cut into runs of q: warp m owns steps [m q, (m + 1) q). */
int frei_b200_debug_plan_query(int64_t n_lam, int32_t B, int32_t L, int64_t resident_warps, int32_t may_relay,
                               int32_t* out4);

/* K0.  Bracket (P_i, T_i) of every level in every species' axes with the rule
 * of scipy.interpolate's find_indices (reached from frei/opacity.py:261-263),
 * build mmr-premultiplied corner weights (zero when out of bounds:
 * fill_value=0, frei/opacity.py:241-244) and the per-layer scalars
 * (p1-p2)/g (frei/twostream.py:227-231) and 1/T.
 * Optional device outputs for bit-exact index tests, each [B][L][S] or NULL:
 * iP, iT (int32), wP, wT (double), oob (uint8). */
int frei_b200_layer_prep(const frei_table* tab, const frei_atmosphere* atm,
                         const frei_workspace* ws,
                         int32_t* d_iP, int32_t* d_iT, double* d_wP, double* d_wT,
                         uint8_t* d_oob, void* stream);

/* K1 (standalone).  k[b][i][j] = sum_s mmr * bilerp(table_s) + sigma, and
 * sigma_out[b][j]: kappa() for every level at once (frei/opacity.py:203-269).
 * Requires layer_prep.  d_k: [B][L][n_lam] doubles, d_sigma: [B][n_lam] doubles. */
int frei_b200_kappa(const frei_table* tab, const frei_spectral* spec,
                    const frei_atmosphere* atm, const frei_workspace* ws,
                    double* d_k, double* d_sigma, void* stream);

/* K2 (standalone).  propagate_fluxes() for one layer, elementwise over n
 * wavelengths with g_0 = 0 (frei/twostream.py:97-177).  All arrays device
 * doubles of length n; lam in cm. */
int frei_b200_propagate(const double* d_lam_cm, const double* d_F1_up,
                        const double* d_F2_down, double T1, double T2,
                        const double* d_delta_tau, const double* d_omega0,
                        double* d_F2_up, double* d_F1_down, int64_t n, void* stream);

/* K2+K3.  One sweep over all layers and this device's wavelengths: the loop
 * bodies of emit() (frei/twostream.py:356-405) or absorb() (:491-533) up to
 * and including the four wavelength integrals, left as per-block partials.
 * Requires layer_prep on the current T. */
int frei_b200_sweep(const frei_table* tab, const frei_spectral* spec,
                    const frei_atmosphere* atm, const frei_flux* flux,
                    int32_t direction, const frei_workspace* ws, void* stream);

/* Deterministic fixed-order reduction of the per-warp partials into ws->sums. */
int frei_b200_reduce(const frei_atmosphere* atm, const frei_workspace* ws,
                     int64_t n_lam, void* stream);

/* K4.  From ws->sums (after the cross-device sum when the wavelength axis is
 * sharded): div_bol_net_flux, convective_flux, delta_t_i, delta_temperature
 * (frei/twostream.py:23-43, 190-287) and T <- T - dT (:407, :536).
 * alpha_override >= 0 replaces atm->alpha (the final emit of
 * Grid.emission_spectrum does not forward alpha, frei/core.py:323-333).
 * d_T_hist (nullable) receives the new T as one extra [B][L] record.
 * tab (nullable): when given, the level records are rebuilt for the new T in
 * the same launch (= layer_prep for the next sweep). */
int frei_b200_update_T(const frei_table* tab, const frei_atmosphere* atm,
                       const frei_workspace* ws, int32_t direction, double alpha_override,
                       double* d_T_hist, void* stream);

/* Single-device fusion of reduce + update_T (+ layer_prep for the next sweep when
 * prep_next != 0) in one launch. */
int frei_b200_post(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                   int64_t n_lam, int32_t direction, double alpha_override, double* d_T_hist,
                   int32_t prep_next, void* stream);

/* Wavelength-sharded mode without a separate collective: reduce + all-reduce + update_T
 * (+ layer_prep) in ONE launch.  Every rank stores its [B][L][4] integrals into all ranks'
 * exchange buffers over NVLink peer memory as self-validating 8-byte words (32 payload bits +
 * the 32-bit epoch: no fence and no separate flag), polls its own buffer until the words of all
 * ranks carry the current epoch and adds the contributions in rank order (bit-identical sums on
 * every rank).
 * peer_bufs: device array of `world` device pointers (peer-mapped, e.g. from
 * torch.distributed._symmetric_memory) to each rank's buffer of 2 * world * B * L*4 * 2 8-byte
 * words, zero-initialised before the first sweep.  epoch = 1, 2, 3, ... must advance identically on
 * all ranks (one per sweep; its low 32 bits must not be 0).  *error (device int, nullable) is set
 * to 1 if a peer does not arrive within a few seconds; the sums are then invalid and the caller
 * must check it before using T (frei_b200/engine.py: Engine.check_errors). */
typedef struct {
    void* const* peer_bufs;
    int32_t* error;
    uint64_t epoch;
    int32_t rank, world;
} frei_p2p;

int frei_b200_post_p2p(const frei_table* tab, const frei_atmosphere* atm, const frei_workspace* ws,
                       int64_t n_lam, int32_t direction, double alpha_override, double* d_T_hist,
                       int32_t prep_next, const frei_p2p* p2p, void* stream);

/* Convenience for one device: [layer_prep if prep_first] + sweep + post, i.e. one
 * emit()/absorb() call of the reference with n_timesteps=1 (frei/core.py:275-299). */
int frei_b200_sweep_step(const frei_table* tab, const frei_spectral* spec,
                         const frei_atmosphere* atm, const frei_flux* flux,
                         int32_t direction, double alpha_override,
                         const frei_workspace* ws, double* d_T_hist,
                         int32_t prep_first, int32_t prep_next, void* stream);

/* Load-time wavelength binning (frei/interp.py:156-202, 270-307 as called from
 * frei/opacity.py:137-139): trapezoid sum of consecutive samples that fall in the same bin, for
 * every leading (temperature, pressure) row.
 * d_a: [n_rows][row_stride] samples (FREI_F32 | FREI_F64), n_samples <= row_stride;
 * d_x: NULL = unit sample spacing (the reference's Trapz aggregation, x = None), else the sample
 *   positions [n_samples]: trapezoids of width x[i+1] - x[i] (xarray's integrate('wavelength') of
 *   the groupies=False branch, frei/opacity.py:29-40);
 * runs of equal consecutive bin codes [run_start[r], run_end[r]) grouped by bin:
 * bin b owns runs bin_first_run[b] .. bin_first_run[b+1]-1 (CSR, n_bins + 1 entries);
 * d_out: [n_rows][n_bins] doubles. */
int frei_b200_bin_trapz(const void* d_a, int32_t dtype, const double* d_x, int64_t n_rows,
                        int64_t n_samples, int64_t row_stride, const int64_t* d_run_start,
                        const int64_t* d_run_end, const int32_t* d_bin_first_run,
                        int32_t n_bins, double* d_out, void* stream);

/* Load-time regridding of a binned table onto the Grid (frei/opacity.py:141-146, 31-33, 163-166):
 * d_out [mT][mP][m] <- d_binned [nT][nP][nb] at the source nodes d_src_T [mT], d_src_P [mP] (the
 * nearest-neighbour indices, host-computed with scipy's interp1d(kind='nearest') rule).
 * d_j0 == NULL: the wavelength axis is copied (m == nb).  Otherwise linear interpolation with
 * extrapolation along wavelength in scipy's interp1d form: between source columns d_j0[j] and
 * d_j0[j] + 1, y = (y_hi - y_lo) / d_dx[j] * d_t[j] + y_lo with d_dx = x_hi - x_lo and
 * d_t = x_new - x_lo (NaN nodes propagate as in scipy). */
int frei_b200_regrid(const double* d_binned, int32_t nT, int32_t nP, int64_t nb,
                     const int32_t* d_src_T, int32_t mT, const int32_t* d_src_P, int32_t mP,
                     const int32_t* d_j0, const double* d_dx, const double* d_t, int64_t m,
                     double* d_out, void* stream);

/* Device-side post-processing of a finished solve, one thread per wavelength of
 * this device's slice; nothing but three sums has to leave the GPU for T_eff.
 *   d_dtaus [L][n_lam], d_spec [n_lam] (fluxes_up of the top level) as returned by the final
 *   emit; d_lam_um / d_w: this slice of the wavelength grid [micron] and of the global
 *   trapezoid weights [cm] (frei_b200_spectral_setup); d_P_bar, d_T [L].
 * Outputs:
 *   d_pressure_milne [n_lam] or NULL: np.interp(2/3, np.exp(-dtaus[:, j]), pressures) with
 *     numpy's own search path (the sequence is not sorted in general), frei/core.py:392-395;
 *   d_cf [L][n_lam] or NULL: contribution function normalised per wavelength, level order
 *     (the reference's cf[::-1]), frei/plot.py:63-79, 83;
 *   d_sums[3] = { sum_j p_milne_j F_j lam_j,  sum_j F_j lam_j,  sum_j w_j F_j }: numerator and
 *     denominator of the flux-weighted mean pressure (frei/core.py:397-401) and
 *     np.trapz(spec.flux, lam) (frei/core.py:413); with a sharded wavelength axis the caller adds
 *     the three numbers over the ranks.
 * d_scratch: frei_b200_diagnostics_scratch_bytes(n_lam) bytes.  Fixed summation order. */
int64_t frei_b200_diagnostics_scratch_bytes(int64_t n_lam);
int frei_b200_diagnostics(const double* d_dtaus, const double* d_spec, const double* d_lam_um,
                          const double* d_w, const double* d_P_bar, const double* d_T, int32_t L,
                          int64_t n_lam, double* d_pressure_milne, double* d_cf, double* d_scratch,
                          double* d_sums, void* stream);

/* Measurement aid (bench.py): thread-level fp64 fused multiply-adds per second this device
 * sustains with register operands and 16 resident warps per scheduler (best of three timed
 * launches, CUDA events on `stream`, synchronises).  It is the ceiling the fp64 sweep is compared
 * with; MEASURED_PEAKS.json has no fp64 entry.  d_scratch: at least 2048 doubles per SM. */
int frei_b200_fp64_peak(double* d_scratch, int64_t scratch_doubles, double* h_dfma_per_s,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FREI_B200_H */
