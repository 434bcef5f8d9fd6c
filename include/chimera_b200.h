/* chimera_b200.h -- C ABI of libchimera_b200.so (sm_100a CUDA implementation of CHIMERA's
 * hierarchical-likelihood hot path).
 *
 * The reference (CosmoStatGW/CHIMERA v2.0.0) is pure Python/JAX and has NO FFI of its own;
 * each entry point below names the reference interface it stands in for (paths relative to
 * the reference checkout).  A maintainer binds these with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All floating data is IEEE double,
 *     pixel ids are int64 (the reference's dtypes: utils/config.py:5, data.py:317).
 *   - every function returns CHB_OK (0) or a negative chb_status; nothing throws across the
 *     ABI.  A human-readable message for the last failure on a handle: chb_last_error().
 *   - numerical failures are VALUES (-inf / -DBL_MAX / NaN / +inf), exactly as in the
 *     reference (likelihood.py:296-297, selection_function.py:47); only invalid arguments
 *     and CUDA failures are errors.
 *   - the caller owns all host buffers; chb_set_* copy them to the device once; the handle
 *     owns the device memory.  A handle is bound to one CUDA device and is not thread-safe
 *     (one handle per rank / GPU).
 *   - there is no CPU fallback: without a CUDA device chb_create fails with CHB_ERR_CUDA.
 */
#ifndef CHIMERA_B200_H
#define CHIMERA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_ABI_VERSION 2     /* 2: chb_set_option (per-handle tuning switches; no environment variables are read) */

typedef enum {
  CHB_OK = 0,
  CHB_ERR_INVALID = -1,    /* bad argument / inconsistent shapes (ValueError in the shim)   */
  CHB_ERR_CUDA = -2,       /* CUDA runtime failure (no device, OOM, launch error)           */
  CHB_ERR_STATE = -3,      /* call order: e.g. chb_eval before chb_set_events               */
  CHB_ERR_UNSUPPORTED = -4 /* option combination the reference does not support either      */
} chb_status;

/* model ids: population/cosmo.py:50-115, mass.py:56-189, rate.py:32-88 */
enum { CHB_COSMO_FLRW = 0, CHB_COSMO_MG_FLRW = 1 };
enum { CHB_MASS_TPL = 0, CHB_MASS_BPL = 1, CHB_MASS_PLP = 2 };
enum { CHB_RATE_POWER_LAW = 0, CHB_RATE_MADAU_DICKINSON = 1, CHB_RATE_TRUNC_MD = 2, CHB_RATE_TRUNC_PL = 3 };
/* likelihood.py:48-62 options */
enum { CHB_KERNEL_EPAN = 0, CHB_KERNEL_GAUSS = 1 };
enum { CHB_BW_SCOTT = 0, CHB_BW_SILVERMAN = 1, CHB_BW_SCALAR = 2 };
enum { CHB_PGW_1D = 0, CHB_PGW_APPROX = 1, CHB_PGW_MARG = 2, CHB_PGW_FULL = 3 };
/* arithmetic of the KDE pair sum: fp64 everywhere, or fp32 pair sums (MUFU ex2) with fp64
 * reweighting/tables/integration.  north_star: 1e-5 relative in fp64 mode, 1e-3 in fp32. */
enum { CHB_FP64 = 0, CHB_FP32 = 1 };

/* One hyper-point = one row of CHB_NPAR doubles, fixed column order.  The Python shim routes
 * keyword hyper-parameters to these slots (pop_wrapper.py:56-64). */
#define CHB_NPAR 32
enum {
  CHB_P_H0 = 0, CHB_P_OM0, CHB_P_OK0, CHB_P_OR0, CHB_P_W0, CHB_P_WA, CHB_P_XI0, CHB_P_N, CHB_P_ZMAX, /* 0..8 */
  CHB_P_MLOW = 9, CHB_P_MHIGH, CHB_P_ALPHA /* tpl/plp alpha, bpl alpha_1 */, CHB_P_BETA, CHB_P_DELTAM,
  CHB_P_ALPHA2, CHB_P_BREAKF, CHB_P_LAMBDAP, CHB_P_MUG, CHB_P_SIGMAG,                                 /* 9..18 */
  CHB_P_GAMMA = 21, CHB_P_KAPPA, CHB_P_ZP, CHB_P_RZMAX,                                               /* 21..24 */
  CHB_P_R0 = 25,
  /* first / last knot of m_grid as the caller's libm evaluates 10**log10(m_low|m_high); the support
   * tests `m_low <= m <= m_high` at those two knots (mass.py:240-245) sit on a rounding knife-edge,
   * so they are taken verbatim when non-zero (0: computed on the device). */
  CHB_P_MGRID_FIRST = 26, CHB_P_MGRID_LAST = 27
};

typedef struct {
  int32_t abi_version;     /* must be CHB_ABI_VERSION                                         */
  int32_t device;          /* CUDA device ordinal                                             */
  int32_t fp_mode;         /* CHB_FP64 | CHB_FP32                                             */
  int32_t cosmo_model, mass_model, rate_model;
  int32_t cosmo_grid_res;  /* z_grid_res of the cosmology table (cosmo.py:77: 1500)           */
  int32_t mass_grid_res;   /* grid_res of the mass tables      (mass.py:20:  1000)            */
  /* hyperlikelihood.__init__ (likelihood.py:48-62) */
  int32_t kind_p_gw;       /* CHB_PGW_*; CHB_PGW_1D = non-pixelated                           */
  int32_t kernel;          /* CHB_KERNEL_*                                                    */
  int32_t bw_method;       /* CHB_BW_*                                                        */
  double  bw_value;        /* scalar factor when bw_method == CHB_BW_SCALAR                   */
  int32_t use_cut_grid;    /* 0: cut_grid=None (evaluate KDE on z_grids directly)             */
  double  cut_grid;
  int32_t binning;
  int32_t num_bins;
  double  pe_neff;
  /* population (pop_wrapper.py:23-43) */
  int32_t scale_free;
  double  Tobs;
  /* catalogue: 0 = empty_catalog(p_bkg='dVdz') (catalog.py:19-43); 1 = pixelated_catalog with
   * dVdz_completeness(z_range) (catalog.py:197-203, completeness.py:43-67) */
  int32_t catalog_kind;
  double  compl_z_lo, compl_z_hi;
  /* selection_function.__init__ (selection_function.py:24-32) */
  double  N_inj;
  int32_t check_neff;      /* 0: N_eff=None                                                   */
  double  N_eff;
} chb_config;

typedef struct chb_handle chb_handle;

/* library / device probes (no reference counterpart) */
int chb_abi_version(void);
int chb_device_count(void);
const char* chb_last_error(const chb_handle* h);      /* h may be NULL: last create error    */

/* hyperlikelihood(...) + selection_function(...) + population(...) construction */
int chb_create(chb_handle** out, const chb_config* cfg);
void chb_destroy(chb_handle* h);

/* Per-handle tuning / diagnostic switches.  They select HOW the same numbers are computed (every setting is held to the
 * same parity tests), never what; the library reads no environment variables.  No reference counterpart.
 *   "fused"      1 (default): fp32 mode, non-pixelated / 'approximate' kinds run as ONE kernel per step with the
 *                reweighted samples in shared memory (csrc/numerator_fused.cu); 0: reweighting kernel -> stage buffer in
 *                global memory -> KDE kernel (csrc/numerator_f32.cu, the form the other kinds use)
 *   "split"      1 (default) | 0: non-fused form as two kernels | as one MODE-0 kernel
 *   "kde_win"    32 (default): windowed Gaussian recurrence, sub-stream iterations per chunk; 0: every sample visits
 *                the whole effective grid
 *   "kde_win_t2" 24 (default): window threshold of the fused kernel in bits -- terms below 2^-t2 of the LARGEST term at a grid
 *                point are dropped (fp32 carries 24 bits); 30 was round 1's setting
 *   "kde_direct" 0 (default) | 1: one MUFU.EX2 per (grid point, sample) pair, no recurrence
 *   "bin_runs"   1 (default) | 0: non-fused binning by runs of sorted samples | one shared-memory atomic per sample
 *   "epan_blocks" 1 (default) | 0: fused kernel, unbinned Epanechnikov KDE by block moments of the sorted samples | direct pair sums
 *   "fused_nt"   0 (default) | 64 | 128 | 256: threads per CTA of the fused 1-D kernel; 0 picks 64 (twelve CTAs per SM) for events
 *                with <= 1024 samples and 128 (six) up to 2048, where the per-unit overheads dominate, else 256 (three)
 *   "zterms_gb"  16 (default): budget of the buffer of precomputed z-grid terms (n_hyper x Nev x Nz x 8 B); the one-launch
 *                kernels take the hyper-points in batches beyond it (large walker batches)
 *   "stage_gb"   12 (default): budget of the stage buffer of the non-fused form; hyper-points are batched beyond it
 * Unknown names / out-of-range values: CHB_ERR_INVALID. */
int chb_set_option(chb_handle* h, const char* name, double value);

/* theta_pe_det fields (data.py:27-47) + z_grids (likelihood.py:51).  Arrays are (Nev, Ns)
 * row-major, z_grids (Nev, Nz).  ra/dec may be NULL unless kind_p_gw == CHB_PGW_FULL. */
int chb_set_events(chb_handle* h, int64_t Nev, int64_t Ns, int64_t Nz,
                   const double* m1det, const double* m2det, const double* dL, const double* pe_prior,
                   const double* ra, const double* dec, const double* z_grids);

/* pixelised fields of theta_pe_det (data.py:37-43), padded with -100 (data.py:348-351).
 * pixels_opt_nsides (Nev,P) int64; pixels_pe_opt_nside (Nev,Ns) int64; ra_pix, dec_pix,
 * gw_loc2d_pdf (Nev,P).  neff_pixels is derived as count(ra_pix != -100) (catalog.py:121). */
int chb_set_pixels(chb_handle* h, int64_t P, const int64_t* pixels_opt_nsides,
                   const int64_t* pixels_pe_opt_nside, const double* ra_pix, const double* dec_pix,
                   const double* gw_loc2d_pdf);

/* pixelated_catalog.p_cat (Nev,P,Nz) with -100 padding and P_compl (Nev,Nz)
 * (catalog.py:180-195). */
int chb_set_catalog(chb_handle* h, const double* p_cat, const double* P_compl);

/* theta_inj_det (data.py:49-53): four (Ninj,) arrays. */
int chb_set_injections(chb_handle* h, int64_t Ninj, const double* m1det, const double* m2det,
                       const double* dL, const double* p_draw);

/* hyperlikelihood.compute_all for a batch of hyper-points (likelihood.py:326-338), split at the
 * cross-rank reduction:
 *   hyper         (n_hyper, CHB_NPAR) host
 *   log_like_evs  (n_hyper, Nev) host or NULL   -- nan_to_num(log(like_ev))  (likelihood.py:328-329)
 *   partials      (n_hyper, 3)   host           -- [sum_ev log_like_evs, sum_inj w, sum_inj w^2]
 *                                                  with w = dN/p_draw (selection_function.py:37-44)
 *   p_gw          optional host output of p_gw1d (n_hyper,Nev,Nz) or p_gw3d (n_hyper,Nev,P,Nz)
 *                 (likelihood.py:105-260); NULL to skip.
 * Host pointers, synchronous: includes the H2D copy of `hyper` and the D2H copies of the
 * outputs.  Events/injections may be absent: the matching partial columns are then 0. */
int chb_eval(chb_handle* h, int64_t n_hyper, const double* hyper,
             double* log_like_evs, double* partials, double* p_gw);
/* compute_numlike_evs (likelihood.py:266-292): the per-event integrals BEFORE log/nan_to_num of
 * the most recent chb_eval* call, (n_hyper, Nev) host. */
int chb_last_numlike_evs(chb_handle* h, double* like_evs);

/* Same, with DEVICE pointers (e.g. torch tensors), asynchronous on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = default stream). */
int chb_eval_device(chb_handle* h, int64_t n_hyper, const double* d_hyper,
                    double* d_log_like_evs, double* d_partials, double* d_p_gw, void* cuda_stream);

/* Host epilogue after the (optional) cross-rank all-reduce of `partials`:
 * N_eff gate + N_exp (selection_function.py:41-47) and log_num - Nev*log(N_exp) or
 * log_num + Nev*log(R0*Tobs) - N_exp (likelihood.py:299-316).  Nev_total = global event count.
 * Outputs (n_hyper,) each; any may be NULL.  N_exp is selection_function.N_exp itself. */
int chb_finalize(const chb_config* cfg, int64_t n_hyper, int64_t Nev_total, const double* hyper,
                 const double* partials, double* log_like_num, double* log_Nexp, double* log_hyper,
                 double* neff_inj, double* N_exp);

/* Plug-in free functions evaluated on the device for ONE parameter row (element-wise, host
 * pointers, n elements).  `which` selects the function; `a`, `b` are its array arguments
 * (b may be NULL).  Stand-ins for the plum-dispatched functions of population/cosmo.py
 * (:122-264), mass.py (:334-345) and rate.py (:96-129). */
enum {
  CHB_F_E_AT_Z = 0,       /* E_at_z(cosmo, a=z)                                  */
  CHB_F_DL_AT_Z,          /* dL_at_z(cosmo, a=z)                                 */
  CHB_F_Z_FROM_DGW,       /* z_from_dGW(cosmo, a=dL)                             */
  CHB_F_DDLDZ_AT_Z,       /* ddLdz_at_z(cosmo, a=z [, b=distances])              */
  CHB_F_DVCDZ_AT_Z,       /* dVcdz_at_z(cosmo, a=z [, b=distances])              */
  CHB_F_VC_AT_Z,          /* Vc_at_z(cosmo, a=z [, b=distances])                 */
  CHB_F_DCT_AT_Z,         /* dCt_at_z(cosmo, a=z)                                */
  CHB_F_P_M1M2,           /* p_m1m2(mass, a=m1, b=m2)                            */
  CHB_F_P_M1_NOTNORM,     /* primary_mass_pdf_notnorm(mass, a=m)                 */
  CHB_F_MERGER_RATE,      /* merger_rate(rate, a=z)                              */
  CHB_F_POP_RATE_DET_INJ  /* pop_rate_det(pop, theta_inj_det): a=m1det,b=m2det + extra */
};
int chb_model_eval(const chb_config* cfg, int which, const double* params /* CHB_NPAR */,
                   int64_t n, const double* a, const double* b, const double* c, double* out);

/* The per-hyper-point interpolation tables themselves (cosmo.py:43-46, mass.py:45-52):
 * z_grid_interp, integral_invE_interp (cosmo_grid_res each), m_grid, cdf_m2_conditioned
 * (mass_grid_res each), norm_p_m1 (1).  Any output may be NULL. */
int chb_model_tables(const chb_config* cfg, const double* params, double* z_grid_interp,
                     double* integral_invE_interp, double* m_grid, double* cdf_m2_conditioned,
                     double* norm_p_m1);

/* Introspection for benchmarks/tests: launches of this library's kernels since creation. */
int64_t chb_kernel_launch_count(const chb_handle* h);
/* Device time (ms, CUDA events on the launching stream) of the last chb_eval*'s kernels:
 * out[0]=tables, out[1]=numerator incl. z-grid terms, out[2]=selection, out[3]=reduce,
 * out[4]=z-grid terms alone, out[5]=reweight+KDE+z-integral kernel(s) alone, out[6..7]=0 (reserved) */
int chb_last_timings(const chb_handle* h, double out[8]);

/* Phase profile of the fused numerator kernel: mean SM-clock cycles per CTA spent in
 * [0] table staging, [1] z-grid terms, [2] reweighting, [3] statistics/grid, [4] KDE+integrand,
 * [5] final reduction, for the most recent evaluation run with profiling enabled.  `enable`
 * switches collection on/off for the following evaluations; `out` may be NULL. */
int chb_phase_profile(chb_handle* h, int enable, double out[8]);

/* Measured MUFU.EX2 throughput [exp/s] of `device` (micro-benchmark run for ~`seconds`): the
 * denominator of the KDE roofline fraction. */
int chb_mufu_peak(int device, double seconds, double* exp_per_s);

/* ---- Setup-side callers of the path, stateless, host pointers ------------------
 * HEALPix RING indexing, replacing healpy.ang2pix / healpy.pix2ang as called at CHIMERA/utils/angles.py:45,58,71
 * (nest=False): theta = colatitude in [0, pi], phi = longitude [rad]; pix int64 in [0, 12 nside^2).
 * nside must be a power of two (ValueError otherwise, like healpy). */
int chb_healpix_ang2pix_ring(int device, int64_t nside, int64_t n, const double* theta, const double* phi,
                             int64_t* pix);
int chb_healpix_pix2ang_ring(int device, int64_t nside, int64_t n, const int64_t* pix, double* theta, double* phi);

/* The per-sample part of pixelize_gw_catalog (CHIMERA/data.py:316-345) for events whose pixel sets are known:
 * pixels_pe_opt_nside (Nev x Ns): the sample's own pixel at opt_nsides[ev] if it is one of the event's pixels,
 *   else the event pixel with the smallest angular separation (utils/angles.py:146-160, numpy.argmin rule);
 * gw_loc2d_pdf (Nev x P): 2-D Gaussian KDE of the (ra, dec) samples at the pixel centres
 *   (jax_gkde_nd, utils/math.py:95-148), -100 in padded slots.
 * Inputs: ra/dec (Nev x Ns) [rad]; pixels_opt_nsides, ra_pix, dec_pix (Nev x P) padded with -100.
 * Either output may be NULL. */
int chb_pixelize_samples(int device, int64_t Nev, int64_t Ns, int64_t P, const int64_t* opt_nsides,
                         const double* ra, const double* dec, const int64_t* pixels_opt_nsides,
                         const double* ra_pix, const double* dec_pix, int64_t* pixels_pe_opt_nside,
                         double* gw_loc2d_pdf);

/* pixelated_catalog.precompute_p_cat (CHIMERA/catalog/catalog.py:143-195, _sum_gaussians_ucv :209-221):
 * p_cat (Nev x P x Nz) = per (event, pixel) sum over the pixel's galaxies with z strictly inside the event grid
 * of w N(z_k; z_gal, z_err) dVdz[ev,k] / trapz_k(...) / sum w; non-finite -> 0; padded pixel slots -100.
 * dVdz (Nev x Nz) is dVc/dz of the catalogue's fiducial cosmology on z_grids (chb_model_eval, CHB_F_DVCDZ_AT_Z).
 * Galaxies (Ngal): ra, dec [rad], z, z_err (= z_err (1+z), catalog.py:113), w.  N_gal (Nev) may be NULL. */
int chb_precompute_p_cat(int device, int64_t Nev, int64_t P, int64_t Nz, const double* z_grids, const double* dVdz,
                         const int64_t* opt_nsides, const int64_t* pixels_opt_nsides, const int32_t* neff_pixels,
                         int64_t Ngal, const double* gal_ra, const double* gal_dec, const double* gal_z,
                         const double* gal_zerr, const double* gal_w, double* p_cat, double* N_gal);

#ifdef __cplusplus
}
#endif
#endif /* CHIMERA_B200_H */
