// common.cuh -- argument blocks and launcher prototypes shared by the kernels and the C ABI.
#pragma once
#include "models.cuh"
#include "devguard.cuh"

struct ModelCfg {
  int cosmo_model, mass_model, rate_model;
  int catalog_kind;
  double compl_z_lo, compl_z_hi;
  TableLayout lay;
};

// Everything the fused numerator kernel needs (device pointers).
struct NumArgs {
  ModelCfg mc;
  // hyperlikelihood options (likelihood.py:48-62)
  int kind, kernel, bw_method, use_cut, binning, num_bins, fp_mode;
  int rec_off;           // 1: disable the Gaussian recurrence (one MUFU.EX2 per pair), env CHB_KDE_DIRECT=1
  int bin_runs;          // 1: binning1d by contiguous runs of the sorted samples (default; env CHB_BIN_RUNS=0 restores one atomic per sample)
  int epan_blocks;       // 1 (default): unbinned Epanechnikov KDE by block moments of the sorted samples (option epan_blocks; 0: direct pair sums)
  float win_t2;          // window threshold of the windowed KDE in bits (option kde_win_t2): terms below 2^-t2 of the largest term at a grid point are dropped
  int kde_win_iters;     // > 0: windowed recurrence over sorted samples, sub-stream iterations per chunk (env CHB_KDE_WIN, default 32; 0 = off)
  double bw_value, cut_grid, pe_neff;
  // event data, samples permuted so that each pixel's samples are contiguous
  int Nev, Ns, Nz, P;
  const double *m1d, *m2d, *dL, *prior, *ra, *dec;   // (Nev, Ns)
  // fp32 mode: packed single-precision copies {dL, m1det, m2det, 1/pe_prior} and {log2 m1det, log2 m2det}
  const float4* s4;
  const float2* l2;
  const double* zgrids;                              // (Nev, Nz)
  const int* pix_off;                                // (Nev, P+2) sample offsets per pixel slot
  const double *ra_pix, *dec_pix, *gw_pdf;           // (Nev, P)
  const int* neff_pix;                               // (Nev,)
  const double* p_cat;                               // (Nev, P, Nz)
  const double* P_compl;                             // (Nev, Nz)
  // 'approximate' kind: hyper-independent pixel sums A[ev,k] = sum_p gw_pdf[p] p_cat[p,k] and
  // B[ev,k] = sum_p gw_pdf[p] [p_cat[p,k] != -100] over the event's valid pixels (catalog_collapse_kernel)
  const double* catA;
  const double* catB;
  // hyper-points
  int n_hyper;
  const double* hyper;   // (n_hyper, CHB_NPAR)
  const double* tabs;    // (n_hyper, lay.total())
  const double* HC;      // (n_hyper, CHB_NHC)
  // outputs
  double* log_like;      // (n_hyper, Nev)
  double* like_raw;      // (n_hyper, Nev) integral before log / nan_to_num
  double* p_gw_out;      // optional (n_hyper, Nev, [P,] Nz)
  // fp32 fast path: precomputed z-grid terms {dVc/dz, ck} for hyper-points [zterms_h0, zterms_h0 + n)
  const float2* zterms;
  // split fast path (numerator_f32_kernel MODE 1 -> MODE 2): reweighted samples {z, w} (units x Ns) and the unit
  // statistics {sum w, sum w^2, min z, max z, std z, -, -, -} (units x 8), unit = ev * n_hyper + h
  float2* zw_stage;
  double* unit_stats;
  float2* zterms_out;
  int zterms_h0;
  // optional phase profile: (gridDim, 8) SM-clock cycles accumulated by thread 0 of every CTA
  unsigned long long* prof;
  // per-CTA global scratch for sample staging when it does not fit in shared memory
  double* scratch;
  long long scratch_stride;   // doubles per CTA (0: staging lives in shared memory)
};

struct SelArgs {
  ModelCfg mc;
  int Ninj, n_hyper, tiles;
  const double *m1d, *m2d, *dL, *p_draw;
  int fp_mode;
  const float4* s4;      // fp32 mode: {dL, m1det, m2det, 1/p_draw}
  const float2* l2;      //            {log2 m1det, log2 m2det}
  const double *hyper, *tabs, *HC;
  double* tile_part;     // (n_hyper, tiles, 2): nansum(w), sum(w^2) per tile
};

// launchers (each returns the cudaError_t of the launch)
cudaError_t launch_build_tables(const ModelCfg& mc, int n_hyper, const double* d_hyper, double* d_tabs,
                                double* d_HC, cudaStream_t s);
cudaError_t launch_selection(const SelArgs& a, cudaStream_t s);
int selection_ctas_per_sm(const ModelCfg& mc, int fp_mode);      // co-resident CTAs of the selection kernel (0: query failed)
cudaError_t launch_numerator(const NumArgs& a, int grid, int block, size_t smem, cudaStream_t s);
size_t numerator_smem_bytes(const NumArgs& a, bool stage_in_smem);
long long numerator_scratch_doubles(const NumArgs& a);
int numerator_block_threads();
cudaError_t numerator_configure(size_t smem);
size_t numerator_f32_smem_bytes(const NumArgs& a, int mode);
cudaError_t numerator_f32_configure(int kind, int mode, size_t smem);
int numerator_f32_ctas_per_sm(int kind, int mode, size_t smem);
cudaError_t launch_numerator_f32(const NumArgs& a, int mode, int grid, size_t smem, cudaStream_t s);
size_t numerator_fused_smem_bytes(const NumArgs& a);
bool numerator_fused_supported(const NumArgs& a);
cudaError_t numerator_fused_configure(size_t smem);
int numerator_fused_ctas_per_sm(size_t smem);
cudaError_t launch_numerator_fused(const NumArgs& a, int grid, size_t smem, cudaStream_t s);
// ... and its 128-thread instantiation (numerator_fused_nt128.cu)
size_t numerator_fused_smem_bytes_nt128(const NumArgs& a);
bool numerator_fused_supported_nt128(const NumArgs& a);
cudaError_t numerator_fused_configure_nt128(size_t smem);
int numerator_fused_ctas_per_sm_nt128(size_t smem);
cudaError_t launch_numerator_fused_nt128(const NumArgs& a, int grid, size_t smem, cudaStream_t s);
size_t numerator_fused_smem_bytes_nt64(const NumArgs& a);
bool numerator_fused_supported_nt64(const NumArgs& a);
cudaError_t numerator_fused_configure_nt64(size_t smem);
int numerator_fused_ctas_per_sm_nt64(size_t smem);
cudaError_t launch_numerator_fused_nt64(const NumArgs& a, int grid, size_t smem, cudaStream_t s);
size_t numerator_marg_smem_bytes(const NumArgs& a);
bool numerator_marg_supported(const NumArgs& a);
cudaError_t numerator_marg_configure(size_t optin);
int numerator_marg_ctas_per_sm(size_t smem);
cudaError_t launch_numerator_marg(const NumArgs& a, int grid, size_t smem, cudaStream_t s);
cudaError_t launch_zgrid_terms(const NumArgs& a, int h0, int nh, cudaStream_t s);
cudaError_t launch_catalog_collapse(int Nev, int P, int Nz, const double* p_cat, const double* gw_pdf, const int* neff_pix,
                                   double* catA, double* catB, cudaStream_t s);
cudaError_t launch_reduce(int n_hyper, int Nev, int tiles, const double* d_log_like, const double* d_tile_part,
                          double* d_partials, cudaStream_t s);
cudaError_t launch_model_eval(const ModelCfg& mc, int which, const double* d_params, const double* d_tabs,
                              const double* d_HC, long long n, const double* a, const double* b, const double* c,
                              double* out, cudaStream_t s);
