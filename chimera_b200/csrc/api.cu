// api.cu -- C ABI of libchimera_b200.so (see include/chimera_b200.h).
// Host-side runtime: device memory ownership, one-time uploads, pixel bucketing of the
// posterior samples, kernel sequencing on the handle's stream, and the host epilogue.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count; else p = nullptr;
    return e;
  }
  cudaError_t upload(const T* src, size_t count) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct chb_handle {
  chb_config cfg;
  ModelCfg mc;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evz = nullptr;               // after the z-grid terms (start of the numerator kernels proper)
  int sm_count = 148;
  int max_smem_optin = 0;
  int64_t launches = 0;
  double timings[4] = {0, 0, 0, 0};

  // events
  int64_t Nev = 0, Ns = 0, Nz = 0, P = 0, Ninj = 0;
  bool have_events = false, have_pixels = false, have_catalog = false, have_inj = false, dirty = true;
  bool sorted = false;      // fp32 1-D kinds: samples of every event sorted by dL (windowed KDE, kde_f32.cuh)
  std::vector<double> h_m1, h_m2, h_dL, h_prior, h_ra, h_dec;      // host copies until prepare()
  std::vector<int64_t> h_pixels, h_pe_pix;
  std::vector<double> h_ra_pix;
  DevBuf<double> m1d, m2d, dL, prior, ra, dec, zgrids, ra_pix, dec_pix, gw_pdf, p_cat, P_compl;
  DevBuf<int> pix_off, neff_pix;
  DevBuf<unsigned long long> prof;
  DevBuf<float4> s4;
  DevBuf<float2> l2;
  DevBuf<float2> zterms;
  DevBuf<float2> zw_stage;
  DevBuf<double> unit_stats;
  int64_t plan_nh = -1, plan_nb = 0;       // split-form plan, valid for plan_nh hyper-points (reset by chb_set_* / options)
  int plan_per1 = 0, plan_per2 = 0;
  bool plan_ok = false;
  int fused_per = -1;                      // co-resident CTAs per SM of numerator_fused_kernel for the current shapes
  int marg_per = -1;                       // likewise numerator_marg_kernel ('marginalized' + binning)
  // tuning / diagnostic options (chb_set_option); they never change what is computed, only how
  int opt_fused = 1;                       // 1-D kinds, fp32: one fused kernel (numerator_fused.cu); 0: round-1 split kernels
  int opt_split = 1;                       // round-1 path: split MODE 1 -> stage -> MODE 2 (0: one MODE 0 kernel)
  int opt_kde_win = 32;                    // windowed recurrence: sub-stream iterations per chunk (0: windows off)
  double opt_kde_win_t2 = 24.0;            // window threshold in bits (fused kernel)
  int opt_kde_direct = 0;                  // 1: one MUFU.EX2 per pair (no recurrence)
  int opt_bin_runs = 1;                    // round-1 path: binning by runs of the sorted samples
  int opt_epan_blocks = 1;                 // fused kernel, unbinned Epanechnikov: block moments (0: direct pair sums)
  int opt_fused_nt = 0;                    // threads per CTA of the fused 1-D kernel: 0 = 64 for Ns <= 1024, 128 for Ns <= 2048, else 256
  double opt_zterms_gb = 16.0;             // budget of the z-grid-term buffer; the one-launch kernels batch the hyper-points beyond it
  double opt_stage_gb = 12.0;              // round-1 path: budget of the {z, w} stage buffer
  DevBuf<double> catA, catB;
  bool cat_collapsed = false;
  bool want_prof = false;
  DevBuf<double> inj_m1, inj_m2, inj_dL, inj_pd;
  DevBuf<float4> inj_s4;
  DevBuf<float2> inj_l2;
  // per-eval buffers
  DevBuf<double> hyper, tabs, HC, log_like, like_raw, tile_part, partials, pgw, scratch;
  int64_t last_n_hyper = 0;
  int num_grid = 0;
  size_t num_smem = 0;
  bool stage_in_smem = true;
};

static int fail(chb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
static int cuda_fail(chb_handle* h, cudaError_t e, const char* what) {
  return fail(h, CHB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int chb_fail_global(int code, const char* msg) { return fail(nullptr, code, msg ? msg : ""); }   // setup.cu
#define CU(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(h, _e, what); } while (0)

static int validate_cfg(const chb_config* c, std::string& why) {
  if (!c) { why = "cfg is NULL"; return CHB_ERR_INVALID; }
  if (c->abi_version != CHB_ABI_VERSION) { why = "abi_version mismatch"; return CHB_ERR_INVALID; }
  if (c->cosmo_model < 0 || c->cosmo_model > CHB_COSMO_MG_FLRW) { why = "unknown cosmo_model"; return CHB_ERR_INVALID; }
  if (c->mass_model < 0 || c->mass_model > CHB_MASS_PLP) { why = "unknown mass_model"; return CHB_ERR_INVALID; }
  if (c->rate_model < 0 || c->rate_model > CHB_RATE_TRUNC_PL) { why = "unknown rate_model"; return CHB_ERR_INVALID; }
  if (c->cosmo_grid_res < 4 || c->mass_grid_res < 4) { why = "table resolution too small"; return CHB_ERR_INVALID; }
  if (c->kind_p_gw < CHB_PGW_1D || c->kind_p_gw > CHB_PGW_FULL) {
    why = "`kind_p_gw3d` must be one of `approximate`, `marginalized`, or `full`"; return CHB_ERR_INVALID; }
  if (c->kernel != CHB_KERNEL_EPAN && c->kernel != CHB_KERNEL_GAUSS) { why = "unknown kernel"; return CHB_ERR_INVALID; }
  if (c->bw_method < CHB_BW_SCOTT || c->bw_method > CHB_BW_SCALAR) {
    why = "bw_method should be 'scott', 'silverman', or a scalar"; return CHB_ERR_INVALID; }
  if (c->binning && c->num_bins < 1) { why = "num_bins must be positive"; return CHB_ERR_INVALID; }
  if (c->kind_p_gw == CHB_PGW_FULL && !c->use_cut_grid) {
    why = "kind_p_gw3d='full' needs a numeric cut_grid"; return CHB_ERR_UNSUPPORTED; }
  if (c->fp_mode != CHB_FP64 && c->fp_mode != CHB_FP32) { why = "unknown fp_mode"; return CHB_ERR_INVALID; }
  return CHB_OK;
}

static ModelCfg model_cfg(const chb_config& c) {
  ModelCfg m;
  m.cosmo_model = c.cosmo_model; m.mass_model = c.mass_model; m.rate_model = c.rate_model;
  m.catalog_kind = c.catalog_kind; m.compl_z_lo = c.compl_z_lo; m.compl_z_hi = c.compl_z_hi;
  m.lay = make_layout(c.cosmo_grid_res, c.mass_grid_res);
  return m;
}

extern "C" {

int chb_abi_version(void) { return CHB_ABI_VERSION; }

int chb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* chb_last_error(const chb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int chb_create(chb_handle** out, const chb_config* cfg) {
  if (!out) return fail(nullptr, CHB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  std::string why;
  int rc = validate_cfg(cfg, why);
  if (rc != CHB_OK) return fail(nullptr, rc, why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, CHB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CHB_ERR_INVALID, "device ordinal out of range");
  DevGuard _dg;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  chb_handle* h = new chb_handle();
  h->cfg = *cfg;
  h->mc = model_cfg(*cfg);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete h; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
  for (auto& evn : h->ev) cudaEventCreate(&evn);
  cudaEventCreate(&h->evz);
  *out = h;
  return CHB_OK;
}

void chb_destroy(chb_handle* h) {
  if (!h) return;
  DevGuard _dg(h->cfg.device);
  DevBuf<double>* dbl[] = {&h->m1d, &h->m2d, &h->dL, &h->prior, &h->ra, &h->dec, &h->zgrids, &h->ra_pix, &h->dec_pix,
                           &h->gw_pdf, &h->p_cat, &h->P_compl, &h->inj_m1, &h->inj_m2, &h->inj_dL, &h->inj_pd,
                           &h->hyper, &h->tabs, &h->HC, &h->log_like, &h->like_raw, &h->tile_part, &h->partials, &h->pgw, &h->scratch};
  for (auto* b : dbl) b->release();
  h->pix_off.release();
  h->prof.release();
  h->s4.release(); h->l2.release(); h->inj_s4.release(); h->inj_l2.release(); h->zterms.release(); h->zw_stage.release(); h->unit_stats.release(); h->catA.release(); h->catB.release();
  h->neff_pix.release();
  for (auto& evn : h->ev) if (evn) cudaEventDestroy(evn);
  if (h->evz) cudaEventDestroy(h->evz);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int chb_set_events(chb_handle* h, int64_t Nev, int64_t Ns, int64_t Nz, const double* m1det, const double* m2det,
                   const double* dL, const double* pe_prior, const double* ra, const double* dec,
                   const double* z_grids) {
  if (!h) return CHB_ERR_INVALID;
  if (Nev < 1 || Ns < 1 || Nz < 2) return fail(h, CHB_ERR_INVALID, "need Nev>=1, Ns>=1, Nz>=2");
  if (Ns > (1 << 30) / 4 || Nev * Ns > (int64_t)1 << 40) return fail(h, CHB_ERR_INVALID, "event arrays too large");
  if (!m1det || !m2det || !dL || !pe_prior || !z_grids) return fail(h, CHB_ERR_INVALID, "NULL event array");
  if (h->cfg.kind_p_gw == CHB_PGW_FULL && (!ra || !dec)) return fail(h, CHB_ERR_INVALID, "kind 'full' needs ra/dec samples");
  if (h->cfg.use_cut_grid && Nz / 2 < 2) return fail(h, CHB_ERR_INVALID, "z_int_res//2 must be >= 2");
  DevGuard _dg(h->cfg.device);
  const size_t n = (size_t)Nev * Ns;
  h->Nev = Nev; h->Ns = Ns; h->Nz = Nz;
  h->h_m1.assign(m1det, m1det + n); h->h_m2.assign(m2det, m2det + n);
  h->h_dL.assign(dL, dL + n); h->h_prior.assign(pe_prior, pe_prior + n);
  if (ra && dec) { h->h_ra.assign(ra, ra + n); h->h_dec.assign(dec, dec + n); } else { h->h_ra.clear(); h->h_dec.clear(); }
  CU(h->zgrids.upload(z_grids, (size_t)Nev * Nz), "upload z_grids");
  h->have_events = true; h->have_pixels = false; h->have_catalog = false; h->dirty = true;
  h->plan_nh = -1; h->fused_per = -1; h->marg_per = -1;
  return CHB_OK;
}

int chb_set_pixels(chb_handle* h, int64_t P, const int64_t* pixels_opt_nsides, const int64_t* pixels_pe_opt_nside,
                   const double* ra_pix, const double* dec_pix, const double* gw_loc2d_pdf) {
  if (!h) return CHB_ERR_INVALID;
  if (!h->have_events) return fail(h, CHB_ERR_STATE, "chb_set_events must come first");
  if (P < 1 || !pixels_opt_nsides || !ra_pix || !dec_pix || !gw_loc2d_pdf)
    return fail(h, CHB_ERR_INVALID, "NULL pixel array or P < 1");
  if (h->cfg.kind_p_gw == CHB_PGW_MARG && !pixels_pe_opt_nside)
    return fail(h, CHB_ERR_INVALID, "kind 'marginalized' needs pixels_pe_opt_nside");
  DevGuard _dg(h->cfg.device);
  h->P = P;
  const size_t np = (size_t)h->Nev * P;
  h->h_pixels.assign(pixels_opt_nsides, pixels_opt_nsides + np);
  if (pixels_pe_opt_nside) h->h_pe_pix.assign(pixels_pe_opt_nside, pixels_pe_opt_nside + (size_t)h->Nev * h->Ns);
  else h->h_pe_pix.clear();
  h->h_ra_pix.assign(ra_pix, ra_pix + np);
  CU(h->ra_pix.upload(ra_pix, np), "upload ra_pix");
  CU(h->dec_pix.upload(dec_pix, np), "upload dec_pix");
  CU(h->gw_pdf.upload(gw_loc2d_pdf, np), "upload gw_loc2d_pdf");
  // neff_pixels = count(ra_pix != -100)  (catalog.py:121)
  std::vector<int> neff(h->Nev);
  for (int64_t e = 0; e < h->Nev; ++e) {
    int c = 0;
    for (int64_t p = 0; p < P; ++p) c += (ra_pix[e * P + p] != -100.0);
    neff[e] = c;
  }
  CU(h->neff_pix.upload(neff.data(), neff.size()), "upload neff_pixels");
  h->have_pixels = true; h->dirty = true;
  h->plan_nh = -1; h->fused_per = -1; h->marg_per = -1;
  return CHB_OK;
}

int chb_set_catalog(chb_handle* h, const double* p_cat, const double* P_compl) {
  if (!h) return CHB_ERR_INVALID;
  if (!h->have_pixels) return fail(h, CHB_ERR_STATE, "chb_set_pixels must come first");
  if (!p_cat || !P_compl) return fail(h, CHB_ERR_INVALID, "NULL catalogue array");
  if (h->cfg.catalog_kind != 1) return fail(h, CHB_ERR_INVALID, "config has catalog_kind=0 (empty catalogue)");
  DevGuard _dg(h->cfg.device);
  CU(h->p_cat.upload(p_cat, (size_t)h->Nev * h->P * h->Nz), "upload p_cat");
  CU(h->P_compl.upload(P_compl, (size_t)h->Nev * h->Nz), "upload P_compl");
  h->have_catalog = true;
  h->cat_collapsed = false;
  return CHB_OK;
}

int chb_set_injections(chb_handle* h, int64_t Ninj, const double* m1det, const double* m2det, const double* dL,
                       const double* p_draw) {
  if (!h) return CHB_ERR_INVALID;
  if (Ninj < 1 || Ninj > 0x7fffffff || !m1det || !m2det || !dL || !p_draw)
    return fail(h, CHB_ERR_INVALID, "bad injection arrays");
  DevGuard _dg(h->cfg.device);
  // The two sums over the injections do not depend on their order (selection_function.py:37-44), so the injections are
  // stored sorted by an estimate of the source-frame secondary mass, m2det / (1 + z_fid(dL)) with z_fid from a fixed flat
  // LCDM (H0 = 70, Om0 = 0.3): the 32 injections a warp evaluates together then lie on the same side of the low-mass
  // taper m_low + delta_m, and the warps above it skip the taper's eight MUFU operations per injection (selection.cu).
  std::vector<int> perm((size_t)Ninj);
  {
    const int NT = 4096;                                 // dL(z) of the fiducial cosmology on a log grid of z in [1e-4, 100]
    std::vector<double> tz(NT), td(NT);
    double dc = 0.0, zp = 0.0;
    for (int i = 0; i < NT; ++i) {
      const double z = std::pow(10.0, -4.0 + 6.0 * i / (NT - 1));
      const int sub = 8;
      for (int k = 0; k < sub; ++k) {                    // midpoint rule for int dz / E
        const double zm = zp + (z - zp) * (k + 0.5) / sub;
        dc += (z - zp) / sub / std::sqrt(0.3 * (1 + zm) * (1 + zm) * (1 + zm) + 0.7);
      }
      zp = z; tz[i] = z; td[i] = 4282.7494 * dc * (1.0 + z);
    }
    std::vector<float> key((size_t)Ninj);
    for (int64_t i = 0; i < Ninj; ++i) {
      const double d = dL[i];
      double z;
      if (!(d > td[0])) z = (d > 0.0) ? d / 4282.7494 : 0.0;
      else if (d >= td[NT - 1]) z = tz[NT - 1];
      else {
        const int k = (int)(std::upper_bound(td.begin(), td.end(), d) - td.begin());
        z = tz[k - 1] + (tz[k] - tz[k - 1]) * (d - td[k - 1]) / (td[k] - td[k - 1]);
      }
      key[i] = (float)(m2det[i] / (1.0 + z));
      if (!(key[i] == key[i])) key[i] = INFINITY;          // NaN input: last, and the comparator stays a strict weak order
      perm[i] = (int)i;
    }
    std::sort(perm.begin(), perm.end(), [&key](int x, int y) { return key[x] < key[y] || (key[x] == key[y] && x < y); });
  }
  std::vector<double> t1((size_t)Ninj), t2((size_t)Ninj), t3((size_t)Ninj), t4((size_t)Ninj);
  for (int64_t i = 0; i < Ninj; ++i) { const int q = perm[i]; t1[i] = m1det[q]; t2[i] = m2det[q]; t3[i] = dL[q]; t4[i] = p_draw[q]; }
  m1det = t1.data(); m2det = t2.data(); dL = t3.data(); p_draw = t4.data();
  CU(h->inj_m1.upload(m1det, Ninj), "upload inj m1det");
  CU(h->inj_m2.upload(m2det, Ninj), "upload inj m2det");
  CU(h->inj_dL.upload(dL, Ninj), "upload inj dL");
  CU(h->inj_pd.upload(p_draw, Ninj), "upload inj p_draw");
  if (h->cfg.fp_mode == CHB_FP32) {
    std::vector<float4> s4(Ninj);
    std::vector<float2> l2(Ninj);
    for (int64_t i = 0; i < Ninj; ++i) {
      s4[i] = make_float4((float)dL[i], (float)m1det[i], (float)m2det[i], (float)(1.0 / p_draw[i]));
      l2[i] = make_float2((float)std::log2(m1det[i]), (float)std::log2(m2det[i]));
    }
    CU(h->inj_s4.upload(s4.data(), Ninj), "upload packed injections");
    CU(h->inj_l2.upload(l2.data(), Ninj), "upload injection log2 masses");
  }
  h->Ninj = Ninj; h->have_inj = true;
  return CHB_OK;
}

// One-time device layout of the event samples.  For the marginalised KDE the samples of every
// event are bucketed by pixel slot (slot i <-> pixels_opt_nsides[ev][i]; unmatched -> slot P),
// so that `pe_pix == pixels[i]` (likelihood.py:176) becomes a contiguous range.  All per-event
// reductions are order-independent, so the other variants use the same permuted arrays.
static int prepare(chb_handle* h) {
  if (!h->dirty) return CHB_OK;
  const int64_t Nev = h->Nev, Ns = h->Ns, P = h->P;
  const size_t n = (size_t)Nev * Ns;
  const int kind = h->cfg.kind_p_gw;
  // 1-D kinds (both arithmetic modes): every per-event reduction is order-independent, so the samples of every event
  // are sorted by dL.  z_from_dGW is monotone in dL, hence the reweighted z's come out sorted for every hyper-point and
  // the windowed KDEs (kde_win.cuh, kde_win64.cuh) only visit the grid points near each chunk of samples.
  // ('full': the 3-D KDE skips blocks of dL-sorted samples that are far from a tile of evaluation points in z)
  h->sorted = (kind == CHB_PGW_1D || kind == CHB_PGW_APPROX || kind == CHB_PGW_FULL);
  const bool bucket = !h->sorted && h->have_pixels && !h->h_pe_pix.empty();
  const bool sky = !h->h_ra.empty();
  std::vector<double> t1, t2, t3, t4, t5, t6;
  const double *p1 = h->h_m1.data(), *p2 = h->h_m2.data(), *p3 = h->h_dL.data(), *p4 = h->h_prior.data();
  const double *p5 = sky ? h->h_ra.data() : nullptr, *p6 = sky ? h->h_dec.data() : nullptr;
  if (bucket || h->sorted) {
    t1.resize(n); t2.resize(n); t3.resize(n); t4.resize(n);
    if (sky) { t5.resize(n); t6.resize(n); }
  }
  if (bucket) {
    std::vector<int> off((size_t)Nev * (P + 2));
    std::vector<int> slot(Ns), cnt(P + 2);
    for (int64_t e = 0; e < Nev; ++e) {
      std::unordered_map<int64_t, int> map;
      for (int64_t p = 0; p < P; ++p) {
        int64_t pid = h->h_pixels[e * P + p];
        if (pid != -100 && !map.count(pid)) map[pid] = (int)p;
      }
      std::fill(cnt.begin(), cnt.end(), 0);
      for (int64_t j = 0; j < Ns; ++j) {
        auto it = map.find(h->h_pe_pix[e * Ns + j]);
        slot[j] = (it == map.end()) ? (int)P : it->second;
        cnt[slot[j] + 1]++;
      }
      for (int64_t p = 0; p <= P; ++p) cnt[p + 1] += cnt[p];
      for (int64_t p = 0; p <= P + 1; ++p) off[e * (P + 2) + p] = cnt[p];
      std::vector<int> cur(cnt.begin(), cnt.end() - 1);
      for (int64_t j = 0; j < Ns; ++j) {
        size_t d = (size_t)e * Ns + cur[slot[j]]++, s = (size_t)e * Ns + j;
        t1[d] = p1[s]; t2[d] = p2[s]; t3[d] = p3[s]; t4[d] = p4[s];
        if (sky) { t5[d] = p5[s]; t6[d] = p6[s]; }
      }
    }
    CU(h->pix_off.upload(off.data(), off.size()), "upload pixel offsets");
  } else if (h->sorted) {
    std::vector<int> perm(Ns);
    for (int64_t e = 0; e < Nev; ++e) {
      const size_t o = (size_t)e * Ns;
      for (int64_t j = 0; j < Ns; ++j) perm[j] = (int)j;
      const double* key = p3 + o;
      // (NaN distances sort last: the comparator must stay a strict weak order whatever the input holds)
      std::sort(perm.begin(), perm.end(), [key](int x, int y) {
        const double a = (key[x] == key[x]) ? key[x] : INFINITY, b = (key[y] == key[y]) ? key[y] : INFINITY;
        return a < b || (a == b && x < y);
      });
      for (int64_t j = 0; j < Ns; ++j) {
        const size_t d = o + j, s = o + perm[j];
        t1[d] = p1[s]; t2[d] = p2[s]; t3[d] = p3[s]; t4[d] = p4[s];
        if (sky) { t5[d] = p5[s]; t6[d] = p6[s]; }
      }
    }
  }
  if (bucket || h->sorted) { p1 = t1.data(); p2 = t2.data(); p3 = t3.data(); p4 = t4.data(); if (sky) { p5 = t5.data(); p6 = t6.data(); } }
  if (sky) { CU(h->ra.upload(p5, n), "upload ra"); CU(h->dec.upload(p6, n), "upload dec"); }
  if (h->cfg.fp_mode == CHB_FP32) {
    // single-precision packed copies for the fp32 reweighting path: {dL, m1det, m2det, 1/pe_prior}, {log2 m1det, log2 m2det}
    // (the fp64 sample arrays are not read in this mode and are not uploaded)
    std::vector<float4> s4(n);
    std::vector<float2> l2(n);
    for (size_t i = 0; i < n; ++i) {
      s4[i] = make_float4((float)p3[i], (float)p1[i], (float)p2[i], (float)(1.0 / p4[i]));
      l2[i] = make_float2((float)std::log2(p1[i]), (float)std::log2(p2[i]));
    }
    CU(h->s4.upload(s4.data(), n), "upload packed samples");
    CU(h->l2.upload(l2.data(), n), "upload log2 masses");
    h->m1d.release(); h->m2d.release(); h->dL.release(); h->prior.release();
  } else {
    CU(h->m1d.upload(p1, n), "upload m1det"); CU(h->m2d.upload(p2, n), "upload m2det");
    CU(h->dL.upload(p3, n), "upload dL"); CU(h->prior.upload(p4, n), "upload pe_prior");
  }
  h->dirty = false;
  h->cat_collapsed = false;
  return CHB_OK;
}

static int eval_impl(chb_handle* h, int64_t n_hyper, const double* d_hyper, double* d_log_like, double* d_partials,
                     double* d_pgw, cudaStream_t s) {
  const chb_config& c = h->cfg;
  const bool do_num = h->have_events, do_sel = h->have_inj;
  if (do_num) {
    if (c.kind_p_gw != CHB_PGW_1D && !h->have_pixels) return fail(h, CHB_ERR_STATE, "pixelated kind needs chb_set_pixels");
    if (c.catalog_kind == 1 && c.kind_p_gw != CHB_PGW_1D && !h->have_catalog)
      return fail(h, CHB_ERR_STATE, "catalog_kind=1 needs chb_set_catalog");
    int rc = prepare(h);
    if (rc != CHB_OK) return rc;
  }
  const TableLayout lay = h->mc.lay;
  CU(h->tabs.alloc((size_t)n_hyper * lay.total()), "alloc tables");
  CU(h->HC.alloc((size_t)n_hyper * CHB_NHC), "alloc constants");
  cudaEventRecord(h->ev[0], s);
  CU(launch_build_tables(h->mc, (int)n_hyper, d_hyper, h->tabs.p, h->HC.p, s), "build_tables launch");
  h->launches++;
  cudaEventRecord(h->ev[1], s);
  cudaEventRecord(h->evz, s);

  double* ll = nullptr;
  if (do_num) {
    ll = d_log_like;
    if (!ll) { CU(h->log_like.alloc((size_t)n_hyper * h->Nev), "alloc log_like"); ll = h->log_like.p; }
    NumArgs a;
    memset(&a, 0, sizeof(a));
    a.mc = h->mc;
    a.kind = c.kind_p_gw; a.kernel = c.kernel; a.bw_method = c.bw_method; a.use_cut = c.use_cut_grid;
    a.binning = (c.kind_p_gw == CHB_PGW_FULL) ? 0 : c.binning; a.num_bins = c.num_bins; a.fp_mode = c.fp_mode;
    a.bw_value = c.bw_value; a.cut_grid = c.cut_grid; a.pe_neff = c.pe_neff;
    a.rec_off = h->opt_kde_direct; a.bin_runs = h->opt_bin_runs; a.epan_blocks = h->opt_epan_blocks;
    a.kde_win_iters = h->sorted ? h->opt_kde_win : 0;
    a.win_t2 = (float)h->opt_kde_win_t2;
    a.Nev = (int)h->Nev; a.Ns = (int)h->Ns; a.Nz = (int)h->Nz; a.P = (int)std::max<int64_t>(h->P, 1);
    a.m1d = h->m1d.p; a.m2d = h->m2d.p; a.dL = h->dL.p; a.prior = h->prior.p; a.ra = h->ra.p; a.dec = h->dec.p;
    a.zgrids = h->zgrids.p; a.pix_off = h->pix_off.p; a.ra_pix = h->ra_pix.p; a.dec_pix = h->dec_pix.p;
    a.gw_pdf = h->gw_pdf.p; a.neff_pix = h->neff_pix.p;
    a.p_cat = h->have_catalog ? h->p_cat.p : nullptr; a.P_compl = h->have_catalog ? h->P_compl.p : nullptr;
    a.s4 = h->s4.p; a.l2 = h->l2.p;
    a.catA = nullptr; a.catB = nullptr;
    if (c.kind_p_gw == CHB_PGW_APPROX && h->have_catalog) {
      if (!h->cat_collapsed) {
        CU(h->catA.alloc((size_t)h->Nev * h->Nz), "alloc catA");
        CU(h->catB.alloc((size_t)h->Nev * h->Nz), "alloc catB");
        CU(launch_catalog_collapse((int)h->Nev, (int)h->P, (int)h->Nz, h->p_cat.p, h->gw_pdf.p, h->neff_pix.p,
                                   h->catA.p, h->catB.p, s), "catalog_collapse launch");
        h->launches++;
        h->cat_collapsed = true;
      }
      a.catA = h->catA.p; a.catB = h->catB.p;
    }
    if (!h->have_catalog) a.mc.catalog_kind = 0;
    a.n_hyper = (int)n_hyper; a.hyper = d_hyper; a.tabs = h->tabs.p; a.HC = h->HC.p;
    CU(h->like_raw.alloc((size_t)n_hyper * h->Nev), "alloc like_raw");
    a.log_like = ll; a.like_raw = h->like_raw.p; a.p_gw_out = d_pgw;
    h->last_n_hyper = n_hyper;
    if (c.kind_p_gw == CHB_PGW_MARG && !h->pix_off.p) return fail(h, CHB_ERR_STATE, "marginalized kind needs pixels_pe_opt_nside");
    const long long units = (long long)h->Nev * n_hyper;
    a.prof = nullptr;
    bool fast = false;
    if (c.fp_mode == CHB_FP32) {
      // The dynamic-shared-memory attribute of a kernel is process-wide: it is only ever set to the device maximum
      // (never to a handle's own footprint), so handles with different shapes cannot invalidate each other's launches.
      const int optin = h->max_smem_optin;
      const size_t fit = (size_t)optin - 1024;     // dynamic footprints must leave room for the kernels' static shared memory
      // (1) 1-D kinds: ONE fused kernel per step (numerator_fused.cu), samples never leave shared memory.
      // Events with few samples run the 128-thread instantiation (six CTAs per SM): the per-unit work every thread
      // repeats and the barriers weigh as much as the sums there (option fused_nt: 0 = by sample count, 128, 256).
      // (measured on C5, 1000 samples per event: 256 threads 1.92 s, 128 threads 1.51 s, 64 threads 1.43 s per step of
      //  4.1e7 units; on C3, 5000 samples: 256 threads 21.7 ms, 128 threads 22.2 ms)
      const bool nt64 = h->opt_fused_nt == 64 || (h->opt_fused_nt == 0 && h->Ns <= 1024);
      const bool nt128 = !nt64 && (h->opt_fused_nt == 128 || (h->opt_fused_nt == 0 && h->Ns <= 2048));
      const size_t ff = nt64 ? numerator_fused_smem_bytes_nt64(a) : nt128 ? numerator_fused_smem_bytes_nt128(a) : numerator_fused_smem_bytes(a);
      bool fused = h->opt_fused && numerator_fused_supported(a) && ff <= fit;
      if (fused && h->fused_per < 0) {
        if (nt64) h->fused_per = (numerator_fused_configure_nt64(optin) == cudaSuccess) ? numerator_fused_ctas_per_sm_nt64(ff) : 0;
        else if (nt128) h->fused_per = (numerator_fused_configure_nt128(optin) == cudaSuccess) ? numerator_fused_ctas_per_sm_nt128(ff) : 0;
        else h->fused_per = (numerator_fused_configure(optin) == cudaSuccess) ? numerator_fused_ctas_per_sm(ff) : 0;
        cudaGetLastError();
      }
      if (fused && h->fused_per < 1) fused = false;
      // (1b) 'marginalized' + binning (the reference's default options): fused kernel, warp per pixel
      const size_t fm = numerator_marg_smem_bytes(a);
      bool marg = !fused && h->opt_fused && numerator_marg_supported(a) && fm <= fit;
      if (marg && h->marg_per < 0) {
        h->marg_per = (numerator_marg_configure(optin) == cudaSuccess) ? numerator_marg_ctas_per_sm(fm) : 0;
        cudaGetLastError();
      }
      if (marg && h->marg_per < 1) marg = false;
      // (2) other kinds (and the A/B switch): split form -- reweighting kernel -> stage buffers in global memory ->
      // KDE/z-integral kernel; fused MODE 0 kernel for odd Ns or when there is no memory for the stage.
      size_t fs = numerator_f32_smem_bytes(a, 0);
      const size_t fs1 = numerator_f32_smem_bytes(a, 1), fs2 = numerator_f32_smem_bytes(a, 2);
      bool split = !fused && !marg && h->opt_split && (h->Ns % 2 == 0) && fs2 <= fit && fs1 <= fit;
      int64_t nb = n_hyper;                        // hyper-points per batch of the split form
      if (split) {
        // the plan (batch size, stage buffers, occupancy) is cached per n_hyper and reset by chb_set_* / chb_set_option:
        // no driver queries inside the timed region of later evaluations
        if (h->plan_nh != n_hyper) {
          const size_t budget = (size_t)(h->opt_stage_gb * (double)((size_t)1 << 30));
          const size_t per_h = (size_t)h->Nev * h->Ns * sizeof(float2);
          nb = std::max<int64_t>(1, std::min<int64_t>(n_hyper, (int64_t)(budget / per_h)));
          size_t free_b = 0, total_b = 0;
          cudaMemGetInfo(&free_b, &total_b);
          while (nb > 1 && h->zw_stage.n < (size_t)nb * h->Nev * h->Ns && (size_t)nb * per_h > free_b / 2) nb = (nb + 1) / 2;
          h->plan_ok = false;
          if (h->zw_stage.alloc((size_t)nb * h->Nev * h->Ns) == cudaSuccess &&
              h->unit_stats.alloc((size_t)nb * h->Nev * 8) == cudaSuccess &&
              numerator_f32_configure(a.kind, 1, optin) == cudaSuccess && numerator_f32_configure(a.kind, 2, optin) == cudaSuccess) {
            h->plan_per1 = numerator_f32_ctas_per_sm(a.kind, 1, fs1);
            h->plan_per2 = numerator_f32_ctas_per_sm(a.kind, 2, fs2);
            h->plan_ok = h->plan_per1 >= 1 && h->plan_per2 >= 1;
          }
          cudaGetLastError();
          h->plan_nb = nb; h->plan_nh = n_hyper;
        }
        nb = h->plan_nb;
        split = h->plan_ok;
      }
      if (fused || marg || split || fs <= fit) {
        // z-grid terms for all (hyper-point, event, k) in one full-occupancy pass when they fit (<= 16 GiB and a
        // quarter of the free memory; otherwise the numerator kernels evaluate them in place)
        const size_t zt_elems = (size_t)n_hyper * h->Nev * h->Nz;
        a.zterms = nullptr; a.zterms_out = nullptr; a.zterms_h0 = 0;
        bool zt_fit = zt_elems * sizeof(float2) <= std::min<size_t>((size_t)2 << 30, (size_t)(h->opt_zterms_gb * (double)((size_t)1 << 30)));
        size_t zt_budget = (size_t)(h->opt_zterms_gb * (double)((size_t)1 << 30));
        if (!zt_fit) {
          size_t free_b = 0, total_b = 0;
          cudaMemGetInfo(&free_b, &total_b);
          zt_budget = std::min(zt_budget, std::max(free_b / 4, h->zterms.n * sizeof(float2)));
          zt_fit = zt_elems * sizeof(float2) <= zt_budget;
        }
        // the one-launch kernels (fused / 'marginalized') take the hyper-points in batches whose z-grid terms fit the
        // budget (large walker batches: C5 on one GPU is 98 GB of terms) -- always the precomputed terms, never the
        // in-kernel evaluation
        const size_t zt_per_h = (size_t)h->Nev * h->Nz;
        int64_t zb = n_hyper;
        if (!zt_fit && (fused || marg)) {
          zb = std::max<int64_t>(1, (int64_t)(zt_budget / (zt_per_h * sizeof(float2))));
          zt_fit = zt_per_h * sizeof(float2) <= zt_budget;
        }
        if (zt_fit) {
          CU(h->zterms.alloc((size_t)std::min<int64_t>(zb, n_hyper) * zt_per_h), "alloc z-grid terms");
          a.zterms_out = h->zterms.p;
          a.zterms = h->zterms.p;
        }
        if (zt_fit && !(fused || marg)) {
          CU(launch_zgrid_terms(a, 0, (int)n_hyper, s), "zgrid_terms launch");
          h->launches++;
          cudaEventRecord(h->evz, s);
        }
        if (fused || marg) {
          const int per = fused ? h->fused_per : h->marg_per;
          const size_t smem = fused ? ff : fm;
          h->num_smem = smem;
          bool first = true;
          for (int64_t h0 = 0; h0 < n_hyper; h0 += zb) {
            const int64_t nh = std::min<int64_t>(zb, n_hyper - h0);
            if (zt_fit) {
              CU(launch_zgrid_terms(a, (int)h0, (int)nh, s), "zgrid_terms launch");
              h->launches++;
              if (first) cudaEventRecord(h->evz, s);      // (timings: the first batch's terms; later batches count as numerator)
            }
            NumArgs b = a;                         // this batch: pointers shifted to hyper-point h0
            b.n_hyper = (int)nh;
            b.hyper = a.hyper + (size_t)h0 * CHB_NPAR; b.tabs = a.tabs + (size_t)h0 * a.mc.lay.total(); b.HC = a.HC + (size_t)h0 * CHB_NHC;
            b.log_like = a.log_like + (size_t)h0 * h->Nev; b.like_raw = a.like_raw + (size_t)h0 * h->Nev;
            if (a.p_gw_out) b.p_gw_out = a.p_gw_out + (size_t)h0 * h->Nev * (size_t)(c.kind_p_gw == CHB_PGW_1D ? 1 : a.P) * h->Nz;
            const long long ub = (long long)h->Nev * nh;
            const int grid = (int)std::min<long long>(ub, (long long)h->sm_count * per);
            h->num_grid = grid;
            if (fused && nt64) { CU(launch_numerator_fused_nt64(b, grid, smem, s), "numerator_fused launch"); }
            else if (fused && nt128) { CU(launch_numerator_fused_nt128(b, grid, smem, s), "numerator_fused launch"); }
            else if (fused) { CU(launch_numerator_fused(b, grid, smem, s), "numerator_fused launch"); }
            else { CU(launch_numerator_marg(b, grid, smem, s), "numerator_marg launch"); }
            if (!first) h->launches++;             // the common `launches++` below counts the first one
            first = false;
          }
          fast = true;
        } else if (split) {
          const int per1 = h->plan_per1, per2 = h->plan_per2;
          h->num_smem = fs2;
          for (int64_t h0 = 0; h0 < n_hyper; h0 += nb) {
            const int64_t nh = std::min<int64_t>(nb, n_hyper - h0);
            NumArgs b = a;                         // this batch: pointers shifted to hyper-point h0
            b.n_hyper = (int)nh;
            b.hyper = a.hyper + (size_t)h0 * CHB_NPAR; b.tabs = a.tabs + (size_t)h0 * a.mc.lay.total(); b.HC = a.HC + (size_t)h0 * CHB_NHC;
            b.log_like = a.log_like + (size_t)h0 * h->Nev; b.like_raw = a.like_raw + (size_t)h0 * h->Nev;
            if (a.p_gw_out) b.p_gw_out = a.p_gw_out + (size_t)h0 * h->Nev * (size_t)(c.kind_p_gw == CHB_PGW_1D ? 1 : a.P) * h->Nz;
            if (a.zterms) b.zterms = a.zterms + (size_t)h0 * h->Nev * h->Nz;
            b.zw_stage = h->zw_stage.p; b.unit_stats = h->unit_stats.p;
            const long long ub = (long long)h->Nev * nh;
            const int g1 = (int)std::min<long long>(ub, (long long)h->sm_count * per1);
            const int g2 = (int)std::min<long long>(ub, (long long)h->sm_count * per2);
            h->num_grid = g2;
            b.prof = nullptr;
            CU(launch_numerator_f32(b, 1, g1, fs1, s), "reweight_f32 launch");
            if (h->want_prof) { CU(h->prof.alloc((size_t)g2 * 8), "alloc profile"); b.prof = h->prof.p; }
            CU(launch_numerator_f32(b, 2, g2, fs2, s), "kde_f32 launch");
            h->launches += 2;
          }
          h->launches--;                           // the common `launches++` below counts one of them
          fast = true;
        } else {
          CU(numerator_f32_configure(a.kind, 0, optin), "numerator_f32 smem opt-in");
          int per_sm = numerator_f32_ctas_per_sm(a.kind, 0, fs);
          if (per_sm >= 1) {
            int grid = (int)std::min<long long>(units, (long long)h->sm_count * per_sm);
            h->num_grid = grid; h->num_smem = fs;
            if (h->want_prof) { CU(h->prof.alloc((size_t)grid * 8), "alloc profile"); a.prof = h->prof.p; }
            CU(launch_numerator_f32(a, 0, grid, fs, s), "numerator_f32 launch");
            fast = true;
          }
        }
      }
    }
    if (!fast) {
      // generic kernel: sample staging in shared memory when it fits, else per-CTA global scratch (L2-resident)
      size_t smem = numerator_smem_bytes(a, true);
      h->stage_in_smem = smem <= (size_t)h->max_smem_optin;
      if (!h->stage_in_smem) smem = numerator_smem_bytes(a, false);
      if (smem > (size_t)h->max_smem_optin) return fail(h, CHB_ERR_UNSUPPORTED, "z grid / tables do not fit in shared memory");
      CU(numerator_configure(smem), "numerator smem opt-in");
      int grid = (int)std::min<long long>(units, (long long)h->sm_count);
      if (!h->stage_in_smem) {
        a.scratch_stride = numerator_scratch_doubles(a);
        CU(h->scratch.alloc((size_t)grid * a.scratch_stride), "alloc staging scratch");
        a.scratch = h->scratch.p;
      }
      h->num_grid = grid; h->num_smem = smem;
      if (h->want_prof) { CU(h->prof.alloc((size_t)grid * 8), "alloc profile"); a.prof = h->prof.p; }
      CU(launch_numerator(a, grid, numerator_block_threads(), smem, s), "numerator launch");
    }
    h->launches++;
  }
  cudaEventRecord(h->ev[2], s);

  int tiles = 0;
  if (do_sel) {
    // enough CTAs to fill the machine even for a single hyper-point; >= 1024 injections per tile
    long long want = (4LL * h->sm_count + n_hyper - 1) / n_hyper;
    long long maxt = std::max<long long>(1, h->Ninj / 1024);
    tiles = (int)std::max<long long>(1, std::min(want, maxt));
    // ... and a CTA count that fills whole waves: among want .. 6 want tiles the one whose last wave is fullest
    // (C3: 256 hyper-points x 3 tiles on 444 resident CTAs was 1.73 waves, i.e. a 14 % tail)
    const int per_sm = selection_ctas_per_sm(h->mc, c.fp_mode);
    if (per_sm > 0) {
      const long long slots = (long long)per_sm * h->sm_count;
      double best = 0.0;
      for (long long t = tiles; t <= std::min(maxt, 6 * std::max<long long>(want, 1)); ++t) {
        const long long ctas = t * n_hyper, waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best + 0.02) { best = eff; tiles = (int)t; }
      }
    }
    CU(h->tile_part.alloc((size_t)n_hyper * tiles * 2), "alloc selection partials");
    SelArgs sa;
    sa.mc = h->mc; sa.Ninj = (int)h->Ninj; sa.n_hyper = (int)n_hyper; sa.tiles = tiles;
    sa.m1d = h->inj_m1.p; sa.m2d = h->inj_m2.p; sa.dL = h->inj_dL.p; sa.p_draw = h->inj_pd.p;
    sa.fp_mode = c.fp_mode; sa.s4 = h->inj_s4.p; sa.l2 = h->inj_l2.p;
    sa.hyper = d_hyper; sa.tabs = h->tabs.p; sa.HC = h->HC.p; sa.tile_part = h->tile_part.p;
    CU(launch_selection(sa, s), "selection launch");
    h->launches++;
  }
  cudaEventRecord(h->ev[3], s);
  CU(launch_reduce((int)n_hyper, (int)h->Nev, tiles, do_num ? ll : nullptr, do_sel ? h->tile_part.p : nullptr,
                   d_partials, s), "reduce launch");
  h->launches++;
  cudaEventRecord(h->ev[4], s);
  return CHB_OK;
}

int chb_set_option(chb_handle* h, const char* name, double value) {
  if (!h || !name) return CHB_ERR_INVALID;
  const std::string n(name);
  if (n == "fused") h->opt_fused = value != 0.0;
  else if (n == "split") h->opt_split = value != 0.0;
  else if (n == "kde_win") { if (value < 0 || value > 4096) return fail(h, CHB_ERR_INVALID, "kde_win out of range"); h->opt_kde_win = (int)value; }
  else if (n == "kde_win_t2") { if (!(value >= 16.0 && value <= 60.0)) return fail(h, CHB_ERR_INVALID, "kde_win_t2 must be in [16, 60]"); h->opt_kde_win_t2 = value; }
  else if (n == "kde_direct") h->opt_kde_direct = value != 0.0;
  else if (n == "bin_runs") h->opt_bin_runs = value != 0.0;
  else if (n == "epan_blocks") h->opt_epan_blocks = value != 0.0;
  else if (n == "fused_nt") { if (value != 0.0 && value != 64.0 && value != 128.0 && value != 256.0) return fail(h, CHB_ERR_INVALID, "fused_nt must be 0, 64, 128 or 256"); h->opt_fused_nt = (int)value; }
  else if (n == "zterms_gb") { if (!(value > 0.0)) return fail(h, CHB_ERR_INVALID, "zterms_gb must be positive"); h->opt_zterms_gb = value; }
  else if (n == "stage_gb") { if (!(value > 0.0)) return fail(h, CHB_ERR_INVALID, "stage_gb must be positive"); h->opt_stage_gb = value; }
  else return fail(h, CHB_ERR_INVALID, "unknown option '" + n + "'");
  h->plan_nh = -1; h->fused_per = -1; h->marg_per = -1;
  return CHB_OK;
}

int chb_eval_device(chb_handle* h, int64_t n_hyper, const double* d_hyper, double* d_log_like, double* d_partials,
                    double* d_p_gw, void* cuda_stream) {
  if (!h) return CHB_ERR_INVALID;
  if (n_hyper < 1 || !d_hyper || !d_partials) return fail(h, CHB_ERR_INVALID, "bad eval arguments");
  if (!h->have_events && !h->have_inj) return fail(h, CHB_ERR_STATE, "nothing to evaluate: set events and/or injections");
  DevGuard _dg(h->cfg.device);
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  return eval_impl(h, n_hyper, d_hyper, d_log_like, d_partials, d_p_gw, s);
}

int chb_eval(chb_handle* h, int64_t n_hyper, const double* hyper, double* log_like_evs, double* partials, double* p_gw) {
  if (!h) return CHB_ERR_INVALID;
  if (n_hyper < 1 || !hyper || !partials) return fail(h, CHB_ERR_INVALID, "bad eval arguments");
  if (!h->have_events && !h->have_inj) return fail(h, CHB_ERR_STATE, "nothing to evaluate: set events and/or injections");
  DevGuard _dg(h->cfg.device);
  cudaStream_t s = h->stream;
  CU(h->hyper.alloc((size_t)n_hyper * CHB_NPAR), "alloc hyper");
  CU(h->partials.alloc((size_t)n_hyper * 3), "alloc partials");
  CU(cudaMemcpyAsync(h->hyper.p, hyper, (size_t)n_hyper * CHB_NPAR * sizeof(double), cudaMemcpyHostToDevice, s), "H2D hyper");
  double* d_ll = nullptr;
  double* d_pgw = nullptr;
  size_t pgw_n = 0;
  if (h->have_events) {
    CU(h->log_like.alloc((size_t)n_hyper * h->Nev), "alloc log_like");
    d_ll = h->log_like.p;
    if (p_gw) {
      pgw_n = (size_t)n_hyper * h->Nev * (h->cfg.kind_p_gw == CHB_PGW_1D ? 1 : h->P) * h->Nz;
      CU(h->pgw.alloc(pgw_n), "alloc p_gw");
      d_pgw = h->pgw.p;
    }
  }
  int rc = eval_impl(h, n_hyper, h->hyper.p, d_ll, h->partials.p, d_pgw, s);
  if (rc != CHB_OK) return rc;
  if (log_like_evs && d_ll)
    CU(cudaMemcpyAsync(log_like_evs, d_ll, (size_t)n_hyper * h->Nev * sizeof(double), cudaMemcpyDeviceToHost, s), "D2H log_like");
  CU(cudaMemcpyAsync(partials, h->partials.p, (size_t)n_hyper * 3 * sizeof(double), cudaMemcpyDeviceToHost, s), "D2H partials");
  if (d_pgw) CU(cudaMemcpyAsync(p_gw, d_pgw, pgw_n * sizeof(double), cudaMemcpyDeviceToHost, s), "D2H p_gw");
  CU(cudaStreamSynchronize(s), "eval synchronize");
  return CHB_OK;
}

int chb_last_numlike_evs(chb_handle* h, double* like_evs) {
  if (!h || !like_evs) return CHB_ERR_INVALID;
  if (!h->have_events || h->last_n_hyper < 1 || !h->like_raw.p) return fail(h, CHB_ERR_STATE, "no numerator evaluated yet");
  DevGuard _dg(h->cfg.device);
  CU(cudaStreamSynchronize(h->stream), "synchronize");
  CU(cudaMemcpy(like_evs, h->like_raw.p, (size_t)h->last_n_hyper * h->Nev * sizeof(double), cudaMemcpyDeviceToHost), "D2H like");
  return CHB_OK;
}

int chb_finalize(const chb_config* cfg, int64_t n_hyper, int64_t Nev_total, const double* hyper, const double* partials,
                 double* log_like_num, double* log_Nexp, double* log_hyper, double* neff_inj, double* N_exp) {
  if (!cfg || n_hyper < 1 || !hyper || !partials) return CHB_ERR_INVALID;
  const double Ninj = cfg->N_inj;
  for (int64_t i = 0; i < n_hyper; ++i) {
    const double R0 = hyper[i * CHB_NPAR + CHB_P_R0];
    double lnum = partials[i * 3 + 0];
    const double s1 = partials[i * 3 + 1], s2 = partials[i * 3 + 2];
    const double xi = s1 / Ninj;                                      // selection_function.py:39
    double Nexp = cfg->Tobs * xi;                                     // :41
    double neff = std::numeric_limits<double>::quiet_NaN();
    if (cfg->check_neff) {
      const double var = s2 / (Ninj * Ninj) - xi * xi / Ninj;         // :44
      neff = xi * xi / var;                                           // :45
      if (neff < cfg->N_eff) Nexp = 0.0;                              // :46-47
    }
    double lh;
    if (!cfg->scale_free) {                                           // likelihood.py:299-300,313-316
      lnum += (double)Nev_total * std::log(R0 * cfg->Tobs);
      lh = lnum - Nexp;
    } else {
      lh = lnum - (double)Nev_total * std::log(Nexp);
    }
    if (log_like_num) log_like_num[i] = lnum;
    if (log_Nexp) log_Nexp[i] = std::log(Nexp);
    if (log_hyper) log_hyper[i] = lh;
    if (neff_inj) neff_inj[i] = neff;
    if (N_exp) N_exp[i] = Nexp;
  }
  return CHB_OK;
}

static int model_tables_device(const chb_config* cfg, const double* params, ModelCfg& mc, DevBuf<double>& dP,
                               DevBuf<double>& dT, DevBuf<double>& dHC) {
  chb_handle* h = nullptr;
  std::string why;
  int rc = validate_cfg(cfg, why);
  if (rc != CHB_OK) return fail(nullptr, rc, why);
  if (!params) return fail(nullptr, CHB_ERR_INVALID, "params is NULL");
  if (chb_device_count() == 0) return fail(nullptr, CHB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  CU(cudaSetDevice(cfg->device), "cudaSetDevice");      // (the callers hold a DevGuard)
  mc = model_cfg(*cfg);
  CU(dP.upload(params, CHB_NPAR), "upload params");
  CU(dT.alloc(mc.lay.total()), "alloc tables");
  CU(dHC.alloc(CHB_NHC), "alloc constants");
  CU(launch_build_tables(mc, 1, dP.p, dT.p, dHC.p, 0), "build_tables launch");
  return CHB_OK;
}

int chb_model_eval(const chb_config* cfg, int which, const double* params, int64_t n, const double* a, const double* b,
                   const double* c, double* out) {
  chb_handle* h = nullptr;
  if (n < 0 || (n > 0 && (!a || !out))) return fail(nullptr, CHB_ERR_INVALID, "bad array arguments");
  if ((which == CHB_F_P_M1M2 && !b) || (which == CHB_F_POP_RATE_DET_INJ && (!b || !c)))
    return fail(nullptr, CHB_ERR_INVALID, "missing array argument");
  ModelCfg mc;
  DevGuard _dg;
  DevBuf<double> dP, dT, dHC, da, db, dc, dout;
  int rc = model_tables_device(cfg, params, mc, dP, dT, dHC);
  if (rc == CHB_OK && n > 0) {
    do {
      cudaError_t e;
      if ((e = da.upload(a, n)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "upload a"); break; }
      if (b && (e = db.upload(b, n)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "upload b"); break; }
      if (c && (e = dc.upload(c, n)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "upload c"); break; }
      if ((e = dout.alloc(n)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "alloc out"); break; }
      if ((e = launch_model_eval(mc, which, dP.p, dT.p, dHC.p, n, da.p, b ? db.p : nullptr, c ? dc.p : nullptr, dout.p, 0)) != cudaSuccess) {
        rc = cuda_fail(nullptr, e, "model_eval launch"); break; }
      if ((e = cudaMemcpy(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) {
        rc = cuda_fail(nullptr, e, "D2H out"); break; }
    } while (0);
  }
  dP.release(); dT.release(); dHC.release(); da.release(); db.release(); dc.release(); dout.release();
  (void)h;
  return rc;
}

int chb_model_tables(const chb_config* cfg, const double* params, double* z_grid_interp, double* integral_invE_interp,
                     double* m_grid, double* cdf_m2_conditioned, double* norm_p_m1) {
  ModelCfg mc;
  DevGuard _dg;
  DevBuf<double> dP, dT, dHC;
  int rc = model_tables_device(cfg, params, mc, dP, dT, dHC);
  if (rc == CHB_OK) {
    std::vector<double> T(mc.lay.total()), HC(CHB_NHC);
    cudaError_t e = cudaMemcpy(T.data(), dT.p, T.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(HC.data(), dHC.p, HC.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(nullptr, e, "D2H tables");
    else {
      if (z_grid_interp) memcpy(z_grid_interp, T.data() + mc.lay.off_zg(), mc.lay.rc * sizeof(double));
      if (integral_invE_interp) memcpy(integral_invE_interp, T.data() + mc.lay.off_iinv(), mc.lay.rc * sizeof(double));
      if (m_grid) memcpy(m_grid, T.data() + mc.lay.off_mg(), mc.lay.rm * sizeof(double));
      if (cdf_m2_conditioned) memcpy(cdf_m2_conditioned, T.data() + mc.lay.off_cdf(), mc.lay.rm * sizeof(double));
      if (norm_p_m1) *norm_p_m1 = HC[HC_NORM_P_M1];
    }
  }
  dP.release(); dT.release(); dHC.release();
  return rc;
}

int chb_phase_profile(chb_handle* h, int enable, double out[8]) {
  if (!h) return CHB_ERR_INVALID;
  DevGuard _dg(h->cfg.device);
  if (out) {
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    if (h->want_prof && h->prof.p && h->num_grid > 0) {
      std::vector<unsigned long long> v((size_t)h->num_grid * 8);
      CU(cudaStreamSynchronize(h->stream), "synchronize");
      CU(cudaDeviceSynchronize(), "synchronize");
      CU(cudaMemcpy(v.data(), h->prof.p, v.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "D2H profile");
      for (int b = 0; b < h->num_grid; ++b) for (int i = 0; i < 8; ++i) out[i] += (double)v[(size_t)b * 8 + i] / h->num_grid;
    }
  }
  h->want_prof = enable != 0;
  return CHB_OK;
}

int64_t chb_kernel_launch_count(const chb_handle* h) { return h ? h->launches : 0; }

int chb_last_timings(const chb_handle* h, double out[8]) {
  if (!h || !out) return CHB_ERR_INVALID;
  DevGuard _dg(h->cfg.device);
  if (cudaEventSynchronize(h->ev[4]) != cudaSuccess) { cudaGetLastError(); return CHB_ERR_STATE; }
  for (int i = 0; i < 8; ++i) out[i] = 0.0;
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); return CHB_ERR_STATE; }
    out[i] = ms;
  }
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[1], h->evz) == cudaSuccess) out[4] = ms; else cudaGetLastError();
  if (cudaEventElapsedTime(&ms, h->evz, h->ev[2]) == cudaSuccess) out[5] = ms; else cudaGetLastError();
  return CHB_OK;
}

}  // extern "C"
