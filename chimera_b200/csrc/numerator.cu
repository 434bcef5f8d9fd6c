// numerator.cu -- fused per-(event, hyper-point) likelihood numerator (kernels 2-4 of DESIGN.md).
//
// One CTA evaluates one unit = log \int K_gw,i(z,Omega|lambda) p_gal(z,Omega|lambda) psi(z)/(1+z) dz dOmega
// for event i and hyper-point lambda, entirely on chip:
//   stage 1  population reweighting of the event's posterior samples
//            (pop_wrapper.py:67-80: z_from_dGW + p_m1m2 / pe_prior)              -> z_j, w_j in smem
//   stage 2  per-event KDE of the reweighted samples on the effective z grid
//            (likelihood.py:105-260, utils/math.py:32-89,154-229), four variants:
//            1-D, 'approximate', 'marginalized' (per pixel), 'full' (3-D whitened Gaussian)
//   stage 3  z-integral against p_gal * psi/(1+z) / (ddL/dz (1+z)^2), pixel sum, log,
//            nan_to_num  (likelihood.py:266-301, pop_wrapper.py:82-90, catalog.py:197-203)
// p_gw never touches HBM unless the caller asks for it (debug output).  The hyper-point's
// tables arrive with one TMA bulk copy.  Grid is persistent (CTAs stride over units, hyper-point
// fastest so that CTAs running concurrently share the same event's samples in L2).
//
// This file holds the fp64 reference-faithful path plus the fp32 pair-sum inner loops
// (CHB_FP32): samples are centred and pre-scaled per unit in fp64, the G x N pair sum runs in
// fp32 with one MUFU.EX2 per pair (Gaussian) or 3 FP32 ops (Epanechnikov), grid points held in
// registers, samples broadcast from shared memory, warp-shuffle + fp64 cross-warp reduction.
#include "common.cuh"
#include "models_f32.cuh"
#include "kde_f32.cuh"
#include "kde_win64.cuh"
#include "stage.cuh"

#define NUM_THREADS 512
#define NUM_WARPS (NUM_THREADS / 32)

int numerator_block_threads() { return NUM_THREADS; }

// shared-memory plan (offsets in doubles), identical on host and device
struct SmemPlan {
  int tab, zgrid, tw, dV, rj, jac, pgw, eg, dens, bc, bs, xwb, part, red, stage, total;
};
__host__ __device__ inline SmemPlan make_plan(int tab_total, int Nz, int B, int Ns, int kind, bool stage_in_smem) {
  SmemPlan p;
  int o = 0;
  p.tab = o; o += tab_total;
  p.zgrid = o; o += Nz;
  p.tw = o; o += Nz;
  p.dV = o; o += Nz;
  p.rj = o; o += Nz;
  p.jac = o; o += Nz;
  p.pgw = o; o += Nz;
  p.eg = o; o += Nz;
  p.dens = o; o += Nz;
  p.bc = o; o += B;
  p.bs = o; o += B;
  p.xwb = o; o += B;                                   // float2 per bin (fp32 mode)
  p.part = o; o += (NUM_WARPS * Nz + 1) / 2;            // float [NUM_WARPS][Nz] cross-warp partials
  p.red = o; o += 64;
  o = (o + 1) & ~1;
  p.stage = o;
  if (stage_in_smem) o += (kind == CHB_PGW_FULL ? 4 : 2) * Ns;
  p.total = o;
  return p;
}
size_t numerator_smem_bytes(const NumArgs& a, bool stage_in_smem) {
  return (size_t)make_plan(a.fp_mode == CHB_FP32 ? a.mc.lay.f32_core() : a.mc.lay.f64_total(), a.Nz, a.binning ? a.num_bins : 0, a.Ns, a.kind, stage_in_smem).total * sizeof(double);
}
long long numerator_scratch_doubles(const NumArgs& a) {
  return (long long)(a.kind == CHB_PGW_FULL ? 4 : 2) * a.Ns;
}

// ------------------------------------------------------------------------------------------
// 1-D KDE pair sums.  dens[g] = scale * sum_j w_j K((eg[g]-x_j)/bw)
// fp64: warp per grid point, lanes stride over the data set.
__device__ __forceinline__ void kde1d_f64(const double* __restrict__ x, const double* __restrict__ w, int n,
                                          const double* __restrict__ eg, int G, double bw, int kernel,
                                          double scale, double* __restrict__ dens) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double inv_bw = 1.0 / bw;
  for (int g = warp; g < G; g += NUM_WARPS) {
    const double gv = eg[g];
    double acc = 0.0;
    if (kernel == CHB_KERNEL_GAUSS) {
      for (int j = lane; j < n; j += 32) {
        double u = (gv - x[j]) / bw;
        acc += w[j] * (exp(-0.5 * u * u) / 2.5066282746310002);
      }
    } else {
      for (int j = lane; j < n; j += 32) {
        double u = (gv - x[j]) / bw;
        double kv = (fabs(u) <= 1.0) ? 0.75 * (1.0 - u * u) : 0.0;
        acc += w[j] * kv;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dens[g] = acc * scale;
  }
  (void)inv_bw;
}

// fp32 pair sums: data (x', w') as float2 in shared memory, x' = (x - c) * s pre-scaled so that the
// Gaussian is 2^-(g'-x')^2 (s = sqrt(log2(e)/2)/bw) and the Epanechnikov support is |g'-x'|<=1
// (s = 1/bw); centring and scaling happen in fp64 before the cast.  Every lane keeps R grid points
// in registers; all lanes of a warp read the same sample (one broadcast LDS.64 per R pairs); the 16
// warps split the samples; per-warp partial sums are combined in fp64.  Per pair the Gaussian costs
// FADD + FMUL + MUFU.EX2 + FFMA: the loop is bound by the MUFU pipe (16 ex2/clk/SM).
// One entry for both arithmetic modes.  x/w: fp64 data set (n entries); in fp32 mode it is converted
// into `xw` (which may alias x: element j of xw overlays element j of x) and the pair sums run in
// fp32.  dens[g] = scale_pdf * sum_j (w_j/W) K((eg[g]-x_j)/bw) / bw, K including its normalisation.
__device__ __forceinline__ void kde1d_any(int fp_mode, double* x, const double* w, int n, float2* xw,
                                          const double* __restrict__ eg, int G, double bw, double W, int kernel,
                                          double scale_pdf, float* part, double* dens) {
  if (fp_mode == CHB_FP64) {
    kde1d_f64(x, w, n, eg, G, bw, kernel, scale_pdf / (W * bw), dens);
    return;
  }
  const double c = 0.5 * (eg[0] + eg[G - 1]);
  const double s = (kernel == CHB_KERNEL_GAUSS) ? 0.8493218002880191 / bw : 1.0 / bw;   // sqrt(log2(e)/2)
  const double knorm = (kernel == CHB_KERNEL_GAUSS) ? 0.3989422804014327 : 0.75;
  const double invW = 1.0 / W;
  for (int j = threadIdx.x; j < n; j += NUM_THREADS) {
    const double xv = x[j], wv = w[j];
    xw[j] = make_float2((float)((xv - c) * s), (float)(wv * invW));
  }
  __syncthreads();
  kde1d_f32<NUM_WARPS>(xw, n, eg, G, c, s, kernel, scale_pdf * knorm / bw, part, dens);
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double nan_to_num_log(double like) {
  double l = log(like);                  // likelihood.py:296-297: jnp.nan_to_num(log, nan=-inf)
  if (isnan(l)) return -INFINITY;        //   NaN -> -inf ; -inf -> -DBL_MAX ; +inf -> +DBL_MAX
  if (isinf(l)) return l > 0 ? CHB_DBL_MAX : -CHB_DBL_MAX;
  return l;
}

// phase profile (only when a.prof != NULL): thread 0 stamps the SM clock at phase boundaries
#define PHASE(i) do { if (a.prof && tid == 0) { long long _t = clock64(); pacc[i] += (unsigned long long)(_t - tlast); tlast = _t; } } while (0)

template <bool STAGE_SMEM, bool F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
numerator_kernel(const NumArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ double L[8];   // full-3D whitening: L00 L10 L11 L20 L21 L22, log_norm, #masked grid points

  const TableLayout lay = a.mc.lay;
  const int Ns = a.Ns, Nz = a.Nz, Pp = a.P, B = a.binning ? a.num_bins : 0;
  constexpr bool in_smem = STAGE_SMEM;
  constexpr int fp_mode = F32 ? CHB_FP32 : CHB_FP64;
  const int tab_total = F32 ? lay.f32_core() : lay.f64_total();
  const SmemPlan pl = make_plan(tab_total, Nz, B, Ns, a.kind, in_smem);
  double* tab = sm + pl.tab;
  double* zgrid = sm + pl.zgrid;
  double* tw = sm + pl.tw;
  double* dV = sm + pl.dV;
  double* rj = sm + pl.rj;
  double* jac = sm + pl.jac;
  double* pgw = sm + pl.pgw;
  double* eg = sm + pl.eg;
  double* dens = sm + pl.dens;
  double* bc = sm + pl.bc;
  double* bs = sm + pl.bs;
  float2* xwb = reinterpret_cast<float2*>(sm + pl.xwb);
  float* part = reinterpret_cast<float*>(sm + pl.part);
  double* red = sm + pl.red;
  double* stage;
  if constexpr (STAGE_SMEM) stage = sm + pl.stage; else stage = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
  double* zs = stage;
  double* ws = stage + Ns;
  double* y1 = stage + 2 * Ns;   // full only
  double* y2 = stage + 3 * Ns;

  const double* zg = tab + lay.off_zg();
  const double* iinv = tab + lay.off_iinv();
  const double* dLt = tab + lay.off_dLt();
  const double* mg = tab + lay.off_mg();
  const double* cdf = tab + lay.off_cdf();
  const int rc = lay.rc, rm = lay.rm;
  const int cm = a.mc.cosmo_model, mm = a.mc.mass_model, rmod = a.mc.rate_model;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tab_bytes = (uint32_t)(tab_total * sizeof(double));
  const bool pixelated = (a.kind != CHB_PGW_1D);
  const bool has_cat = (a.mc.catalog_kind == 1);

  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  unsigned long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();

  const long long units = (long long)a.Nev * a.n_hyper;
  for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int ev = (int)(unit / a.n_hyper), h = (int)(unit % a.n_hyper);
    __syncthreads();                       // everyone is done with the previous unit's smem
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bar, tab_bytes);
      bulk_g2s(tab, a.tabs + (size_t)h * lay.total() + (F32 ? lay.off_f32() : 0), tab_bytes, &bar);
    }
    if (tid < CHB_NPAR) P[tid] = a.hyper[(size_t)h * CHB_NPAR + tid];
    if (tid >= 64 && tid < 64 + CHB_NHC) HC[tid - 64] = a.HC[(size_t)h * CHB_NHC + tid - 64];
    const double* zgr = a.zgrids + (size_t)ev * Nz;
    for (int k = tid; k < Nz; k += NUM_THREADS) zgrid[k] = zgr[k];
    __syncthreads();
    mbar_wait(&bar, phase);
    phase ^= 1;
    PHASE(0);   // table staging + parameter loads

    // ---- z-grid quantities: Jacobian, dVc/dz, psi/(1+z), trapezoid weights --------------
    F32Consts fc;
    if constexpr (F32) fc = make_f32_consts(a.mc, P, HC, tab);
    for (int k = tid; k < Nz; k += NUM_THREADS) {
      const double z = zgrid[k];
      double ii;
      if constexpr (F32) ii = (double)iinv_at_z_f32(fc, (float)z); else ii = interp_clamped(z, zg, iinv, rc);
      const double dCt = dCt_from_dCr(P, HC, HC[HC_DH] * ii);
      const double Ez = E_at_z(P, HC, z);
      jac[k] = ddLdz_from(cm, P, HC, z, dCt, Ez) * ((1.0 + z) * (1.0 + z));   // likelihood.py:272,289
      dV[k] = dVcdz_from(HC, dCt, Ez);
      rj[k] = merger_rate(rmod, P, HC, z) / (1.0 + z);                        // pop_wrapper.py:85
      const double zl = (k > 0) ? zgrid[k - 1] : z, zr = (k < Nz - 1) ? zgrid[k + 1] : z;
      tw[k] = 0.5 * (zr - zl);                                                // trapezoid rule as a dot product
      if constexpr (F32) rj[k] = rj[k] * tw[k] / jac[k];                      // one division per grid point, not per pixel
    }

    PHASE(1);   // z-grid Jacobian / dVc/dz / rate
    // ---- stage 1: reweighting ----------------------------------------------------------
    const size_t so = (size_t)ev * Ns;
    double s1 = 0.0, s2 = 0.0, sz = 0.0, zmn = INFINITY, zmx = -INFINITY;
    if constexpr (F32) {
      const float4* s4 = a.s4 + so;
      const float2* l2 = a.l2 + so;
#pragma unroll 2
      for (int j = tid; j < Ns; j += NUM_THREADS) {
        const float4 sv = __ldg(s4 + j);
        const float2 lv = __ldg(l2 + j);
        const float zf = z_from_dL_f32(fc, sv.x);
        const float opz = 1.f + zf;
        const float r = rcpf_(opz), lz = lg2f_(opz);
        const float wf = weight_f32(fc, sv.y * r, sv.z * r, lv.x - lz, lv.y - lz, sv.w);
        const double z = (double)zf, w = (double)wf;
        zs[j] = z;
        ws[j] = w;
        s1 += w; s2 += w * w; sz += z;
        zmn = fmin(zmn, z); zmx = fmax(zmx, z);
      }
    } else {
      // table intervals without binary searches (bit-identical interpolants): float-bits LUT of the fp32 block for
      // dL -> z (read through L1), log-spaced direct index for the conditional cdf
      const unsigned short* lut = reinterpret_cast<const unsigned short*>(a.tabs + (size_t)h * lay.total() + lay.off_f32() + lay.f32_lut());
      const int lut_b0 = (int)HC[HC_LUT_B0], lut_nb = (int)HC[HC_LUT_NB];
      for (int j = tid; j < Ns; j += NUM_THREADS) {
        const double dL = __ldg(a.dL + so + j);
        const double z = interp_at(dL, dLt, zg, rc, upper_index_lut(dLt, rc, dL, lut, lut_b0, lut_nb));
        const double opz = 1.0 + z;
        const double m1 = __ldg(a.m1d + so + j) / opz, m2 = __ldg(a.m2d + so + j) / opz;
        const double w = p_m1m2_logidx(mm, P, HC, mg, cdf, rm, m1, m2) / __ldg(a.prior + so + j);
        zs[j] = z;
        ws[j] = w;
        s1 += w; s2 += w * w; sz += z;
        zmn = fmin(zmn, z); zmx = fmax(zmx, z);
      }
    }
    PHASE(2);   // reweighting loop (thread 0's share; includes waiting at the first reduction)
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    sz = block_sum(sz, red);
    zmn = block_min(zmn, red);
    zmx = block_max(zmx, red);
    const double zmean = sz / Ns;
    double sv = 0.0;
    for (int j = tid; j < Ns; j += NUM_THREADS) { double d = zs[j] - zmean; sv += d * d; }
    sv = block_sum(sv, red);
    const double zstd = sqrt(sv / Ns);
    const double norm = s1 / Ns;                  // likelihood.py:111
    const double neff = s1 * s1 / s2;             // likelihood.py:112
    const bool ok = (a.kind == CHB_PGW_FULL) ? !(neff < a.pe_neff) : (neff >= a.pe_neff);

    double* pout = nullptr;
    if (a.p_gw_out) pout = a.p_gw_out + ((size_t)h * a.Nev + ev) * (size_t)(pixelated ? Pp : 1) * Nz;
    const int npix = pixelated ? a.neff_pix[ev] : 1;

    if (!ok) {                                    // lax.cond false branch: zeros
      if (pout) for (int i = tid; i < (pixelated ? Pp : 1) * Nz; i += NUM_THREADS) pout[i] = 0.0;
      if (tid == 0) { a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log(0.0); a.like_raw[(size_t)h * a.Nev + ev] = 0.0; }
      continue;
    }

    // ---- effective grid (likelihood.py:115-123 / 186-190) -----------------------------
    int G = Nz;
    double ustep = 0.0, ulb = 0.0;       // spacing / first point of the effective grid when it is our own linspace
    if (a.kind != CHB_PGW_FULL) {
      if (a.use_cut) {
        G = Nz / 2;
        const double lb = (zmn - a.cut_grid * zstd > 0.0) ? zmn - a.cut_grid * zstd : 1.e-8;
        const double ub = zmx + a.cut_grid * zstd;
        const double step = (ub - lb) / (double)(G - 1);
        ustep = step; ulb = lb;
        for (int i = tid; i < G; i += NUM_THREADS) eg[i] = (i == G - 1) ? ub : __dadd_rn(__dmul_rn((double)i, step), lb);
      } else {
        for (int i = tid; i < G; i += NUM_THREADS) eg[i] = zgrid[i];
      }
    }
    __syncthreads();

    PHASE(3);   // statistics + effective grid
    double like_acc = 0.0;      // per-thread partial of sum_p trapz_k(...)
    const double fR = HC[HC_FR];
    const double* pcat_ev = has_cat ? a.p_cat + (size_t)ev * Pp * Nz : nullptr;
    const double* pcompl_ev = has_cat ? a.P_compl + (size_t)ev * Nz : nullptr;

    if (a.kind == CHB_PGW_1D || a.kind == CHB_PGW_APPROX) {
      // ---- p_gw1d ------------------------------------------------------------------------
      const double* dx = zs;
      const double* dw = ws;
      int dn = Ns;
      double W = s1, Q = s2, dstd = zstd;
      if (a.binning) {                           // utils/math.py:32-46
        const double step = (zmx - zmn) / (double)B;
        for (int i = tid; i < B; i += NUM_THREADS) {
          double e0 = __dadd_rn(__dmul_rn((double)i, step), zmn);
          double e1 = (i + 1 == B) ? zmx : __dadd_rn(__dmul_rn((double)(i + 1), step), zmn);
          bc[i] = (e0 + e1) / 2;
          bs[i] = 0.0;
        }
        __syncthreads();
        for (int j = tid; j < Ns; j += NUM_THREADS) {
          double f = floor((zs[j] - zmn) / (zmx - zmn) * B);
          if (!isnan(f)) atomicAdd(&bs[(int)fmin(fmax(f, 0.0), (double)(B - 1))], ws[j]);
        }
        __syncthreads();
        double t1 = 0, t2 = 0, tc = 0;
        for (int i = tid; i < B; i += NUM_THREADS) { t1 += bs[i]; t2 += bs[i] * bs[i]; tc += bc[i]; }
        W = block_sum(t1, red); Q = block_sum(t2, red);
        const double cmean = block_sum(tc, red) / B;
        double tv = 0;
        for (int i = tid; i < B; i += NUM_THREADS) { double d = bc[i] - cmean; tv += d * d; }
        dstd = sqrt(block_sum(tv, red) / B);
        dx = bc; dw = bs; dn = B;
      }
      const double neff_k = 1.0 / (Q / (W * W));                     // 1/sum((w/W)^2)
      double bw;
      if (a.bw_method == CHB_BW_SCOTT) bw = pow(neff_k, -0.2) * dstd;
      else if (a.bw_method == CHB_BW_SILVERMAN) bw = pow(neff_k * 3.0 / 4.0, -0.2) * dstd;
      else bw = a.bw_value * dstd;
      bool done = false;
      if constexpr (!F32) {
        // fp64 mode, unbinned Gaussian on the uniform effective grid: windowed recurrence over the (sorted) samples
        // (kde_win64.cuh) -- 2 exp2 per 8 pairs and ~1/3 of the pairs, exact to ~1e-15; scratch: chunk tables in pgw,
        // per-warp rows in `part` (NUM_WARPS * G doubles = its NUM_WARPS * Nz floats)
        if (!a.binning && a.kernel == CHB_KERNEL_GAUSS && ustep > 0.0 && a.kde_win_iters > 0 && Nz >= 112 && 2 * G <= Nz) {
          float4* summ = reinterpret_cast<float4*>(pgw);
          int2* win = reinterpret_cast<int2*>(summ + 32);
          done = kde1d_f64_win<NUM_WARPS>(zs, ws, Ns, G, ulb, ustep, bw, W, norm * 0.3989422804014327 / bw, summ, win, pgw + 96,
                                          reinterpret_cast<double*>(part), dens);
        }
      }
      if (!done)
        kde1d_any(fp_mode, const_cast<double*>(dx), dw, dn, a.binning ? xwb : reinterpret_cast<float2*>(zs), eg, G, bw, W,
                  a.kernel, norm, part, dens);
      __syncthreads();
      for (int k = tid; k < Nz; k += NUM_THREADS) pgw[k] = interp_lr(zgrid[k], eg, dens, G, 0.0, 0.0);
      __syncthreads();
      if (a.kind == CHB_PGW_1D) {
        for (int k = tid; k < Nz; k += NUM_THREADS) {
          const double pz = dV[k] * rj[k];
          if constexpr (F32) like_acc += pgw[k] * pz; else like_acc += (pgw[k] * pz / jac[k]) * tw[k];
          if (pout) pout[k] = pgw[k];
        }
      } else {
        const double* gwp = a.gw_pdf + (size_t)ev * Pp;
        if (pout) for (int i = tid; i < Pp * Nz; i += NUM_THREADS) pout[i] = pgw[i % Nz] * gwp[i / Nz];
        if (a.catA) {
          // separable case: sum_p gw_pdf[p] p_gal[p,k] = fR A[k] + (1 - P_compl[k]) dVc/dz[k] B[k] with the
          // hyper-independent pixel sums A, B precomputed once (the trapezoid rule and the pixel sum commute)
          const double* A = a.catA + (size_t)ev * Nz;
          const double* Bk = a.catB + (size_t)ev * Nz;
          for (int k = tid; k < Nz; k += NUM_THREADS) {
            const double pgs = has_cat ? fR * A[k] + (1.0 - pcompl_ev[k]) * dV[k] * Bk[k] : dV[k] * Bk[k];
            if constexpr (F32) like_acc += pgw[k] * pgs * rj[k]; else like_acc += ((pgw[k] * pgs) * rj[k] / jac[k]) * tw[k];
          }
        } else {
          for (int i = tid; i < npix * Nz; i += NUM_THREADS) {
            const int p = i / Nz, k = i - p * Nz;
            const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
            if (pc == -100.0) continue;            // likelihood.py:274 sentinel mask
            const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
            const double pz = pgal * rj[k];
            if constexpr (F32) like_acc += (pgw[k] * gwp[p]) * pz; else like_acc += ((pgw[k] * gwp[p]) * pz / jac[k]) * tw[k];
          }
        }
      }
    } else if (a.kind == CHB_PGW_MARG) {
      // ---- p_gw3dmarg (always Epanechnikov, likelihood.py:192) ----------------------------
      const int* off = a.pix_off + (size_t)ev * (Pp + 2);
      const double* gwp = a.gw_pdf + (size_t)ev * Pp;
      if (pout) for (int i = tid; i < Pp * Nz; i += NUM_THREADS) pout[i] = 0.0;
      for (int p = 0; p < npix; ++p) {
        const int o0 = off[p], o1 = off[p + 1], nin = o1 - o0;
        double t1 = 0, t2 = 0, tz = 0, tm = -INFINITY;
        for (int j = o0 + tid; j < o1; j += NUM_THREADS) { double w = ws[j]; t1 += w; t2 += w * w; tz += zs[j]; tm = fmax(tm, zs[j]); }
        double W = block_sum(t1, red), Q = block_sum(t2, red);
        const double zsum = block_sum(tz, red);
        const double zmax_in = fmax(block_max(tm, red), zmn);       // masked samples sit at min(z)
        const double* dx = zs + o0;
        const double* dw = ws + o0;
        int dn = nin;
        double dstd;
        if (a.binning) {
          const double step = (zmax_in - zmn) / (double)B;
          for (int i = tid; i < B; i += NUM_THREADS) {
            double e0 = __dadd_rn(__dmul_rn((double)i, step), zmn);
            double e1 = (i + 1 == B) ? zmax_in : __dadd_rn(__dmul_rn((double)(i + 1), step), zmn);
            bc[i] = (e0 + e1) / 2;
            bs[i] = 0.0;
          }
          __syncthreads();
          for (int j = o0 + tid; j < o1; j += NUM_THREADS) {
            double f = floor((zs[j] - zmn) / (zmax_in - zmn) * B);
            if (!isnan(f)) atomicAdd(&bs[(int)fmin(fmax(f, 0.0), (double)(B - 1))], ws[j]);
          }
          __syncthreads();
          double u1 = 0, u2 = 0, uc = 0;
          for (int i = tid; i < B; i += NUM_THREADS) { u1 += bs[i]; u2 += bs[i] * bs[i]; uc += bc[i]; }
          W = block_sum(u1, red); Q = block_sum(u2, red);
          const double cmean = block_sum(uc, red) / B;
          double tv = 0;
          for (int i = tid; i < B; i += NUM_THREADS) { double d = bc[i] - cmean; tv += d * d; }
          dstd = sqrt(block_sum(tv, red) / B);
          dx = bc; dw = bs; dn = B;
        } else {
          // std of the masked data set: in-pixel samples + (Ns - nin) copies of min(z)
          const double mm_ = (zsum + (double)(Ns - nin) * zmn) / Ns;
          double tv = 0;
          for (int j = o0 + tid; j < o1; j += NUM_THREADS) { double d = zs[j] - mm_; tv += d * d; }
          tv = block_sum(tv, red) + (double)(Ns - nin) * (zmn - mm_) * (zmn - mm_);
          dstd = sqrt(tv / Ns);
        }
        const double neff_k = 1.0 / (Q / (W * W));
        double bw;
        if (a.bw_method == CHB_BW_SCOTT) bw = pow(neff_k, -0.2) * dstd;
        else if (a.bw_method == CHB_BW_SILVERMAN) bw = pow(neff_k * 3.0 / 4.0, -0.2) * dstd;
        else bw = a.bw_value * dstd;
        // W == 0 (no weight in the pixel) -> w/W = NaN for every sample in the reference
        const double scale = (W != 0.0) ? (norm * gwp[p]) : nan("");
        kde1d_any(fp_mode, const_cast<double*>(dx), dw, dn, a.binning ? xwb : reinterpret_cast<float2*>(zs + o0), eg, G, bw, W,
                  CHB_KERNEL_EPAN, 1.0, part, dens);
        __syncthreads();
        for (int k = tid; k < Nz; k += NUM_THREADS) {
          const double raw = interp_lr(zgrid[k], eg, dens, G, 0.0, 0.0);
          const double inside = (zgrid[k] >= eg[0] && zgrid[k] <= eg[G - 1]);
          const double v = inside ? raw * scale : 0.0;          // interp(left=0,right=0) of kde*... then *norm*gw_pdf
          const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
          if (pout) pout[(size_t)p * Nz + k] = v;
          if (pc == -100.0) continue;
          const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
          if constexpr (F32) like_acc += v * (pgal * rj[k]); else like_acc += (v * (pgal * rj[k]) / jac[k]) * tw[k];
        }
        __syncthreads();
      }
    } else {
      // ---- p_gw3dfull: 3-D whitened Gaussian KDE (likelihood.py:211-260, math.py:154-229) --
      const double* ra = a.ra + so;
      const double* dec = a.dec + so;
      const double W = s1;
      const double Qn = s2 / (W * W);                 // sum of squared normalised weights
      const double neff_k = 1.0 / Qn;
      double factor;
      if (a.bw_method == CHB_BW_SCOTT) factor = pow(neff_k, -1.0 / 7.0);
      else if (a.bw_method == CHB_BW_SILVERMAN) factor = pow(neff_k * 5.0 / 4.0, -1.0 / 7.0);
      else factor = a.bw_value;
      double m0 = 0, m1 = 0, m2 = 0;
      for (int j = tid; j < Ns; j += NUM_THREADS) { double wn = ws[j] / W; m0 += wn * zs[j]; m1 += wn * ra[j]; m2 += wn * dec[j]; }
      m0 = block_sum(m0, red); m1 = block_sum(m1, red); m2 = block_sum(m2, red);
      double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
      for (int j = tid; j < Ns; j += NUM_THREADS) {
        double wn = ws[j] / W, r0 = zs[j] - m0, r1 = ra[j] - m1, r2 = dec[j] - m2;
        c00 += wn * r0 * r0; c01 += wn * r0 * r1; c02 += wn * r0 * r2;
        c11 += wn * r1 * r1; c12 += wn * r1 * r2; c22 += wn * r2 * r2;
      }
      c00 = block_sum(c00, red); c01 = block_sum(c01, red); c02 = block_sum(c02, red);
      c11 = block_sum(c11, red); c12 = block_sum(c12, red); c22 = block_sum(c22, red);
      if (tid == 0) {
        const double dn_ = 1.0 - Qn;
        c00 /= dn_; c01 /= dn_; c02 /= dn_; c11 /= dn_; c12 /= dn_; c22 /= dn_;
        // inverse of the symmetric 3x3 covariance, / factor^2
        const double a00 = c11 * c22 - c12 * c12, a01 = c02 * c12 - c01 * c22, a02 = c01 * c12 - c02 * c11;
        const double a11 = c00 * c22 - c02 * c02, a12 = c01 * c02 - c00 * c12, a22 = c00 * c11 - c01 * c01;
        const double det = c00 * a00 + c01 * a01 + c02 * a02;
        const double f2 = factor * factor;
        const double i00 = a00 / det / f2, i01 = a01 / det / f2, i02 = a02 / det / f2;
        const double i11 = a11 / det / f2, i12 = a12 / det / f2, i22 = a22 / det / f2;
        // lower Cholesky factor of inv_cov
        const double l00 = sqrt(i00), l10 = i01 / l00, l20 = i02 / l00;
        const double l11 = sqrt(i11 - l10 * l10), l21 = (i12 - l20 * l10) / l11;
        const double l22 = sqrt(i22 - l20 * l20 - l21 * l21);
        L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
        L[6] = log(l00) + log(l11) + log(l22) - 1.5 * log(2.0 * CHB_PI);
      }
      __syncthreads();
      const double l00 = L[0], l10 = L[1], l11 = L[2], l20 = L[3], l21 = L[4], l22 = L[5], lognorm = L[6];
      // whiten samples about the weighted mean: y = (x - mu)^T L ; weights normalised in place
      const bool f32 = (fp_mode == CHB_FP32);
      float4* yw = reinterpret_cast<float4*>(y1);      // fp32 mode: (y0,y1,y2,w') per sample over the y1|y2 region
      const double ps = 0.8493218002880191;            // sqrt(log2(e)/2): exp(-d^2/2) = 2^-(ps d)^2
      for (int j = tid; j < Ns; j += NUM_THREADS) {
        const double r0 = zs[j] - m0, r1 = ra[j] - m1, r2 = dec[j] - m2;
        const double w0 = r0 * l00 + r1 * l10 + r2 * l20, w1 = r1 * l11 + r2 * l21, w2 = r2 * l22;
        const double wn = ws[j] / W;
        if (f32) {
          yw[j] = make_float4((float)(w0 * ps), (float)(w1 * ps), (float)(w2 * ps), (float)wn);
        } else {
          zs[j] = w0; y1[j] = w1; y2[j] = w2; ws[j] = wn;
        }
      }
      if (pout) for (int i = tid; i < Pp * Nz; i += NUM_THREADS) pout[i] = 0.0;
      const double zlo = zmn - a.cut_grid * zstd, zhi = zmx + a.cut_grid * zstd;   // likelihood.py:225
      // compact list of the grid points inside the cut window
      int* kmask = reinterpret_cast<int*>(dens);
      if (tid == 0) {
        int c = 0;
        for (int k = 0; k < Nz; ++k) if (zgrid[k] <= zhi && zgrid[k] >= zlo) kmask[c++] = k;
        L[7] = (double)c;
      }
      __syncthreads();
      const int nmask = (int)L[7];
      const double* rap = a.ra_pix + (size_t)ev * Pp;
      const double* dep = a.dec_pix + (size_t)ev * Pp;
      const int npts = npix * nmask;
      if (!f32) {
        for (int i = warp; i < npts; i += NUM_WARPS) {
          const int p = i / nmask, k = kmask[i - p * nmask];
          const double r0 = zgrid[k] - m0, r1 = rap[p] - m1, r2 = dep[p] - m2;
          const double q0 = r0 * l00 + r1 * l10 + r2 * l20, q1 = r1 * l11 + r2 * l21, q2 = r2 * l22;
          double acc = 0.0;
          for (int j = lane; j < Ns; j += 32) {
            const double d0 = zs[j] - q0, d1 = y1[j] - q1, d2 = y2[j] - q2;
            acc += ws[j] * exp(lognorm - 0.5 * (d0 * d0 + d1 * d1 + d2 * d2));
          }
          acc = warp_sum(acc);
          if (lane == 0) {
            const double v = acc * norm;
            if (pout) pout[(size_t)p * Nz + k] = v;
            const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
            if (pc != -100.0) {
              const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
              if constexpr (F32) like_acc += v * (pgal * rj[k]); else like_acc += (v * (pgal * rj[k]) / jac[k]) * tw[k];
            }
          }
        }
      } else {
        // fp32: every lane owns FR evaluation points, the warp streams all samples (broadcast LDS.128);
        // per pair 3 FADD + FMUL + 2 FFMA + MUFU.EX2 + FFMA.
        constexpr int FR = 2;
        const double enorm = exp(lognorm) * norm;
        const int ntiles = (npts + 32 * FR - 1) / (32 * FR);
        for (int t = warp; t < ntiles; t += NUM_WARPS) {
          float q0[FR], q1[FR], q2[FR], acc[FR];
          int pk[FR];
#pragma unroll
          for (int r = 0; r < FR; ++r) {
            const int i = t * 32 * FR + r * 32 + lane;
            pk[r] = -1; acc[r] = 0.f;
            q0[r] = q1[r] = q2[r] = 1.0e18f;
            if (i < npts) {
              const int p = i / nmask, k = kmask[i - p * nmask];
              pk[r] = p * Nz + k;
              const double r0 = zgrid[k] - m0, r1 = rap[p] - m1, r2 = dep[p] - m2;
              q0[r] = (float)((r0 * l00 + r1 * l10 + r2 * l20) * ps);
              q1[r] = (float)((r1 * l11 + r2 * l21) * ps);
              q2[r] = (float)((r2 * l22) * ps);
            }
          }
#pragma unroll 4
          for (int j = 0; j < Ns; ++j) {
            const float4 v = yw[j];
#pragma unroll
            for (int r = 0; r < FR; ++r) {
              const float d0 = v.x - q0[r], d1 = v.y - q1[r], d2 = v.z - q2[r];
              const float e = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
              acc[r] = fmaf(v.w, ex2_ftz(-e), acc[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < FR; ++r) {
            if (pk[r] < 0) continue;
            const int p = pk[r] / Nz, k = pk[r] - p * Nz;
            const double v = (double)acc[r] * enorm;
            if (pout) pout[(size_t)p * Nz + k] = v;
            const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
            if (pc != -100.0) {
              const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
              if constexpr (F32) like_acc += v * (pgal * rj[k]); else like_acc += (v * (pgal * rj[k]) / jac[k]) * tw[k];
            }
          }
        }
      }
    }

    PHASE(4);   // KDE + interpolation + integrand
    const double like = block_sum(like_acc, red);
    if (tid == 0) { a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log(like); a.like_raw[(size_t)h * a.Nev + ev] = like; }
    PHASE(5);   // final reduction + store
  }
  if (a.prof && tid == 0) for (int i = 0; i < 8; ++i) a.prof[(size_t)blockIdx.x * 8 + i] = pacc[i];
}

cudaError_t numerator_configure(size_t smem) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(numerator_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(numerator_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(numerator_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(numerator_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t launch_numerator(const NumArgs& a, int grid, int block, size_t smem, cudaStream_t s) {
  const bool in_smem = (a.scratch_stride == 0), f32 = (a.fp_mode == CHB_FP32);
  if (in_smem && f32) numerator_kernel<true, true><<<grid, block, smem, s>>>(a);
  else if (in_smem) numerator_kernel<true, false><<<grid, block, smem, s>>>(a);
  else if (f32) numerator_kernel<false, true><<<grid, block, smem, s>>>(a);
  else numerator_kernel<false, false><<<grid, block, smem, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Setup kernel for kind 'approximate': collapse the catalogue term over the event's pixels,
//   A[ev,k] = sum_p gw_loc2d_pdf[ev,p] p_cat[ev,p,k],   B[ev,k] = sum_p gw_loc2d_pdf[ev,p] [p_cat[ev,p,k] != -100]
// (p over valid pixels).  p_gw3d = p_gw1d[k] gw_pdf[p] is separable there (likelihood.py:150-154), so
// sum_p trapz_k(p_gw3d p_z / jac) needs only A and B.  One coalesced pass over p_cat, once per run.
__global__ void catalog_collapse_kernel(int Nev, int P, int Nz, const double* __restrict__ p_cat,
                                        const double* __restrict__ gw_pdf, const int* __restrict__ neff_pix,
                                        double* __restrict__ catA, double* __restrict__ catB) {
  const int ev = blockIdx.x;
  const int npix = neff_pix[ev];
  for (int k = threadIdx.x; k < Nz; k += blockDim.x) {
    double sa = 0.0, sb = 0.0;
    for (int p = 0; p < npix; ++p) {
      const double pc = p_cat[((size_t)ev * P + p) * Nz + k];
      const double g = gw_pdf[(size_t)ev * P + p];
      if (pc != -100.0) { sa += g * pc; sb += g; }
    }
    catA[(size_t)ev * Nz + k] = sa;
    catB[(size_t)ev * Nz + k] = sb;
  }
}
cudaError_t launch_catalog_collapse(int Nev, int P, int Nz, const double* p_cat, const double* gw_pdf, const int* neff_pix,
                                   double* catA, double* catB, cudaStream_t s) {
  catalog_collapse_kernel<<<Nev, 128, 0, s>>>(Nev, P, Nz, p_cat, gw_pdf, neff_pix, catA, catB);
  return cudaGetLastError();
}
