// tables.cu -- per-hyper-point interpolation tables (kernel 0).
//
// One CTA per hyper-point rebuilds what `population.update(**hyper_lambdas)` rebuilds in the
// reference for every likelihood call:
//   cosmology: z_grid_interp = [0] U logspace(-10, log10 z_max, res-1), integral_invE_interp =
//              cumtrapz(1/E, z)                               (population/cosmo.py:43-46)
//              + the dL(z) table z_from_dGW inverts            (cosmo.py:260-264)
//   mass:      m_grid = logspace(log10 m_low, log10 m_high, res), cdf_m2_conditioned =
//              cumtrapz(p2(m_grid | m1 = m_high)), norm_p_m1 = trapz(p1(m_grid))   (mass.py:45-52)
// plus the scalar constants (HC row) the other kernels reuse.  fp64 throughout; the two prefix
// sums are sequential (one thread each, different warps) so they associate exactly like
// numpy's cumsum -- ~25 us per CTA, all hyper-points in parallel, negligible next to the KDE.
#include "common.cuh"

__global__ void __launch_bounds__(256)
build_tables_kernel(ModelCfg mc, int n_hyper, const double* __restrict__ hyper, double* __restrict__ tabs,
                    double* __restrict__ HCg) {
  extern __shared__ double sm[];
  const int rc = mc.lay.rc, rm = mc.lay.rm;
  double* zs = sm;            // rc
  double* ys = zs + rc;       // rc : 1/E, then integral_invE in place
  double* p2 = ys + rc;       // rm : secondary pdf at m1 = m_high, then cdf in place
  double* p1 = p2 + rm;       // rm : primary pdf
  double* ms = p1 + rm;       // rm : m_grid
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];

  const int h = blockIdx.x;
  if (h >= n_hyper) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < CHB_NPAR) P[tid] = hyper[(size_t)h * CHB_NPAR + tid];
  if (tid < CHB_NHC) HC[tid] = 0.0;
  __syncthreads();

  if (tid == 0) {
    HC[HC_DH] = 299792.458e-3 / P[CHB_P_H0];
    HC[HC_ODE0] = 1.0 - P[CHB_P_OM0] - P[CHB_P_OR0] - P[CHB_P_OK0];
    HC[HC_SQRTOK] = sqrt(fabs(P[CHB_P_OK0] + 1.e-10));
    HC[HC_DE_CONST] = (P[CHB_P_W0] == -1.0 && P[CHB_P_WA] == 0.0) ? 1.0 : 0.0;
    const double lo = P[CHB_P_MLOW], hi = P[CHB_P_MHIGH];
    HC[HC_LOG10_MLOW] = log10(lo);
    HC[HC_DLOG10_M] = (log10(hi) - log10(lo)) / (double)(rm - 1);
    HC[HC_LOG10_ZSTEP] = (log10(P[CHB_P_ZMAX]) + 10.0) / (double)(rc - 2);
    if (mc.mass_model == CHB_MASS_PLP) {
      double a = -P[CHB_P_ALPHA];
      HC[HC_PL_NORM] = (a == -1.0) ? (log(lo) - log(hi)) : (pow(hi, 1.0 + a) - pow(lo, 1.0 + a)) / (1.0 + a);
      double mu = P[CHB_P_MUG], sg = P[CHB_P_SIGMAG];
      double up = ((mu + 5.0 * sg) - mu) / (sg * sqrt(2.0));
      double dn = (lo - mu) / (sg * sqrt(2.0));
      HC[HC_TG_NORM] = 0.5 * erf(up) - 0.5 * erf(dn);
    } else if (mc.mass_model == CHB_MASS_BPL) {
      double mb = lo + P[CHB_P_BREAKF] * (hi - lo);
      HC[HC_MBREAK] = mb;
      HC[HC_BPL_RATIO] = tpl_notnorm(mb, -P[CHB_P_ALPHA], lo, mb) / tpl_notnorm(mb, -P[CHB_P_ALPHA2], mb, hi);
    }
    double g = P[CHB_P_GAMMA];
    if (mc.rate_model == CHB_RATE_MADAU_DICKINSON || mc.rate_model == CHB_RATE_TRUNC_MD)
      HC[HC_RATE_NORM] = 1.0 + pow(1.0 + P[CHB_P_ZP], -g - P[CHB_P_KAPPA]);
    else if (mc.rate_model == CHB_RATE_TRUNC_PL)
      HC[HC_RATE_NORM] = 1.0 / ((pow(1.0 + P[CHB_P_RZMAX], g + 1.0) - 1.0) / (g + 1.0));
    HC[HC_NORM_P_M1] = 1.0;
  }
  __syncthreads();

  // knots (numpy.linspace: arange*step + start, last = stop; then 10**y)
  const double lzmax = log10(P[CHB_P_ZMAX]);
  const double zstep = (lzmax - (-10.0)) / (double)(rc - 2);
  for (int i = tid; i < rc; i += nt) {
    double z = 0.0;
    if (i > 0) {
      double y = (i - 1 == rc - 2) ? lzmax : __dadd_rn(__dmul_rn((double)(i - 1), zstep), -10.0);
      z = pow(10.0, y);
    }
    zs[i] = z;
    ys[i] = 1.0 / E_at_z(P, HC, z);
  }
  const double l0 = log10(P[CHB_P_MLOW]), l1 = log10(P[CHB_P_MHIGH]);
  const double mstep = (l1 - l0) / (double)(rm - 1);
  for (int i = tid; i < rm; i += nt) {
    double y = (i == rm - 1) ? l1 : __dadd_rn(__dmul_rn((double)i, mstep), l0);
    double m = pow(10.0, y);
    if (i == 0 && P[CHB_P_MGRID_FIRST] != 0.0) m = P[CHB_P_MGRID_FIRST];
    if (i == rm - 1 && P[CHB_P_MGRID_LAST] != 0.0) m = P[CHB_P_MGRID_LAST];
    ms[i] = m;
    p2[i] = secondary_notnorm(mc.mass_model, P, m, P[CHB_P_MHIGH]);
    p1[i] = primary_notnorm(mc.mass_model, P, HC, m);
  }
  __syncthreads();

  // cumtrapz(1/E, z) (utils/math.py:22-26) by warp 0, cumtrapz(p2) and trapz(p1) by warp 1: every lane sums the
  // trapezoid increments of a contiguous segment left to right, the segment totals are scanned with shuffles and each
  // lane then writes its running sums starting from its offset.  (One thread walking the ~1000-knot tables alone was
  // 80 us per step, the largest per-step cost that does not shrink when the events are sharded over GPUs; the
  // association differs from a strictly sequential cumsum by a few ulp.)
  {
    const int lane = tid & 31, warp = tid >> 5;
    if (warp < 2) {
      double* y = (warp == 0) ? ys : p2;
      const double* x = (warp == 0) ? zs : ms;
      const int n = (warp == 0) ? rc : rm;
      const int seg = (n - 1 + 31) / 32;
      const int a0 = min(n, 1 + lane * seg), a1 = min(n, a0 + seg);
      const double first_prev = (a0 < n) ? y[a0 - 1] : 0.0;      // the ORIGINAL value left of the segment
      double tot = 0.0, nrm = 0.0, prev = first_prev;
      for (int i = a0; i < a1; ++i) {
        const double cur = y[i], dx = x[i] - x[i - 1];
        tot += 0.5 * (prev + cur) * dx;
        if (warp == 1) nrm += dx * (p1[i] + p1[i - 1]) / 2.0;
        prev = cur;
      }
      double off = tot;                                        // inclusive scan of the segment totals -> exclusive offset
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
      off -= tot;
      if (warp == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0) HC[HC_NORM_P_M1] = nrm;
      }
      __syncwarp();                                            // every lane has read its originals
      double acc = off;
      prev = first_prev;
      for (int i = a0; i < a1; ++i) {
        const double cur = y[i];
        acc += 0.5 * (prev + cur) * (x[i] - x[i - 1]);
        y[i] = acc;
        prev = cur;
      }
      if (lane == 0) y[0] = 0.0;
    }
  }
  __syncthreads();

  double* T = tabs + (size_t)h * mc.lay.total();
  for (int i = tid; i < rc; i += nt) {
    double z = zs[i];
    double dCt = dCt_from_dCr(P, HC, HC[HC_DH] * ys[i]);
    double dL = dCt * (1.0 + z);
    if (mc.cosmo_model == CHB_COSMO_MG_FLRW) dL *= Xi_at_z(P, z);
    T[mc.lay.off_zg() + i] = z;
    T[mc.lay.off_iinv() + i] = ys[i];
    T[mc.lay.off_dLt() + i] = dL;
  }
  for (int i = tid; i < rm; i += nt) {
    T[mc.lay.off_mg() + i] = ms[i];
    T[mc.lay.off_cdf() + i] = p2[i];
  }
  // ---- fp32 fast-path block: packed float4 rows + float-bits LUT -----------------------------
  {
    float4* zi4 = reinterpret_cast<float4*>(T + mc.lay.off_f32() + mc.lay.f32_zi4());
    float4* dl4 = reinterpret_cast<float4*>(T + mc.lay.off_f32() + mc.lay.f32_dl4());
    float4* cd4 = reinterpret_cast<float4*>(T + mc.lay.off_f32() + mc.lay.f32_cd4());
    unsigned short* lut = reinterpret_cast<unsigned short*>(T + mc.lay.off_f32() + mc.lay.f32_lut());
    const double* dLg = T + mc.lay.off_dLt();       // written above by this CTA
    __syncthreads();
    for (int i = tid; i < rc; i += nt) {
      const int j = min(i + 1, rc - 1);
      const double z0 = zs[i], z1 = zs[j], d0 = dLg[i], d1 = dLg[j];
      const double si = (z1 > z0) ? (ys[j] - ys[i]) / (z1 - z0) : 0.0;
      const double sd = (d1 != d0) ? (z1 - z0) / (d1 - d0) : 0.0;
      zi4[i] = make_float4((float)z0, (float)ys[i], (float)si, (float)z1);
      dl4[i] = make_float4((float)d0, (float)z0, (float)sd, (float)d1);
    }
    for (int i = tid; i < rm; i += nt) {
      const int j = min(i + 1, rm - 1);
      const double m0 = ms[i], m1 = ms[j];
      const double sc = (m1 > m0) ? (p2[j] - p2[i]) / (m1 - m0) : 0.0;
      cd4[i] = make_float4((float)m0, (float)p2[i], (float)sc, (float)m1);
    }
    // LUT over float-bits buckets of dL: lut[b] = last knot k with dLt[k] <= lower edge of bucket b
    const unsigned b0 = __float_as_uint((float)dLg[1]) >> CHB_LUT_SHIFT;
    const unsigned b1 = __float_as_uint((float)dLg[rc - 1]) >> CHB_LUT_SHIFT;
    const bool lut_ok = (dLg[1] > 0.0) && (b1 >= b0) && (b1 - b0 + 2 <= CHB_LUT_CAP) && (rc <= 65535);
    const int nb = lut_ok ? (int)(b1 - b0 + 2) : 0;
    for (int b = tid; b < nb; b += nt) {
      const double edge = (double)__uint_as_float((b0 + (unsigned)b) << CHB_LUT_SHIFT);
      int k = upper_index(dLg, rc, edge) - 1;       // dLt[k] <= edge < dLt[k+1] (clamped)
      if (b == 0) k = 0;
      lut[b] = (unsigned short)max(0, min(k, rc - 2));
    }
    if (tid == 0 && !lut_ok) lut[0] = 0;            // the scan then starts at the first knot
    if (tid == 0) {
      float* FC = reinterpret_cast<float*>(T + mc.lay.off_f32() + mc.lay.f32_fc());
      const double inv_norm = 1.0 / HC[HC_NORM_P_M1];
      for (int i = 0; i < CHB_NFC; ++i) FC[i] = 0.f;
      FC[FC_LG2_M0] = (float)log2(ms[0]);
      FC[FC_INV_LG2_MSTEP] = (float)((double)(rm - 1) / (log2(ms[rm - 1]) - log2(ms[0])));
      FC[FC_LO] = (float)P[CHB_P_MLOW]; FC[FC_HI] = (float)P[CHB_P_MHIGH];
      FC[FC_NEG_ALPHA] = (float)(-P[CHB_P_ALPHA]); FC[FC_BETA] = (float)P[CHB_P_BETA]; FC[FC_DM] = (float)P[CHB_P_DELTAM];
      if (mc.mass_model == CHB_MASS_PLP) {
        const double lam = P[CHB_P_LAMBDAP], sg = P[CHB_P_SIGMAG];
        FC[FC_KA] = (float)log2((1.0 - lam) / HC[HC_PL_NORM] * inv_norm);
        FC[FC_KG] = (float)log2(lam / (sg * 2.5066282746310002 * HC[HC_TG_NORM]) * inv_norm);
        FC[FC_MU] = (float)P[CHB_P_MUG]; FC[FC_G_HI] = (float)(P[CHB_P_MUG] + 5.0 * sg);
        FC[FC_G_C] = (float)(-1.4426950408889634 / (2.0 * sg * sg));
      } else {
        FC[FC_KA] = (float)log2(inv_norm);
        if (mc.mass_model == CHB_MASS_BPL) {
          FC[FC_KG] = (float)log2(HC[HC_BPL_RATIO] * inv_norm);
          FC[FC_MB] = (float)HC[HC_MBREAK]; FC[FC_NEG_ALPHA2] = (float)(-P[CHB_P_ALPHA2]);
        }
      }
      FC[FC_CDL_X] = (float)ms[rm - 1]; FC[FC_CDL_Y] = (float)p2[rm - 1];
      FC[FC_LUT_B0] = __int_as_float((int)b0); FC[FC_LUT_NB] = __int_as_float(nb);
      FC[FC_Z_TOP] = (float)zs[rc - 1];
      HC[HC_LUT_B0] = (double)b0;
      HC[HC_LUT_NB] = (double)nb;
      HC[HC_LG2_M0] = log2(ms[0]);
      HC[HC_INV_LG2_MSTEP] = (double)(rm - 1) / (log2(ms[rm - 1]) - log2(ms[0]));
      HC[HC_LG2_Z1] = log2(zs[1]);
      HC[HC_INV_LG2_ZSTEP] = (double)(rc - 2) / (log2(zs[rc - 1]) - log2(zs[1]));
    }
  }
  if (tid == 0 && mc.catalog_kind == 1) {   // fR = Vc(z_hi) - Vc(z_lo)  (completeness.py:54-58)
    double dlo = dCt_from_dCr(P, HC, HC[HC_DH] * interp_clamped(mc.compl_z_lo, zs, ys, rc));
    double dhi = dCt_from_dCr(P, HC, HC[HC_DH] * interp_clamped(mc.compl_z_hi, zs, ys, rc));
    HC[HC_FR] = Vc_from_dCt(P, HC, dhi) - Vc_from_dCt(P, HC, dlo);
  }
  __syncthreads();
  if (tid < CHB_NHC) HCg[(size_t)h * CHB_NHC + tid] = HC[tid];
}

cudaError_t launch_build_tables(const ModelCfg& mc, int n_hyper, const double* d_hyper, double* d_tabs,
                                double* d_HC, cudaStream_t s) {
  size_t smem = (size_t)(2 * mc.lay.rc + 3 * mc.lay.rm) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(build_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  build_tables_kernel<<<n_hyper, 256, smem, s>>>(mc, n_hyper, d_hyper, d_tabs, d_HC);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// element-wise plug-in functions for one parameter row (chb_model_eval)
__global__ void model_eval_kernel(ModelCfg mc, int which, const double* __restrict__ P, const double* __restrict__ T,
                                  const double* __restrict__ HC, long long n, const double* __restrict__ a,
                                  const double* __restrict__ b, const double* __restrict__ c, double* __restrict__ out) {
  const int rc = mc.lay.rc, rm = mc.lay.rm;
  const double* zg = T + mc.lay.off_zg();
  const double* iinv = T + mc.lay.off_iinv();
  const double* dLt = T + mc.lay.off_dLt();
  const double* mg = T + mc.lay.off_mg();
  const double* cdf = T + mc.lay.off_cdf();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double x = a[i], r = 0.0;
    switch (which) {
      case CHB_F_E_AT_Z: r = E_at_z(P, HC, x); break;
      case CHB_F_DCT_AT_Z: r = dCt_from_dCr(P, HC, HC[HC_DH] * interp_clamped(x, zg, iinv, rc)); break;
      case CHB_F_DL_AT_Z: {
        r = dCt_from_dCr(P, HC, HC[HC_DH] * interp_clamped(x, zg, iinv, rc)) * (1.0 + x);
        if (mc.cosmo_model == CHB_COSMO_MG_FLRW) r *= Xi_at_z(P, x);
      } break;
      case CHB_F_Z_FROM_DGW: r = interp_clamped(x, dLt, zg, rc); break;
      case CHB_F_DDLDZ_AT_Z:
      case CHB_F_DVCDZ_AT_Z:
      case CHB_F_VC_AT_Z: {
        double dCt = b ? dL2dCt(mc.cosmo_model, P, b[i], x)
                       : dCt_from_dCr(P, HC, HC[HC_DH] * interp_clamped(x, zg, iinv, rc));
        if (which == CHB_F_VC_AT_Z) r = Vc_from_dCt(P, HC, dCt);
        else {
          double Ez = E_at_z(P, HC, x);
          r = (which == CHB_F_DDLDZ_AT_Z) ? ddLdz_from(mc.cosmo_model, P, HC, x, dCt, Ez) : dVcdz_from(HC, dCt, Ez);
        }
      } break;
      case CHB_F_P_M1M2: r = p_m1m2(mc.mass_model, P, HC, mg, cdf, rm, x, b[i]); break;
      case CHB_F_P_M1_NOTNORM: r = primary_notnorm(mc.mass_model, P, HC, x); break;
      case CHB_F_MERGER_RATE: r = merger_rate(mc.rate_model, P, HC, x); break;
      case CHB_F_POP_RATE_DET_INJ: {   // pop_wrapper.py:102-111 ; a=m1det b=m2det c=dL
        double dL = c[i];
        double z = interp_clamped(dL, dLt, zg, rc);
        double m1 = x / (1.0 + z), m2 = b[i] / (1.0 + z);
        double dCt = dL2dCt(mc.cosmo_model, P, dL, z);
        double Ez = E_at_z(P, HC, z);
        double pz = dVcdz_from(HC, dCt, Ez) * (merger_rate(mc.rate_model, P, HC, z) / (1.0 + z));
        double dN = P[CHB_P_R0] * p_m1m2(mc.mass_model, P, HC, mg, cdf, rm, m1, m2) * pz;
        double jac = fabs(ddLdz_from(mc.cosmo_model, P, HC, z, dCt, Ez)) * ((1.0 + z) * (1.0 + z));
        r = dN / jac;
      } break;
      default: r = nan(""); break;
    }
    out[i] = r;
  }
}

cudaError_t launch_model_eval(const ModelCfg& mc, int which, const double* d_params, const double* d_tabs,
                              const double* d_HC, long long n, const double* a, const double* b, const double* c,
                              double* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int block = 256;
  long long grid = (n + block - 1) / block;
  if (grid > 148 * 16) grid = 148 * 16;
  model_eval_kernel<<<(int)grid, block, 0, s>>>(mc, which, d_params, d_tabs, d_HC, n, a, b, c, out);
  return cudaGetLastError();
}
