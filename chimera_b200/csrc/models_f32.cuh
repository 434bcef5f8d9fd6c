// models_f32.cuh -- fp32 fast path of the population reweighting (CHB_FP32 mode).
//
// Same functions as models.cuh (pop_wrapper.py:67-80: z_from_dGW, theta_det2src, p_m1m2 / pe_prior)
// evaluated in single precision on the FP32 + MUFU pipes instead of the FP64 pipe:
//   * powers x^a as ex2(a * lg2 x) (MUFU.LG2 / MUFU.EX2), log2 of the detector-frame masses
//     precomputed once at upload, so a source-frame mass power costs FADD + FMUL + MUFU;
//   * table look-ups through the packed float4 rows of the fp32 table block (models.cuh):
//     one LDS.128 per candidate interval; the dL -> z inversion finds its interval with a
//     float-bits LUT and a short forward scan instead of an 11-step binary search;
//   * the interpolants are the reference's own piecewise-linear functions on the reference's
//     knots (knots and slopes are rounded from the fp64 tables), so the only deviations from the
//     fp64 path are fp32 roundings (~1e-7 relative per sample); north_star budget for this mode: 1e-3.
#pragma once
#include "models.cuh"

__device__ __forceinline__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#define CHB_LOG2E_F 1.4426950408889634f

// per-hyper-point constants of the fast path, derived from the parameter row / HC row
struct F32Consts {
  // tables
  const float4* zi4; const float4* dl4; const float4* cd4; const unsigned short* lut;
  int rc, rm, nb; unsigned b0;
  float lg2_m0, inv_lg2_mstep, lg2_z1, inv_lg2_zstep;
  float z_last;
  // mass
  int mass_model;
  float lo, hi, neg_alpha, beta, dm, inv_norm_p1;
  float lam, mu, g_hi, g_c, g_pref, inv_plnorm;       // plp
  float mb, neg_alpha2, ratio;                          // bpl
  float4 cd4_last;                                      // last row of cd4 (clamp)
};

__device__ __forceinline__ F32Consts make_f32_consts(const ModelCfg& mc, const double* __restrict__ P,
                                                     const double* __restrict__ HC, const double* __restrict__ f32blk) {
  F32Consts c;
  const TableLayout lay = mc.lay;
  c.zi4 = reinterpret_cast<const float4*>(f32blk + lay.f32_zi4());
  c.dl4 = reinterpret_cast<const float4*>(f32blk + lay.f32_dl4());
  c.cd4 = reinterpret_cast<const float4*>(f32blk + lay.f32_cd4());
  c.lut = reinterpret_cast<const unsigned short*>(f32blk + lay.f32_lut());
  c.rc = lay.rc; c.rm = lay.rm;
  c.nb = (int)HC[HC_LUT_NB]; c.b0 = (unsigned)HC[HC_LUT_B0];
  c.lg2_m0 = (float)HC[HC_LG2_M0]; c.inv_lg2_mstep = (float)HC[HC_INV_LG2_MSTEP];
  c.lg2_z1 = (float)HC[HC_LG2_Z1]; c.inv_lg2_zstep = (float)HC[HC_INV_LG2_ZSTEP];
  c.z_last = (float)P[CHB_P_ZMAX];
  c.mass_model = mc.mass_model;
  c.lo = (float)P[CHB_P_MLOW]; c.hi = (float)P[CHB_P_MHIGH];
  c.neg_alpha = (float)(-P[CHB_P_ALPHA]); c.beta = (float)P[CHB_P_BETA]; c.dm = (float)P[CHB_P_DELTAM];
  c.inv_norm_p1 = (float)(1.0 / HC[HC_NORM_P_M1]);
  c.lam = (float)P[CHB_P_LAMBDAP]; c.mu = (float)P[CHB_P_MUG];
  const double sg = P[CHB_P_SIGMAG];
  c.g_hi = (float)(P[CHB_P_MUG] + 5.0 * sg);
  c.g_c = (float)(-1.4426950408889634 / (2.0 * sg * sg));
  c.g_pref = (float)(1.0 / (sg * 2.5066282746310002 * HC[HC_TG_NORM]));
  c.inv_plnorm = (float)(1.0 / HC[HC_PL_NORM]);
  c.mb = (float)HC[HC_MBREAK]; c.neg_alpha2 = (float)(-P[CHB_P_ALPHA2]); c.ratio = (float)HC[HC_BPL_RATIO];
  c.cd4_last = make_float4((float)P[CHB_P_MHIGH], 0.f, 0.f, 0.f);   // callers with cd4 resident overwrite it
  return c;
}

// z_from_dGW (cosmo.py:260-264): LUT bucket from the float bits, forward scan, one FFMA.
__device__ __forceinline__ float z_from_dL_f32(const F32Consts& c, float dL) {
  int b = (int)(__float_as_uint(dL) >> CHB_LUT_SHIFT) - (int)c.b0;
  b = max(0, min(b, c.nb - 1));
  int k = c.lut[b];
  float4 e = c.dl4[k];
  while (dL >= e.w && k < c.rc - 2) { ++k; e = c.dl4[k]; }
  float z = fmaf(dL - e.x, e.z, e.y);
  if (dL >= e.w) z = c.zi4[c.rc - 1].x;          // beyond the last knot: clamp (np.interp)
  if (dL <= 0.f) z = 0.f;
  return z;
}

// integral_invE at z by direct indexing of the log-spaced grid (cosmo.py:132-133)
__device__ __forceinline__ float iinv_at_z_f32(const F32Consts& c, float z) {
  int k = 0;
  if (z > 1.0e-10f) k = 1 + (int)((lg2f_(z) - c.lg2_z1) * c.inv_lg2_zstep);
  k = max(0, min(k, c.rc - 2));
  float4 e = c.zi4[k];
  if (z < e.x && k > 0) e = c.zi4[--k];          // one-step fix-up for index rounding
  else if (z >= e.w && k < c.rc - 2) e = c.zi4[++k];
  float v = fmaf(z - e.x, e.z, e.y);
  if (z >= e.w) v = c.zi4[c.rc - 1].y;
  return v;
}

// mass.py:255-264
__device__ __forceinline__ float smoothing_f32(float m, float dm, float lo) {
  const float x = m - lo;
  if (x < 0.f) return 0.f;
  if (x > dm) return 1.f;
  const float t = dm * (rcpf_(x) + rcpf_(x - dm));
  return rcpf_(1.f + ex2f_(t * CHB_LOG2E_F));
}

// p_m1m2 / pe_prior (mass.py:334-345, pop_wrapper.py:79).  lg2m1/lg2m2: log2 of the SOURCE-frame masses.
__device__ __forceinline__ float weight_f32(const F32Consts& c, float m1, float m2, float lg2m1, float lg2m2,
                                            float inv_prior) {
  if (!(c.lo <= m1 && m1 <= c.hi)) return 0.f;        // primary support
  float p1;
  if (c.mass_model == CHB_MASS_TPL) {
    p1 = ex2f_(c.neg_alpha * lg2m1);
  } else if (c.mass_model == CHB_MASS_BPL) {
    p1 = (m1 <= c.mb) ? ex2f_(c.neg_alpha * lg2m1) : 0.f;
    if (m1 >= c.mb) p1 += ex2f_(c.neg_alpha2 * lg2m1) * c.ratio;
    p1 *= smoothing_f32(m1, c.dm, c.lo);
  } else {
    const float Ppl = ex2f_(c.neg_alpha * lg2m1) * c.inv_plnorm;
    const float d = m1 - c.mu;
    const float G = (m1 <= c.g_hi) ? ex2f_(c.g_c * d * d) * c.g_pref : 0.f;
    p1 = ((1.f - c.lam) * Ppl + c.lam * G) * smoothing_f32(m1, c.dm, c.lo);
  }
  // p1 == 0 (outside the support, or an fp32 underflow) must give weight 0 even when 1/cdf overflows (0 * inf)
  if (!(p1 > 0.f)) return 0.f;
  if (!(c.lo <= m2 && m2 <= m1)) return 0.f;          // secondary support (tpl_notnorm(m2, beta, m_low, m1))
  float p2 = ex2f_(c.beta * lg2m2);
  if (c.mass_model != CHB_MASS_TPL) p2 *= smoothing_f32(m2, c.dm, c.lo);
  int i = (int)((lg2m1 - c.lg2_m0) * c.inv_lg2_mstep);
  i = max(0, min(i, c.rm - 2));
  const float4 e = c.cd4[i];
  float cdf = fmaf(m1 - e.x, e.z, e.y);
  if (m1 >= c.cd4[c.rm - 1].x) cdf = c.cd4[c.rm - 1].y;
  p2 = p2 * rcpf_(cdf);
  if (p2 != p2) p2 = 0.f;                              // 0/0 -> 0 (mass.py:340)
  return p1 * c.inv_norm_p1 * p2 * inv_prior;
}

// Branch-free variants for software-pipelined loops (several samples in flight per thread): every
// sample executes the same instruction stream, support tests become selects.
__device__ __forceinline__ float smoothing_bf(float m, float dm, float lo) {
  const float x = m - lo;
  // dm/x + dm/(x-dm) = dm (2x - dm) / (x (x - dm)): one reciprocal
  const float t = dm * (2.f * x - dm) * rcpf_(x * (x - dm));
  float S = rcpf_(1.f + ex2f_(t * CHB_LOG2E_F));
  S = (x > dm) ? 1.f : S;
  return (x < 0.f) ? 0.f : S;
}
// SMOOTH = false: the caller has checked that every mass is above m_low + delta_m (smoothing == 1 exactly,
// mass.py:255-264), so the taper is skipped.
template <bool SMOOTH>
__device__ __forceinline__ float weight_bf(const F32Consts& c, float m1, float m2, float lg2m1, float lg2m2,
                                           float inv_prior) {
  const bool in1 = (c.lo <= m1) && (m1 <= c.hi);
  const bool taper = SMOOTH && (c.mass_model != CHB_MASS_TPL);
  float p1;
  if (c.mass_model == CHB_MASS_TPL) {
    p1 = ex2f_(c.neg_alpha * lg2m1);
  } else if (c.mass_model == CHB_MASS_BPL) {
    const float a = (m1 <= c.mb) ? ex2f_(c.neg_alpha * lg2m1) : 0.f;
    const float b = (m1 >= c.mb) ? ex2f_(c.neg_alpha2 * lg2m1) * c.ratio : 0.f;
    p1 = a + b;
  } else {
    const float Ppl = ex2f_(c.neg_alpha * lg2m1) * c.inv_plnorm;
    const float d = m1 - c.mu;
    const float G = (m1 <= c.g_hi) ? ex2f_(c.g_c * d * d) * c.g_pref : 0.f;
    p1 = (1.f - c.lam) * Ppl + c.lam * G;
  }
  if (taper) p1 *= smoothing_bf(m1, c.dm, c.lo);
  float p2 = ex2f_(c.beta * lg2m2);
  if (taper) p2 *= smoothing_bf(m2, c.dm, c.lo);
  int i = (int)((lg2m1 - c.lg2_m0) * c.inv_lg2_mstep);
  i = max(0, min(i, c.rm - 2));
  const float4 e = c.cd4[i];
  float cdf = fmaf(m1 - e.x, e.z, e.y);
  cdf = (m1 >= c.cd4_last.x) ? c.cd4_last.y : cdf;
  p2 = p2 * rcpf_(cdf);
  p2 = (p2 != p2) ? 0.f : p2;                          // 0/0 -> 0 (mass.py:340)
  const bool in2 = (c.lo <= m2) && (m2 <= m1);
  const float w = p1 * c.inv_norm_p1 * p2 * inv_prior;
  return (in1 && in2 && p1 > 0.f) ? w : 0.f;     // p1 == 0: never 0 * inf
}
// dL -> z with a fixed two-step scan (covers every bucket of a monotone table at 32 buckets/octave);
// `more` tells the caller that a longer scan is needed (non-monotone / unusually dense tables).
__device__ __forceinline__ float z_lookup2(const F32Consts& c, float dL, float z_top, int& k, float4& e, bool& more) {
  int b = (int)(__float_as_uint(dL) >> CHB_LUT_SHIFT) - (int)c.b0;
  b = max(0, min(b, c.nb - 1));
  k = c.lut[b];
  e = c.dl4[k];
  if (dL >= e.w && k < c.rc - 2) { ++k; e = c.dl4[k]; }
  if (dL >= e.w && k < c.rc - 2) { ++k; e = c.dl4[k]; }
  more = (dL >= e.w && k < c.rc - 2);
  float z = fmaf(dL - e.x, e.z, e.y);
  z = (dL >= e.w) ? z_top : z;
  return (dL <= 0.f) ? 0.f : z;
}

// ------------------------------------------------------------------------------------------
// cosmology / rate terms of the detector-frame rate (pop_wrapper.py:102-111) in single precision
struct CosmoRateF32 {
  int cosmo_model, rate_model;
  float Om, Or, Ok, Ode, w0, wa, dH, Xi0, n, R0;
  bool de_const;
  float gamma, gk, lg2_opzp, rnorm, rzmax;
};
__device__ __forceinline__ CosmoRateF32 make_cosmo_rate_f32(const ModelCfg& mc, const double* __restrict__ P,
                                                            const double* __restrict__ HC) {
  CosmoRateF32 c;
  c.cosmo_model = mc.cosmo_model; c.rate_model = mc.rate_model;
  c.Om = (float)P[CHB_P_OM0]; c.Or = (float)P[CHB_P_OR0]; c.Ok = (float)P[CHB_P_OK0]; c.Ode = (float)HC[HC_ODE0];
  c.w0 = (float)P[CHB_P_W0]; c.wa = (float)P[CHB_P_WA]; c.dH = (float)HC[HC_DH];
  c.Xi0 = (float)P[CHB_P_XI0]; c.n = (float)P[CHB_P_N]; c.R0 = (float)P[CHB_P_R0];
  c.de_const = HC[HC_DE_CONST] != 0.0;
  c.gamma = (float)P[CHB_P_GAMMA]; c.gk = (float)(P[CHB_P_GAMMA] + P[CHB_P_KAPPA]);
  c.lg2_opzp = (float)log2(1.0 + P[CHB_P_ZP]); c.rnorm = (float)HC[HC_RATE_NORM]; c.rzmax = (float)P[CHB_P_RZMAX];
  return c;
}
// E(z) (cosmo.py:122-130); lz = log2(1+z)
__device__ __forceinline__ float E_at_z_f32(const CosmoRateF32& c, float z, float opz, float lz) {
  const float x2 = opz * opz;
  float de = 1.f;
  if (!c.de_const) de = ex2f_(3.f * (1.f + c.w0 + c.wa * z * rcpf_(opz)) * lz);
  return sqrtf(c.Om * (x2 * opz) + c.Or * (x2 * x2) + c.Ok * x2 + c.Ode * de);
}
// merger_rate (rate.py:96-129)
__device__ __forceinline__ float merger_rate_f32(const CosmoRateF32& c, float z, float lz) {
  const float pl = ex2f_(c.gamma * lz);
  if (c.rate_model == CHB_RATE_POWER_LAW) return pl;
  if (c.rate_model == CHB_RATE_TRUNC_PL) return (z < c.rzmax) ? pl * c.rnorm : 0.f;
  const float v = c.rnorm * pl * rcpf_(1.f + ex2f_(c.gk * (lz - c.lg2_opzp)));
  if (c.rate_model == CHB_RATE_TRUNC_MD) return (z < c.rzmax) ? v : 0.f;
  return v;
}
// dN/dtheta_det / (R0 p_m1m2) for an injection with its ORIGINAL distance dL (pop_wrapper.py:105-110):
// dVc/dz psi/(1+z) / (|ddL/dz| (1+z)^2)
__device__ __forceinline__ float zterm_inj_f32(const CosmoRateF32& c, float z, float opz, float lz, float dL) {
  float Xi = 1.f, dCt = dL * rcpf_(opz);
  if (c.cosmo_model == CHB_COSMO_MG_FLRW) {
    Xi = c.Xi0 + (1.f - c.Xi0) * ex2f_(-c.n * lz);
    dCt = dCt * rcpf_(Xi);
  }
  const float Ez = E_at_z_f32(c, z, opz, lz);
  const float dHE = c.dH * rcpf_(Ez);
  float ddL = dCt + dHE * opz;
  if (c.cosmo_model == CHB_COSMO_MG_FLRW)
    ddL = ddL * Xi + (dCt * opz) * (c.n * (c.Xi0 - 1.f) * ex2f_(-(c.n + 1.f) * lz));
  const float dV = 12.566370614359172f * dHE * dCt * dCt;
  return dV * merger_rate_f32(c, z, lz) * rcpf_(opz) * rcpf_(fabsf(ddL) * opz * opz);
}

// z-grid terms of one (hyper-point, z): {dVc/dz, psi/(1+z) * trapezoid weight / (ddL/dz (1+z)^2)}
// (cosmo.py:188-221,245-257, rate.py:96-129, likelihood.py:272,289).  E, ddL/dz, psi in fp32; the
// comoving distance through the packed zi4 table.
__device__ __forceinline__ float2 zgrid_terms_f32(const F32Consts& fc, const CosmoRateF32& cr, const double* __restrict__ P,
                                                  const double* __restrict__ HC, int cm, double z, double tw) {
  const float zf = (float)z, opz = 1.f + zf, lz = lg2f_(opz);
  const double dCt = dCt_from_dCr(P, HC, HC[HC_DH] * (double)iinv_at_z_f32(fc, zf));
  const float Ez = E_at_z_f32(cr, zf, opz, lz);
  const float dHE = cr.dH * rcpf_(Ez);
  float ddL = (float)dCt + dHE * opz;
  if (cm == CHB_COSMO_MG_FLRW) {
    const float Xi = cr.Xi0 + (1.f - cr.Xi0) * ex2f_(-cr.n * lz);
    ddL = ddL * Xi + ((float)dCt * opz) * (cr.n * (cr.Xi0 - 1.f) * ex2f_(-(cr.n + 1.f) * lz));
  }
  const double dVv = 12.566370614359172 * (double)dHE * dCt * dCt;
  const double ckv = (double)(merger_rate_f32(cr, zf, lz) * rcpf_(opz) * rcpf_(ddL * opz * opz)) * tw;
  return make_float2((float)dVv, (float)ckv);
}

