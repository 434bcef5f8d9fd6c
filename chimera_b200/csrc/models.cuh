// models.cuh -- device-side population models (cosmology / mass / rate) and table interpolation.
//
// B200-native restatement of the plug-in functions the reference dispatches with plum
// (CHIMERA/population/cosmo.py:122-264, mass.py:240-345, rate.py:96-129).  A hyper-point is a
// row of CHB_NPAR doubles (include/chimera_b200.h); per-hyper-point derived scalars live in a
// row of CHB_NHC doubles ("HC") filled by the table kernel (tables.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/chimera_b200.h"

#define CHB_NHC 24
enum {
  HC_DH = 0,        // 299792.458e-3 / H0                       cosmo.py:83-84
  HC_ODE0,          // 1 - Om0 - Or0 - Ok0                       cosmo.py:80-81
  HC_SQRTOK,        // sqrt(|Ok0 + 1e-10|)                       cosmo.py:145
  HC_FR,            // Vc(z_hi) - Vc(z_lo)                       completeness.py:54-58
  HC_NORM_P_M1,     // trapz(p1_notnorm, m_grid)                 mass.py:50-52
  HC_PL_NORM,       // tpl_cdf(-alpha, m_low, m_high)            mass.py:300
  HC_TG_NORM,       // truncated-gaussian normalisation          mass.py:272-274
  HC_MBREAK,        // bpl break mass                            mass.py:291
  HC_BPL_RATIO,     // pl1(m_break)/pl2(m_break)                 mass.py:292-295
  HC_RATE_NORM,     // MD: 1+(1+zp)^(-g-k); trunc PL: 1/norm     rate.py:104,112
  HC_LOG10_MLOW,    // log10(m_low)
  HC_DLOG10_M,      // (log10 m_high - log10 m_low)/(rm-1)
  HC_LOG10_ZSTEP,   // (log10 z_max + 10)/(rc-2)
  HC_DE_CONST,      // 1 if w0==-1 && wa==0 (dark-energy term constant)
  // fp32 fast-path lookup constants (packed tables below)
  HC_LUT_B0,        // float-bits bucket of dLt[1]:  (bits >> CHB_LUT_SHIFT)
  HC_LUT_NB,        // number of buckets in use (0: fast path unavailable for this hyper-point)
  HC_LG2_M0,        // log2(m_grid[0])
  HC_INV_LG2_MSTEP, // (rm-1) / (log2 m_grid[rm-1] - log2 m_grid[0])
  HC_LG2_Z1,        // log2(z_grid_interp[1]) = log2(1e-10)
  HC_INV_LG2_ZSTEP  // (rc-2) / (log2 z_max - log2 1e-10)
};
// single-precision constants of the fused fp32 kernel, one row of CHB_NFC floats per hyper-point (written by
// build_tables_kernel next to the packed tables): normalisations are folded into log2 offsets so that a mass pdf
// is one FFMA + one MUFU.EX2 and cannot leave the fp32 range before it is normalised.
#define CHB_NFC 32
enum {
  FC_LG2_M0 = 0, FC_INV_LG2_MSTEP, FC_LO, FC_HI, FC_NEG_ALPHA, FC_BETA, FC_DM,
  FC_KA,          // log2 of the factor of the first power law:  tpl/bpl 1/norm_p1;  plp (1-lambda)/(pl_norm norm_p1)
  FC_KG,          // plp: log2(lambda / (sigma sqrt(2 pi) tg_norm norm_p1));  bpl: log2(ratio / norm_p1)
  FC_MU, FC_G_HI, FC_G_C, FC_MB, FC_NEG_ALPHA2, FC_CDL_X, FC_CDL_Y, FC_LUT_B0 /* int bits */, FC_LUT_NB /* int bits */,
  FC_Z_TOP
};
#define CHB_LUT_SHIFT 17      // 64 buckets per octave of dL: <= 0.7 knots per bucket at the default resolution, so <= 1 scan step
#define CHB_LUT_CAP 4096      // uint16 entries

#define CHB_PI 3.141592653589793238462643383279502884
#define CHB_DBL_MAX 1.7976931348623157e308

struct ModelIds {
  int cosmo, mass, rate;
};

// ------------------------------------------------------------------------------------------
// small math helpers
// x^y for x > 0 through exp/log: relative error ~ |y ln x| * 2^-53, far inside the 1e-5 budget,
// and ~4x cheaper than pow() in fp64.
__device__ __forceinline__ double pow_pos(double x, double y) { return exp(y * log(x)); }

// index i in [1, n-1] = clip(searchsorted(xp, x, side='right'), 1, n-1)   (jnp.interp)
__device__ __forceinline__ int upper_index(const double* __restrict__ xp, int n, double x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return min(max(lo, 1), n - 1);
}

// jnp.interp(x, xp, fp) with the default edge handling (clamp to fp[0] / fp[n-1]).
__device__ __forceinline__ double interp_at(double x, const double* __restrict__ xp,
                                            const double* __restrict__ fp, int n, int i) {
  double x0 = xp[i - 1], dx = xp[i] - x0;
  double f0 = fp[i - 1], df = fp[i] - f0;
  double f = (fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((x - x0) / dx) * df;
  if (x < xp[0]) f = fp[0];
  if (x > xp[n - 1]) f = fp[n - 1];
  return f;
}
__device__ __forceinline__ double interp_clamped(double x, const double* __restrict__ xp,
                                                 const double* __restrict__ fp, int n) {
  return interp_at(x, xp, fp, n, upper_index(xp, n, x));
}
// jnp.interp(..., left=l, right=r)
__device__ __forceinline__ double interp_lr(double x, const double* __restrict__ xp,
                                            const double* __restrict__ fp, int n, double l, double r) {
  int i = upper_index(xp, n, x);
  double x0 = xp[i - 1], dx = xp[i] - x0;
  double f0 = fp[i - 1], df = fp[i] - f0;
  double f = (fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((x - x0) / dx) * df;
  if (x < xp[0]) f = l;
  if (x > xp[n - 1]) f = r;
  return f;
}

// The same interval as `upper_index` (i - 1 = last knot <= x), found without a binary search:
//  * dL table: the float-bits LUT of the fp32 block gives the last knot at or below the lower edge of the bucket of
//    float_round_down(x) <= x, then a forward scan in the fp64 table (<= 1 step at the default resolution);
//  * log-spaced table (m_grid): direct index from log10(x), then a fix-up scan in either direction.
// Exactly the interval of the binary search, hence bit-identical interpolants.
__device__ __forceinline__ int upper_index_lut(const double* __restrict__ xp, int n, double x,
                                               const unsigned short* __restrict__ lut, int b0, int nb) {
  if (nb <= 0 || !(x > 0.0)) return upper_index(xp, n, x);
  int b = (int)(__float_as_uint(__double2float_rd(x)) >> CHB_LUT_SHIFT) - b0;
  b = max(0, min(b, nb - 1));
  int k = lut[b];
  while (k < n - 2 && xp[k + 1] <= x) ++k;
  return k + 1;
}
__device__ __forceinline__ int upper_index_log(const double* __restrict__ xp, int n, double x, double log10_x,
                                               double log10_first, double inv_dlog10) {
  int k = (int)((log10_x - log10_first) * inv_dlog10);
  k = max(0, min(k, n - 2));
  while (k < n - 2 && xp[k + 1] <= x) ++k;
  while (k > 0 && xp[k] > x) --k;
  return k + 1;
}

// ------------------------------------------------------------------------------------------
// cosmology
__device__ __forceinline__ double E_at_z(const double* __restrict__ P, const double* __restrict__ HC, double z) {
  double x = 1.0 + z, x2 = x * x;
  double de = 1.0;
  if (HC[HC_DE_CONST] == 0.0) {
    double wz = P[CHB_P_W0] + P[CHB_P_WA] * z / (1.0 + z);
    de = pow(x, 3.0 * (1.0 + wz));
  }
  return sqrt(P[CHB_P_OM0] * (x2 * x) + P[CHB_P_OR0] * (x2 * x2) + P[CHB_P_OK0] * x2 + HC[HC_ODE0] * de);
}

// transverse comoving distance from the radial one (cosmo.py:142-153)
__device__ __forceinline__ double dCt_from_dCr(const double* __restrict__ P, const double* __restrict__ HC, double dCr) {
  double Ok0 = P[CHB_P_OK0];
  if (Ok0 == 0.0) return dCr;
  double dH = HC[HC_DH], s = HC[HC_SQRTOK];
  return (Ok0 > 0.0) ? (dH / s) * sinh(s * dCr / dH) : (dH / s) * sin(s * dCr / dH);
}

__device__ __forceinline__ double Xi_at_z(const double* __restrict__ P, double z) {
  return P[CHB_P_XI0] + (1.0 - P[CHB_P_XI0]) / pow_pos(1.0 + z, P[CHB_P_N]);
}

// cosmo.py:201-203 / 230-235
__device__ __forceinline__ double dL2dCt(int cosmo_model, const double* __restrict__ P, double dist, double z) {
  if (cosmo_model == CHB_COSMO_MG_FLRW) return (dist / Xi_at_z(P, z)) / (1.0 + z);
  return dist / (1.0 + z);
}

// ddL/dz given dCt and E (cosmo.py:212-221 / 245-257)
__device__ __forceinline__ double ddLdz_from(int cosmo_model, const double* __restrict__ P, const double* __restrict__ HC,
                                             double z, double dCt, double Ez) {
  double ddL = dCt + (HC[HC_DH] / Ez) * (1.0 + z);
  if (cosmo_model == CHB_COSMO_MG_FLRW) {
    double dLflrw = dCt * (1.0 + z);
    double n = P[CHB_P_N], Xi0 = P[CHB_P_XI0];
    double dXi = n * (Xi0 - 1.0) / pow_pos(1.0 + z, n + 1.0);
    return ddL * Xi_at_z(P, z) + dLflrw * dXi;
  }
  return ddL;
}

__device__ __forceinline__ double dVcdz_from(const double* __restrict__ HC, double dCt, double Ez) {
  return 4.0 * CHB_PI * HC[HC_DH] * dCt * dCt / Ez;
}

// cosmo.py:166-186
__device__ __forceinline__ double Vc_from_dCt(const double* __restrict__ P, const double* __restrict__ HC, double dCt) {
  double Ok0 = P[CHB_P_OK0];
  if (Ok0 == 0.0) return 4.0 * CHB_PI * dCt * dCt * dCt / 3.0;
  double reg = Ok0 + 1e-10, s = HC[HC_SQRTOK], dH = HC[HC_DH];
  double pref = 4.0 * CHB_PI * dH * dH * dH / (2.0 * reg);
  double t = (dCt / dH) * sqrt(1.0 + reg * dCt * dCt / (dH * dH));
  return (Ok0 > 0.0) ? pref * (t - asinh(s * dCt / dH) / s) : pref * (t - asin(s * dCt / dH) / s);
}

// ------------------------------------------------------------------------------------------
// mass
__device__ __forceinline__ double tpl_notnorm(double m, double alpha, double lo, double hi) {
  return (lo <= m && m <= hi) ? pow_pos(m, alpha) : 0.0;
}
// mass.py:255-264: exp(-logaddexp(0,t)) == 1/(1+e^t)
__device__ __forceinline__ double smoothing(double m, double dm, double lo) {
  if (m < lo) return 0.0;
  if (m > lo + dm) return 1.0;
  double t = dm / (m - lo + 1e-99) + dm / (m - lo - dm + 1e-99);
  return 1.0 / (1.0 + exp(t));
}
__device__ __forceinline__ double gaussian_pdf(double x, double mu, double sigma) {
  double d = x - mu;
  return exp(-0.9189385332046727 - log(sigma) - d * d / (2.0 * sigma * sigma));
}

// mass.py:285-305
__device__ __forceinline__ double primary_notnorm(int mass_model, const double* __restrict__ P,
                                                  const double* __restrict__ HC, double m) {
  double lo = P[CHB_P_MLOW], hi = P[CHB_P_MHIGH];
  if (mass_model == CHB_MASS_TPL) return tpl_notnorm(m, -P[CHB_P_ALPHA], lo, hi);
  if (mass_model == CHB_MASS_BPL) {
    double mb = HC[HC_MBREAK];
    double pdf = tpl_notnorm(m, -P[CHB_P_ALPHA], lo, mb);
    pdf += tpl_notnorm(m, -P[CHB_P_ALPHA2], mb, hi) * HC[HC_BPL_RATIO];
    return pdf * smoothing(m, P[CHB_P_DELTAM], lo);
  }
  // plp
  double Ppl = tpl_notnorm(m, -P[CHB_P_ALPHA], lo, hi) / HC[HC_PL_NORM];
  double mu = P[CHB_P_MUG], sg = P[CHB_P_SIGMAG];
  double G = (lo <= m && m <= mu + 5.0 * sg) ? gaussian_pdf(m, mu, sg) / HC[HC_TG_NORM] : 0.0;
  double lam = P[CHB_P_LAMBDAP];
  return ((1.0 - lam) * Ppl + lam * G) * smoothing(m, P[CHB_P_DELTAM], lo);
}

// mass.py:320-328
__device__ __forceinline__ double secondary_notnorm(int mass_model, const double* __restrict__ P, double m2, double m1) {
  double pdf = tpl_notnorm(m2, P[CHB_P_BETA], P[CHB_P_MLOW], m1);
  if (mass_model != CHB_MASS_TPL) pdf *= smoothing(m2, P[CHB_P_DELTAM], P[CHB_P_MLOW]);
  return pdf;
}

// mass.py:334-345.  mg/cdf: the hyper-point's m_grid and cdf_m2_conditioned tables (rm entries).
__device__ __forceinline__ double p_m1m2(int mass_model, const double* __restrict__ P, const double* __restrict__ HC,
                                         const double* __restrict__ mg, const double* __restrict__ cdf, int rm,
                                         double m1, double m2) {
  double p1 = primary_notnorm(mass_model, P, HC, m1) / HC[HC_NORM_P_M1];
  double p2 = secondary_notnorm(mass_model, P, m2, m1) / interp_clamped(m1, mg, cdf, rm);
  if (isnan(p2)) p2 = 0.0;
  return p1 * p2;
}

// same with the log-spaced index (no binary search over m_grid); bit-identical to p_m1m2
__device__ __forceinline__ double p_m1m2_logidx(int mass_model, const double* __restrict__ P, const double* __restrict__ HC,
                                                const double* __restrict__ mg, const double* __restrict__ cdf, int rm,
                                                double m1, double m2) {
  double p1 = primary_notnorm(mass_model, P, HC, m1) / HC[HC_NORM_P_M1];
  const int i = upper_index_log(mg, rm, m1, log10(m1), HC[HC_LOG10_MLOW], 1.0 / HC[HC_DLOG10_M]);
  double p2 = secondary_notnorm(mass_model, P, m2, m1) / interp_at(m1, mg, cdf, rm, i);
  if (isnan(p2)) p2 = 0.0;
  return p1 * p2;
}

// ------------------------------------------------------------------------------------------
// rate (rate.py:96-129)
__device__ __forceinline__ double merger_rate(int rate_model, const double* __restrict__ P,
                                              const double* __restrict__ HC, double z) {
  double x = 1.0 + z, g = P[CHB_P_GAMMA];
  if (rate_model == CHB_RATE_POWER_LAW) return pow_pos(x, g);
  if (rate_model == CHB_RATE_TRUNC_PL) return (z < P[CHB_P_RZMAX]) ? pow_pos(x, g) * HC[HC_RATE_NORM] : 0.0;
  double k = P[CHB_P_KAPPA], zp = P[CHB_P_ZP];
  double md = pow_pos(x, g) / (1.0 + pow_pos(x / (1.0 + zp), g + k));
  double val = HC[HC_RATE_NORM] * md;
  if (rate_model == CHB_RATE_TRUNC_MD) return (z < P[CHB_P_RZMAX]) ? val : 0.0;
  return val;
}

// ------------------------------------------------------------------------------------------
// per-hyper-point table block in global memory: [zg | iinv | dLt | mg | cdf], strides padded to
// even counts so every sub-table is 16-byte aligned (bulk-copy requirement).
//
// fp32 fast-path block (same hyper-point, appended): float4 rows so that one LDS.128 brings the
// knot, the value, the slope and the next knot:
//   zi4[k] = {z_k, integral_invE_k, slope_k, z_{k+1}}      (index from log2 z: the grid is log-spaced)
//   dl4[k] = {dL_k, z_k, dz/ddL slope_k, dL_{k+1}}         (index from a float-bits LUT + short scan)
//   cd4[k] = {m_k, cdf_k, slope_k, m_{k+1}}                (index from log2 m)
//   lut[b] = first candidate interval for dL in float-bits bucket b (uint16)
struct TableLayout {
  int rc, rm;        // table resolutions
  int rcs, rms;      // padded strides
  __host__ __device__ int f64_total() const { return 3 * rcs + 2 * rms; }
  __host__ __device__ int off_zg() const { return 0; }
  __host__ __device__ int off_iinv() const { return rcs; }
  __host__ __device__ int off_dLt() const { return 2 * rcs; }
  __host__ __device__ int off_mg() const { return 3 * rcs; }
  __host__ __device__ int off_cdf() const { return 3 * rcs + rms; }
  // fp32 block, offsets in doubles relative to the start of the hyper-point's table block
  __host__ __device__ int off_f32() const { return f64_total(); }
  __host__ __device__ int f32_zi4() const { return 0; }
  __host__ __device__ int f32_dl4() const { return 2 * rcs; }
  __host__ __device__ int f32_cd4() const { return 4 * rcs; }
  __host__ __device__ int f32_lut() const { return 4 * rcs + 2 * rms; }
  __host__ __device__ int f32_fc() const { return 4 * rcs + 2 * rms + CHB_LUT_CAP / 4; }   // CHB_NFC floats (FC_* below)
  // br: one float4 row {dL_k, z_k, slope_k, dL_k+1} PER LUT BUCKET (the dl4 row of the knot interval at the bucket's lower edge):
  // z_from_dGW reads br[b] and br[b + 1] with two independent loads instead of the dependent chain lut[b] -> dl4[k] -> dl4[k + 1]
  __host__ __device__ int f32_total() const { return 4 * rcs + 2 * rms + CHB_LUT_CAP / 4 + CHB_NFC / 2; }
  __host__ __device__ int f32_core() const { return f32_total(); }    // what the staging kernels copy
  __host__ __device__ int total() const { return f64_total() + f32_total(); }
};
static inline TableLayout make_layout(int rc, int rm) {
  TableLayout t;
  t.rc = rc; t.rm = rm;
  t.rcs = (rc + 1) & ~1; t.rms = (rm + 1) & ~1;
  return t;
}

// ------------------------------------------------------------------------------------------
// block-wide reductions (deterministic for a fixed block size)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// red: >= 32 doubles of shared memory. Result broadcast to all threads.
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  return warp_sum(r);
}
__device__ __forceinline__ double block_min(double v, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : INFINITY;
  return warp_min(r);
}
__device__ __forceinline__ double block_max(double v, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : -INFINITY;
  return warp_max(r);
}
