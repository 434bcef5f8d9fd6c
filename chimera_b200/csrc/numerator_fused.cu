// numerator_fused.cu -- ONE kernel per step for the 1-D kinds of the fp32 mode (non-pixelated and 'approximate'):
// population reweighting -> statistics -> [binning] -> KDE -> interpolation -> z-integral -> log, per (event, hyper-point)
// unit, with the reweighted samples living only in shared memory (likelihood.py:105-154, 266-301; pop_wrapper.py:67-80;
// utils/math.py:32-89).
//
// Why it replaces the split MODE 1 -> stage buffer -> MODE 2 pair of numerator_f32.cu (round 1):
//   * no stage buffer: the split form wrote and re-read 8 B per (sample, hyper-point) through HBM (2 x 10.2 GB per C3 step);
//   * no 42 KB table block next to the 40 KB sample stage.  The packed tables (dl4 | cd4 | lut, models.cuh) are read
//     with ld.global.nc through L1: the samples of an event are sorted by dL at upload, so a warp's 64 consecutive
//     samples fall into one or two rows of dl4, and an event's source-frame masses into a few dozen rows of cd4 -- the
//     working set per CTA is a few KB.  53 KB of shared memory per CTA -> 3-4 co-resident CTAs per SM;
//   * no rescaling pass: the stage keeps {z - z0, log2 w} pairs, the scale 1/bw rides in the FFMA2 that forms x' - g
//     (kde_win.cuh, RAW), 1/sum(w) in the exponent offset; the per-chunk summaries the windows need are taken while the
//     samples are still in registers (one redux per 64 samples);
//   * bandwidth and tiling are derived by every thread from the block-reduced statistics (no one-thread section and
//     no barrier behind it); the per-hyper-point single-precision constants come from the FC row build_tables_kernel
//     wrote (no fp64 divisions per thread per unit);
//   * binning (utils/math.py:32-46) runs on the shared-memory stage by contiguous runs of the sorted samples.
// fp64 is kept for: tables, block-level statistics, bandwidth, grid geometry, cross-warp accumulation, interpolation,
// z-integral and the final reduction -- as in numerator_f32.cu.
#include "common.cuh"
#include "models_f32.cuh"
#include "kde_f32.cuh"
#include "kde_win.cuh"
#include <algorithm>

#ifndef FU_NT
#define FU_NT 256
#endif
// This file is compiled three times: as is (256 threads per CTA) and through numerator_fused_nt128.cu / _nt64.cu, which
// define FU_NT and CHB_FU_VARIANT -- further instantiations of the 1-D fused kernel for events with few samples (walker
// batches with ~1000 samples per event), where the per-unit work every thread repeats (statistics, bandwidth, window plan)
// and the barriers weigh as much as the sums: fewer warps per unit, more units in flight.  A variant renames the
// kernel and its five host entry points and leaves out the 'marginalized' kernel.
#ifdef CHB_FU_VARIANT
#define FU_CAT2(a, b) a##b
#define FU_CAT(a, b) FU_CAT2(a, b)
#define FU_NAME(x) FU_CAT(x, CHB_FU_VARIANT)
#else
#define FU_NAME(x) x
#endif
#define numerator_fused_kernel FU_NAME(numerator_fused_kernel)
#define numerator_fused_smem_bytes FU_NAME(numerator_fused_smem_bytes)
#define numerator_fused_supported FU_NAME(numerator_fused_supported)
#define numerator_fused_configure FU_NAME(numerator_fused_configure)
#define numerator_fused_ctas_per_sm FU_NAME(numerator_fused_ctas_per_sm)
#define launch_numerator_fused FU_NAME(launch_numerator_fused)
#define FU_NW (FU_NT / 32)
#define FU_SUB 64                 // samples per warp iteration of the reweighting loop = granularity of the chunk summaries
#define FU_INB (FU_SUB * 24)      // bytes of one block of packed samples: 64 x (float4 s4 + float2 l2)
#ifndef CHB_FU_MINB
#define CHB_FU_MINB 3             // co-resident CTAs per SM the kernel is compiled for
#endif
#ifndef CHB_FU_L1PF
#define CHB_FU_L1PF 0             // 1: at the start of a unit every thread issues prefetch.global.L1 for the table rows the
#endif                            //    unit is about to touch (a hint: the rows' first uses then hit L1 instead of waiting on L2)
#ifndef CHB_FU_CPASYNC
#define CHB_FU_CPASYNC 1          // 1: every warp stages its NEXT 64-sample block of packed samples in shared memory with
#endif                            //    cp.async (1.5 KB per warp, issued a whole block ahead: the L2 latency of the sample stream is
                                  //    off the critical path and no second register set is needed); 0: direct ld.global.cg
#ifndef CHB_FU_TAILPF
#define CHB_FU_TAILPF 0           // 1: right after the reweighting every 16th thread issues prefetch.global.L1 for the rows the
#endif                            //    z-integral reads once the KDE is done (event grid, z-grid terms, collapsed catalogue rows) and
                                  //    for the next unit's constants: their L2 latency hides behind the KDE instead of standing
                                  //    at the end of the unit, where two warps walk the 300-point grid alone.  Measured: 0.6 % SLOWER
                                  //    (the co-resident CTAs already cover that latency); kept as a switch
#ifndef CHB_FU_PREFETCH
#define CHB_FU_PREFETCH 0         // 1: the next block's packed samples are requested one block ahead (two register sets);
#endif                            // 0 (default, measured 1 % faster): every block requests its successor's samples right after
                                  //    its own evaluation (12 fewer live registers, the L2 latency is covered by the other warps)

struct FusedPlan {                // byte offsets into dynamic shared memory
  int stage, rows, dens, bc, bs, bx, sub, summ, win, cr, red, total;
};
static __host__ __device__ inline FusedPlan make_fused_plan(int Ns, int Nz, int B) {
  FusedPlan p;
  const int G = Nz / 2;
  const int Bp = (B + 1) & ~1;
  int o = 0;
  p.stage = o; o += Ns * 8;                       // float4 per two samples {dz_a, dz_b, v_a, v_b}
  p.rows = o; o += max(FU_NW * G * 8, FU_NW * FU_INB);   // per-warp partial rows (doubles); during the reweighting: the
                                                         // cp.async landing zone of the packed samples (FU_INB bytes per warp)
  p.dens = o; o += ((G + 1) & ~1) * 8;
  p.bc = o; o += 3 * (B + 2) * 8;                 // inclusive prefix sums S0 | S1 | S2 over the bins (Epanechnikov)
  p.bs = o; o += Bp * 8;
  p.bx = o; o += Bp * 8;                          // binned data set in pair layout
  p.sub = o; o += ((Ns + FU_SUB - 1) / FU_SUB) * 16;
  p.summ = o; o += 32 * 16;
  p.win = o; o += 32 * 8;
  p.cr = o; o += 16 * 4;
  p.red = o; o += 96 * 8;
  p.total = o;
  return p;
}
size_t numerator_fused_smem_bytes(const NumArgs& a) {
  return (size_t)make_fused_plan(a.Ns, a.Nz, a.binning ? a.num_bins : 0).total;
}
bool numerator_fused_supported(const NumArgs& a) {
  return a.fp_mode == CHB_FP32 && (a.kind == CHB_PGW_1D || a.kind == CHB_PGW_APPROX) && a.use_cut && (a.Ns % 2 == 0) &&
         a.Nz / 2 >= 2 && a.s4 != nullptr;
}

__device__ __forceinline__ double nan_to_num_log_fu(double like) {
  double l = log(like);                  // likelihood.py:296-297
  if (isnan(l)) return -INFINITY;
  if (isinf(l)) return l > 0 ? CHB_DBL_MAX : -CHB_DBL_MAX;
  return l;
}

struct FuStats { double a, b, c, d; float mn, mx; };
__device__ __forceinline__ FuStats fu_block_stats(FuStats v, double* red /* >= 6*FU_NW doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.a += __shfl_xor_sync(0xffffffffu, v.a, o);
    v.b += __shfl_xor_sync(0xffffffffu, v.b, o);
    v.c += __shfl_xor_sync(0xffffffffu, v.c, o);
    v.d += __shfl_xor_sync(0xffffffffu, v.d, o);
  }
  v.mn = warp_min_f32(v.mn); v.mx = warp_max_f32(v.mx);
  if (lane == 0) {
    red[w * 6 + 0] = v.a; red[w * 6 + 1] = v.b; red[w * 6 + 2] = v.c; red[w * 6 + 3] = v.d;
    red[w * 6 + 4] = (double)v.mn; red[w * 6 + 5] = (double)v.mx;
  }
  __syncthreads();
  FuStats r = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < FU_NW; ++i) {
    r.a += red[i * 6 + 0]; r.b += red[i * 6 + 1]; r.c += red[i * 6 + 2]; r.d += red[i * 6 + 3];
    r.mn = fminf(r.mn, (float)red[i * 6 + 4]); r.mx = fmaxf(r.mx, (float)red[i * 6 + 5]);
  }
  return r;
}

// Range [k0, k1) of the event's z grid inside [lb, ub] (whole-warp call, every warp for itself).  p_gw is zero outside the
// effective grid (likelihood.py:139-141 interpolates with left = right = 0), so the z-integral only has to visit these
// points: an event grid laid out for the whole H0 prior is two to three times wider than the samples' support at one
// hyper-point.  The grid is ascending in the reference (np.linspace / logspace); if the count of inside points does not
// match the range (a grid that is not ascending), the whole grid is returned.
__device__ __forceinline__ void fu_grid_range(const double* __restrict__ zgr, int Nz, double lb, double ub, int& k0, int& k1) {
  const int lane = threadIdx.x & 31;
  int below = 0, upto = 0, inside = 0;
  for (int kb = 0; kb < Nz; kb += 32) {
    const int k = kb + lane;
    const double x = (k < Nz) ? zgr[k] : INFINITY;
    below += __popc(__ballot_sync(0xffffffffu, x < lb));
    upto += __popc(__ballot_sync(0xffffffffu, x <= ub));
    inside += __popc(__ballot_sync(0xffffffffu, x >= lb && x <= ub));
  }
  k0 = below; k1 = upto;
  if (k1 - k0 != inside || k1 < k0 || !(lb <= ub)) { k0 = 0; k1 = Nz; }      // (NaN bounds: every point, so that NaN propagates as before)
}

// ---- population reweighting in single precision (pop_wrapper.py:67-80) ---------------------------------------
// per-hyper-point view of the packed tables (global memory, read through L1) and the FC constants
struct FuTab {
  const float4* __restrict__ dl4; const float4* __restrict__ cd4; const unsigned short* __restrict__ lut;
  int rc, rm, nb, b0;
  float z_top, lg2_m0, inv_lg2_mstep, lo, hi, dm, neg_alpha, beta, kA, kG, mu, g_hi, g_c, mb, neg_alpha2, cdl_x, cdl_y;
};
__device__ __forceinline__ FuTab make_fu_tab(const TableLayout& lay, const double* __restrict__ f32blk, const float* __restrict__ FC) {
  FuTab t;
  t.dl4 = reinterpret_cast<const float4*>(f32blk + lay.f32_dl4());
  t.cd4 = reinterpret_cast<const float4*>(f32blk + lay.f32_cd4());
  t.lut = reinterpret_cast<const unsigned short*>(f32blk + lay.f32_lut());
  t.rc = lay.rc; t.rm = lay.rm;
  t.b0 = __float_as_int(FC[FC_LUT_B0]); t.nb = __float_as_int(FC[FC_LUT_NB]);
  t.z_top = FC[FC_Z_TOP]; t.lg2_m0 = FC[FC_LG2_M0]; t.inv_lg2_mstep = FC[FC_INV_LG2_MSTEP];
  t.lo = FC[FC_LO]; t.hi = FC[FC_HI]; t.dm = FC[FC_DM]; t.neg_alpha = FC[FC_NEG_ALPHA]; t.beta = FC[FC_BETA];
  t.kA = FC[FC_KA]; t.kG = FC[FC_KG]; t.mu = FC[FC_MU]; t.g_hi = FC[FC_G_HI]; t.g_c = FC[FC_G_C];
  t.mb = FC[FC_MB]; t.neg_alpha2 = FC[FC_NEG_ALPHA2]; t.cdl_x = FC[FC_CDL_X]; t.cdl_y = FC[FC_CDL_Y];
  return t;
}

// z_from_dGW (cosmo.py:260-264): float-bits LUT -> candidate row, one scan step (64 buckets per octave of dL hold at
// most one knot of the 140-per-decade table), rare longer scans in a loop; clamped ends like numpy.interp.
__device__ __forceinline__ void fu_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ int fu_lut_bucket(const FuTab& t, float dL) {
  const int b = (int)(__float_as_uint(dL) >> CHB_LUT_SHIFT) - t.b0;
  return max(0, min(b, t.nb - 1));
}
__device__ __forceinline__ float fu_z_from_dL(const FuTab& t, float dL) {
  const int b = fu_lut_bucket(t, dL);
  // (a per-bucket copy of the rows -- two independent loads instead of lut -> dl4 -> dl4 -- was measured: no gain, the
  //  other warps of the SM cover this chain)
  int k = __ldg(t.lut + b);
  float4 e = __ldg(t.dl4 + k);
  if (dL >= e.w && k < t.rc - 2) { ++k; e = __ldg(t.dl4 + k); }
  while (dL >= e.w && k < t.rc - 2) { ++k; e = __ldg(t.dl4 + k); }
  float z = fmaf(dL - e.x, e.z, e.y);
  z = (dL >= e.w) ? t.z_top : z;
  return (dL <= 0.f) ? 0.f : z;
}

// mass.py:255-264 with one reciprocal: dm/x + dm/(x-dm) = dm (2x - dm) / (x (x - dm))
__device__ __forceinline__ float fu_smoothing(float m, float dm, float lo) {
  const float x = m - lo;
  const float tt = dm * (2.f * x - dm) * rcpf_(x * (x - dm));
  float S = rcpf_(1.f + ex2f_(tt * CHB_LOG2E_F));
  S = (x > dm) ? 1.f : S;
  return (x < 0.f) ? 0.f : S;
}

// p_m1m2 / pe_prior (mass.py:334-345, pop_wrapper.py:79); lm1/lm2 = log2 of the source-frame masses.
// SMOOTH = false: the caller has checked that every mass of the warp is above m_low + delta_m (taper == 1 exactly).
template <int MASS, bool SMOOTH>
__device__ __forceinline__ float fu_weight(const FuTab& t, float m1, float m2, float lm1, float lm2, float inv_prior) {
  float p1;                                                  // primary pdf, already divided by norm_p_m1
  if (MASS == CHB_MASS_TPL) {
    p1 = ex2f_(fmaf(t.neg_alpha, lm1, t.kA));
  } else if (MASS == CHB_MASS_BPL) {
    const float pa = ex2f_(fmaf(t.neg_alpha, lm1, t.kA)), pb = ex2f_(fmaf(t.neg_alpha2, lm1, t.kG));
    p1 = ((m1 <= t.mb) ? pa : 0.f) + ((m1 >= t.mb) ? pb : 0.f);
  } else {
    const float d = m1 - t.mu;
    const float g = ex2f_(fmaf(t.g_c * d, d, t.kG));
    p1 = ex2f_(fmaf(t.neg_alpha, lm1, t.kA)) + ((m1 <= t.g_hi) ? g : 0.f);
  }
  float p2 = ex2f_(t.beta * lm2);
  if (SMOOTH && MASS != CHB_MASS_TPL) { p1 *= fu_smoothing(m1, t.dm, t.lo); p2 *= fu_smoothing(m2, t.dm, t.lo); }
  int i = (int)((lm1 - t.lg2_m0) * t.inv_lg2_mstep);
  i = max(0, min(i, t.rm - 2));
  // (an index that is one off at a knot extrapolates the neighbouring segment by an ulp: same value to rounding)
  const float4 e = __ldg(t.cd4 + i);
  const float cdf = fmaf(m1 - e.x, e.z, e.y);                // (m1 <= m_high = last knot inside the support: no clamp needed)
  p2 = p2 * rcpf_(cdf);
  p2 = (p2 != p2) ? 0.f : p2;                                // 0/0 -> 0 (mass.py:340)
  const bool in = (m1 <= t.hi) && (t.lo <= m2) && (m2 <= m1) && (p1 > 0.f);   // (lo <= m2 <= m1 implies lo <= m1); p1 == 0: never 0 * inf
  return in ? p1 * p2 * inv_prior : 0.f;
}

__device__ __forceinline__ void fu_cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fu_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fu_cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One pass over the event's packed samples: {z - z0, v} pairs to the shared-memory stage (v = log2 w when `want_lw`,
// else w), per-warp statistics {sum w, sum w^2, sum dz, sum dz^2, min dz, max dz} to red[warp*6..] (+ z0 in red[95]), and
// per 64-sample block {min dz, max dz, max log2 w, dz of that sample} over the samples with w > 0.  Static round-robin
// of blocks over the warps: bit-reproducible.  NOT inlined: the loop gets the kernel's whole register budget to itself
// (the unit-level state of the caller is saved around the call once per unit instead of spilling inside the loop).
template <int MASS>
static __device__ __noinline__ void fu_reweight(const double* __restrict__ f32blk, int rc, int rcs, int rms, int rm,
                                         const float* __restrict__ FC, const float4* __restrict__ s4,
                                         const float2* __restrict__ l2, int Ns, int want_lw_i, float4* __restrict__ stage,
                                         float4* __restrict__ sub, double* __restrict__ red, float4* __restrict__ inb) {
  TableLayout lay; lay.rc = rc; lay.rm = rm; lay.rcs = rcs; lay.rms = rms;
  const FuTab t = make_fu_tab(lay, f32blk, FC);
  const bool want_lw = want_lw_i != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // z shifted by the redshift of the event's median-dL sample: {z - z0} is exact in fp32 and small, so the one-pass
  // variance does not cancel and the scaled coordinates of the pair sums keep ~1e-6 absolute accuracy
  const float4 smed = __ldg(&s4[Ns / 2]);
  const float z0 = fu_z_from_dL(t, smed.x);
#if CHB_FU_L1PF
  {
    // L1 prefetch of the table rows this unit is about to touch (hints only; 128-byte lines = 8 float4 rows / 64 LUT
    // entries): the LUT entries and dl4 rows between the smallest and the largest dL of the event (its first and last
    // sample when sorted; min/max otherwise cost nothing but coverage), and the cd4 rows within +-30 % of the median
    // sample's source-frame primary mass.  Without it the first use of every row waits on L2 inside the dependency
    // chain of a block (2-3 new dl4 rows per 64 sorted samples).
    const float dlo = fminf(__ldg(&s4[0].x), __ldg(&s4[Ns - 1].x)), dhi = fmaxf(__ldg(&s4[0].x), __ldg(&s4[Ns - 1].x));
    const int blo = fu_lut_bucket(t, dlo), bhi = fu_lut_bucket(t, dhi);
    const int tid = threadIdx.x;
    if (tid * 64 <= bhi - blo + 64) fu_prefetch_l1(t.lut + min(blo + tid * 64, t.nb - 1));
    const int klo = __ldg(t.lut + blo), khi = min((int)__ldg(t.lut + bhi) + 3, t.rc - 1);
    if (tid * 8 <= khi - klo + 8) fu_prefetch_l1(t.dl4 + min(klo + tid * 8, t.rc - 1));
    const float m1s = smed.y * rcpf_(1.f + z0);
    const int imed = (int)((lg2f_(m1s) - t.lg2_m0) * t.inv_lg2_mstep);
    const int ilo = max(0, min(imed - 128, t.rm - 1)), ihi = max(0, min(imed + 128, t.rm - 1));
    if (tid * 8 <= ihi - ilo + 8) fu_prefetch_l1(t.cd4 + min(ilo + tid * 8, t.rm - 1));
  }
#endif
  float fa = 0.f, fb = 0.f, fcs = 0.f, fd = 0.f, mn = INFINITY, mx = -INFINITY;
  // One 64-sample block: `sa, sb, ll` hold the block's packed samples (requested one block earlier), the NEXT block
  // of this warp is requested into `na, nb, nl` before this one is evaluated (L2 latency behind ~300 instructions).
  // The loop below alternates two register sets, so no block is ever copied between registers.
  auto block = [&](int cb, const float4& sa, const float4& sb, const float4& ll, float4& na, float4& nb, float4& nl) {
    const int nbase = cb + FU_NW * FU_SUB;
    if (CHB_FU_PREFETCH && nbase < Ns) {
      // (streamed once per unit: ld.global.cg keeps them out of L1, which then holds only the table rows in use)
      const int j = min(nbase + 2 * lane, Ns - 2);             // tail lanes recompute the last pair, never stored
      na = __ldcg(s4 + j); nb = __ldcg(s4 + j + 1); nl = __ldcg(reinterpret_cast<const float4*>(l2 + j));
    }
    const bool valid = cb + 2 * lane < Ns;
    const float za = fu_z_from_dL(t, sa.x), zb = fu_z_from_dL(t, sb.x);
    const float opa = 1.f + za, opb = 1.f + zb;
    const float ra = rcpf_(opa), rb = rcpf_(opb), lza = lg2f_(opa), lzb = lg2f_(opb);
    const float m1a = sa.y * ra, m2a = sa.z * ra, m1b = sb.y * rb, m2b = sb.z * rb;
    float wa, wb;
    // (m2 <= m1 inside the support, so the taper test on m2 covers m1 whenever the weight is not already zero)
    const bool taper = (MASS != CHB_MASS_TPL) && (!(m2a - t.lo > t.dm) || !(m2b - t.lo > t.dm));
    if (MASS != CHB_MASS_TPL && __any_sync(0xffffffffu, taper)) {
      wa = fu_weight<MASS, true>(t, m1a, m2a, ll.x - lza, ll.y - lza, sa.w);
      wb = fu_weight<MASS, true>(t, m1b, m2b, ll.z - lzb, ll.w - lzb, sb.w);
    } else {
      wa = fu_weight<MASS, false>(t, m1a, m2a, ll.x - lza, ll.y - lza, sa.w);
      wb = fu_weight<MASS, false>(t, m1b, m2b, ll.z - lzb, ll.w - lzb, sb.w);
    }
    const float dza = za - z0, dzb = zb - z0;
    if (valid) {
      fa += wa + wb; fb = fmaf(wa, wa, fmaf(wb, wb, fb)); fcs += dza + dzb; fd = fmaf(dza, dza, fmaf(dzb, dzb, fd));
      mn = fminf(mn, fminf(dza, dzb)); mx = fmaxf(mx, fmaxf(dza, dzb));
    }
    if (want_lw) {
      // log2 w: lg2(0) = -inf natively (a zero weight adds exactly 0); NaN weights are forced to -inf as well
      float lwa = lg2f_(wa), lwb = lg2f_(wb);
      lwa = (valid && wa > 0.f) ? lwa : -INFINITY; lwb = (valid && wb > 0.f) ? lwb : -INFINITY;
      if (valid) stage[(cb >> 1) + lane] = make_float4(dza, dzb, lwa, lwb);
      // block summary over ALL its samples (a superset of the weighted ones: still a valid hull for the windows)
      const float lo = warp_min_f32(valid ? fminf(dza, dzb) : INFINITY);
      const float hi = warp_max_f32(valid ? fmaxf(dza, dzb) : -INFINITY);
      const float lm = fmaxf(lwa, lwb), xm = (lwa >= lwb) ? dza : dzb;
      const float lmw = warp_max_f32(lm);
      const unsigned pick = __ballot_sync(0xffffffffu, lm == lmw);
      const float xmw = __shfl_sync(0xffffffffu, xm, __ffs(pick) - 1);
      if (lane == 0) sub[cb / FU_SUB] = make_float4(lo, hi, lmw, xmw);
    } else if (valid) {
      stage[(cb >> 1) + lane] = make_float4(dza, dzb, wa, wb);
    }
  };
  int cb = warp * FU_SUB;
  float4 a0, a1, a2, b0, b1, b2;
#if CHB_FU_CPASYNC
  // `inb`: this warp's landing zone, 96 float4 = [lane] s4 of the lane's first sample | [32 + lane] second sample |
  // [64 + lane] the two {log2 m1det, log2 m2det} pairs.  Every lane copies and reads ONLY its own three 16-byte pieces, so
  // cp.async.wait_group orders everything (no warp barrier); the block after next is requested as soon as the current
  // one sits in registers, a whole block of arithmetic before it is needed.
  auto request = [&](int c) {
    const int j = min(c + 2 * lane, Ns - 2);                   // tail lanes recompute the last pair, never stored
    fu_cp_async16(inb + lane, s4 + j); fu_cp_async16(inb + 32 + lane, s4 + j + 1);
    fu_cp_async16(inb + 64 + lane, reinterpret_cast<const float4*>(l2 + j));
    fu_cp_async_commit();
  };
  if (cb < Ns) request(cb);
  while (cb < Ns) {
    fu_cp_async_wait();
    a0 = inb[lane]; a1 = inb[32 + lane]; a2 = inb[64 + lane];
    const int nb_ = cb + FU_NW * FU_SUB;
    if (nb_ < Ns) request(nb_);
    block(cb, a0, a1, a2, b0, b1, b2);
    cb = nb_;
  }
#else
  if (cb < Ns) {
    const int j = min(cb + 2 * lane, Ns - 2);
    a0 = __ldcg(s4 + j); a1 = __ldcg(s4 + j + 1); a2 = __ldcg(reinterpret_cast<const float4*>(l2 + j));
  }
#if CHB_FU_PREFETCH
  while (cb < Ns) {
    block(cb, a0, a1, a2, b0, b1, b2);
    cb += FU_NW * FU_SUB;
    if (cb >= Ns) break;
    block(cb, b0, b1, b2, a0, a1, a2);
    cb += FU_NW * FU_SUB;
  }
#else
  while (cb < Ns) {
    block(cb, a0, a1, a2, b0, b1, b2);
    cb += FU_NW * FU_SUB;
    if (cb < Ns) {
      const int j = min(cb + 2 * lane, Ns - 2);
      a0 = __ldcg(s4 + j); a1 = __ldcg(s4 + j + 1); a2 = __ldcg(reinterpret_cast<const float4*>(l2 + j));
    }
  }
#endif
#endif
  // warp-level reduction; fp64 from here on
  double da = (double)fa, db = (double)fb, dc = (double)fcs, dd = (double)fd;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    da += __shfl_xor_sync(0xffffffffu, da, o);
    db += __shfl_xor_sync(0xffffffffu, db, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
    dd += __shfl_xor_sync(0xffffffffu, dd, o);
  }
  mn = warp_min_f32(mn); mx = warp_max_f32(mx);
  if (lane == 0) {
    red[warp * 6 + 0] = da; red[warp * 6 + 1] = db; red[warp * 6 + 2] = dc; red[warp * 6 + 3] = dd;
    red[warp * 6 + 4] = (double)mn; red[warp * 6 + 5] = (double)mx;
    if (warp == 0) red[95] = (double)z0;
  }
}

// Direct pair sums for data sets without a window plan (few samples, coarse grids, the 200 bins of the reference's
// default options, Epanechnikov).  pairs[i] = {x_a, x_b, v_a, v_b}; x' = x * sf; grid point g sits at gfirst + g hd
// (scaled units); Gaussian: v = log2 w', term 2^(v + koff - d^2); Epanechnikov: v = w, term v max(1 - d^2, 0).
// Lane owns R grid points of a 32R-point pass; the warps split the pairs; rows[warp][g] (doubles) receive the sums.
template <int R, bool GAUSS>
__device__ __forceinline__ void fu_direct_pass(const float4* __restrict__ pairs, int npairs, int G, int g_base, double gfirst,
                                               double hd, float sf, float koff, double* __restrict__ rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gp[R], acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int g = g_base + r * 32 + lane;
    gp[r] = (g < G) ? (float)(gfirst + (double)g * hd) : 3.0e18f;
    acc[r] = 0.f;
  }
  const int per = (npairs + FU_NW - 1) / FU_NW;
  const int j0 = min(npairs, warp * per), j1 = min(npairs, j0 + per);
#pragma unroll 2
  for (int j = j0; j < j1; ++j) {
    const float4 v = pairs[j];
    const float xa = v.x * sf, xb = v.y * sf;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float da = gp[r] - xa, db = gp[r] - xb;
      if (GAUSS) {
        acc[r] += ex2_ftz(fmaf(-da, da, v.z + koff)) + ex2_ftz(fmaf(-db, db, v.w + koff));
      } else {
        acc[r] = fmaf(v.z, fmaxf(fmaf(-da, da, 1.f), 0.f), acc[r]);
        acc[r] = fmaf(v.w, fmaxf(fmaf(-db, db, 1.f), 0.f), acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int g = g_base + r * 32 + lane;
    if (g < G) rows[warp * G + g] = (double)acc[r];
  }
}
// dens[g] = scale * sum over the data set; whole-CTA call; ends with dens[] written (no barrier after)
static __device__ __noinline__ void fu_direct(const float4* __restrict__ pairs, int npairs, int G, double gfirst, double hd,
                                          float sf, float koff, bool gauss, double scale, double* __restrict__ rows,
                                          double* __restrict__ dens) {
  if (G <= 160) {
    if (gauss) fu_direct_pass<5, true>(pairs, npairs, G, 0, gfirst, hd, sf, koff, rows);
    else fu_direct_pass<5, false>(pairs, npairs, G, 0, gfirst, hd, sf, koff, rows);
  } else {
    for (int gb = 0; gb < G; gb += 256) {
      if (gauss) fu_direct_pass<8, true>(pairs, npairs, G, gb, gfirst, hd, sf, koff, rows);
      else fu_direct_pass<8, false>(pairs, npairs, G, gb, gfirst, hd, sf, koff, rows);
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += FU_NT) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < FU_NW; ++w) acc += rows[w * G + g];
    dens[g] = acc * scale;
  }
}

// ---- Epanechnikov KDE of the staged samples WITHOUT binning (utils/math.py:52-85, kernel 'epan') ---------------------------
// The kernel has compact support, |g - x| <= 1 in units of the bandwidth, and the samples are sorted, so for a grid point
// almost every 64-sample block lies entirely inside or entirely outside the support:
//     inside:   sum_j w_j (1 - (g - x_j)^2) = (1 - D^2) M0 + 2 D M1 - M2,   D = g - c_b,  M_k = sum_j w_j (x_j - c_b)^k
//               about the block's own centre c_b (|D| <= 1 and a narrow block: nothing cancels in fp32);
//     outside:  skipped;      straddling the edge of the support (two blocks per grid point, more in sparse tails): summed directly.
// O(Ns + G Ns / 64) instead of the O(G Ns) pair sums.  The classification uses the block's hull {min x, max x}, so the
// result does not depend on the samples actually being sorted -- only the cost does.
// blk: 8 floats per block {lo, hi, c, M0, M1, M2, -, -} (the per-warp rows are idle in this path).  Whole-CTA call; ends
// with dens[] written (no barrier after).
static __device__ __noinline__ void fu_epan_blocks(const float4* __restrict__ pairs, int npairs, int G, double gfirst, double hd,
                                            float sf, double scale, float* __restrict__ blk, double* __restrict__ dens) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (npairs + 31) / 32;
  for (int b = warp; b < nblk; b += FU_NW) {
    const int j = b * 32 + lane;
    const bool on = j < npairs;
    const float4 v = on ? pairs[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float xa = v.x * sf, xb = v.y * sf;
    const float lo = warp_min_f32(on ? fminf(xa, xb) : INFINITY), hi = warp_max_f32(on ? fmaxf(xa, xb) : -INFINITY);
    const float c = 0.5f * (lo + hi);
    const float da = xa - c, db = xb - c;
    float m0 = v.z + v.w, m1 = fmaf(v.z, da, v.w * db), m2 = fmaf(v.z * da, da, v.w * db * db);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m0 += __shfl_xor_sync(0xffffffffu, m0, o); m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    if (lane == 0) {
      reinterpret_cast<float4*>(blk)[2 * b] = make_float4(lo, hi, c, m0);
      reinterpret_cast<float4*>(blk)[2 * b + 1] = make_float4(m1, m2, 0.f, 0.f);
    }
  }
  __syncthreads();
  // One grid point per warp and pass: the lanes classify the blocks (32 per round); a block inside the support adds its
  // moment form on the lane that owns it, the blocks straddling the support's edge are then summed by the WHOLE warp
  // (one pair per lane), and one shuffle reduction closes the grid point -- no lane ever walks a block alone.
  const float4* __restrict__ blk4 = reinterpret_cast<const float4*>(blk);
  for (int g = warp; g < G; g += FU_NW) {
    const float gp = (float)(gfirst + (double)g * hd);
    const float glo = gp - 1.f, ghi = gp + 1.f;
    float acc = 0.f;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      const int b = b0 + lane;
      bool part = false;
      if (b < nblk) {
        const float4 h4 = blk4[2 * b];
        const bool outside = (h4.y < glo) || (h4.x > ghi);
        const bool inside = (h4.x >= glo) && (h4.y <= ghi);
        if (inside) {
          const float4 m4 = blk4[2 * b + 1];
          const float D = gp - h4.z;
          acc += fmaf(fmaf(-D, D, 1.f), h4.w, fmaf(2.f * D, m4.x, -m4.y));
        }
        part = !outside && !inside;                            // (a NaN hull fails both tests and is summed directly)
      }
      unsigned m = __ballot_sync(0xffffffffu, part);
      while (m) {
        const int j = (b0 + __ffs(m) - 1) * 32 + lane;
        m &= m - 1;
        if (j < npairs) {
          const float4 v = pairs[j];
          const float da = gp - v.x * sf, db = gp - v.y * sf;
          acc = fmaf(v.z, fmaxf(fmaf(-da, da, 1.f), 0.f), acc);
          acc = fmaf(v.w, fmaxf(fmaf(-db, db, 1.f), 0.f), acc);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dens[g] = (double)acc * scale;
  }
}

// ---- Epanechnikov KDE of B equally spaced bin centres by prefix sums (utils/math.py:32-46 + 52-85) -------------------
// kde1d with the compact kernel K(u) = 3/4 (1 - u^2) [|u| <= 1] over bin centres c_b = first + b * bstep with weights
// w_b: the bins inside the support of a grid point form a contiguous index range, so
//     sum_b w'_b (1 - (gu - u_b)^2) = (1 - gu^2) dS0 + 2 gu dS1 - dS2,   u = (c - cmid) / bw,
// with dS* differences of inclusive prefix sums of {w', w' u, w' u^2} (fp64; coordinates centred on the bin range and
// scaled by 1/bw so the cancellation costs ~2 digits of 16).  O(B) to build (one warp), O(1) per grid point.
struct EpanBins { double zlo, zhi, bstep, cmid, inv_bw, u_first, inv_du; int B; };
__device__ __forceinline__ double epan_bin_centre(const EpanBins& eb, int i) {
  const double e0 = __dadd_rn(__dmul_rn((double)i, eb.bstep), eb.zlo);             // linspace edges (utils/math.py:35)
  const double e1 = (i + 1 == eb.B) ? eb.zhi : __dadd_rn(__dmul_rn((double)(i + 1), eb.bstep), eb.zlo);
  return (e0 + e1) / 2;
}
__device__ __forceinline__ EpanBins make_epan_bins(double zlo, double zhi, int B, double bw) {
  EpanBins eb;
  eb.zlo = zlo; eb.zhi = zhi; eb.B = B; eb.bstep = (zhi - zlo) / (double)B;
  eb.cmid = 0.5 * (zlo + zhi); eb.inv_bw = 1.0 / bw;
  eb.u_first = (epan_bin_centre(eb, 0) - eb.cmid) * eb.inv_bw;
  eb.inv_du = 1.0 / (eb.bstep * eb.inv_bw);
  return eb;
}
// one warp: S0/S1/S2[0..B] from the bin sums (float), normalised by 1/W
__device__ __forceinline__ void epan_bins_prefix(const EpanBins& eb, const float* __restrict__ bins, double invW,
                                                 double* __restrict__ S0, double* __restrict__ S1, double* __restrict__ S2) {
  const int lane = threadIdx.x & 31;
  const int seg = (eb.B + 31) / 32, b0 = min(eb.B, lane * seg), b1 = min(eb.B, b0 + seg);
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  for (int i = b0; i < b1; ++i) {
    const double wn = (double)bins[i] * invW, u = (epan_bin_centre(eb, i) - eb.cmid) * eb.inv_bw;
    t0 += wn; t1 += wn * u; t2 += wn * u * u;
  }
  double e0 = t0, e1 = t1, e2 = t2;                           // inclusive scan of the segment totals, made exclusive below
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y0 = __shfl_up_sync(0xffffffffu, e0, o), y1 = __shfl_up_sync(0xffffffffu, e1, o), y2 = __shfl_up_sync(0xffffffffu, e2, o);
    if (lane >= o) { e0 += y0; e1 += y1; e2 += y2; }
  }
  e0 -= t0; e1 -= t1; e2 -= t2;
  if (lane == 0) { S0[0] = 0.0; S1[0] = 0.0; S2[0] = 0.0; }
  for (int i = b0; i < b1; ++i) {
    const double wn = (double)bins[i] * invW, u = (epan_bin_centre(eb, i) - eb.cmid) * eb.inv_bw;
    e0 += wn; e1 += wn * u; e2 += wn * u * u;
    S0[i + 1] = e0; S1[i + 1] = e1; S2[i + 1] = e2;
  }
}
// the same prefix tables by the WHOLE CTA (the 1-D fused kernel): every thread takes a segment of ceil(B / threads) bins,
// warp scans + the warp totals through `wtot` (3 * FU_NW doubles) -- one warp alone spent ~600 dependent fp64
// instructions here with the other seven waiting on the barrier behind it (6 % of the reference-default configuration).
// Contains one __syncthreads(); the caller's barrier publishes the tables.
__device__ __forceinline__ void epan_bins_prefix_cta(const EpanBins& eb, const float* __restrict__ bins, double invW,
                                                     double* __restrict__ S0, double* __restrict__ S1, double* __restrict__ S2,
                                                     double* __restrict__ wtot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = (eb.B + FU_NT - 1) / FU_NT, b0 = min(eb.B, tid * seg), b1 = min(eb.B, b0 + seg);
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  for (int i = b0; i < b1; ++i) {
    const double wn = (double)bins[i] * invW, u = (epan_bin_centre(eb, i) - eb.cmid) * eb.inv_bw;
    t0 += wn; t1 += wn * u; t2 += wn * u * u;
  }
  double e0 = t0, e1 = t1, e2 = t2;                           // inclusive scan over the warp's threads
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y0 = __shfl_up_sync(0xffffffffu, e0, o), y1 = __shfl_up_sync(0xffffffffu, e1, o), y2 = __shfl_up_sync(0xffffffffu, e2, o);
    if (lane >= o) { e0 += y0; e1 += y1; e2 += y2; }
  }
  if (lane == 31) { wtot[warp * 3] = e0; wtot[warp * 3 + 1] = e1; wtot[warp * 3 + 2] = e2; }
  __syncthreads();
  double o0 = 0.0, o1 = 0.0, o2 = 0.0;
  for (int w = 0; w < warp; ++w) { o0 += wtot[w * 3]; o1 += wtot[w * 3 + 1]; o2 += wtot[w * 3 + 2]; }
  e0 += o0 - t0; e1 += o1 - t1; e2 += o2 - t2;                // exclusive prefix of this thread's segment
  if (tid == 0) { S0[0] = 0.0; S1[0] = 0.0; S2[0] = 0.0; }
  for (int i = b0; i < b1; ++i) {
    const double wn = (double)bins[i] * invW, u = (epan_bin_centre(eb, i) - eb.cmid) * eb.inv_bw;
    e0 += wn; e1 += wn * u; e2 += wn * u * u;
    S0[i + 1] = e0; S1[i + 1] = e1; S2[i + 1] = e2;
  }
}
// sum_b w'_b (1 - ((g - c_b)/bw)^2) over the bins with |g - c_b| <= bw (bins exactly on the edge contribute 0 either way)
__device__ __forceinline__ double epan_bins_sum(const EpanBins& eb, double g, const double* __restrict__ S0,
                                                const double* __restrict__ S1, const double* __restrict__ S2) {
  const double gu = (g - eb.cmid) * eb.inv_bw;
  int lo = (int)ceil((gu - 1.0 - eb.u_first) * eb.inv_du), hi = (int)floor((gu + 1.0 - eb.u_first) * eb.inv_du);
  lo = max(lo, 0); hi = min(hi, eb.B - 1);
  if (!(hi >= lo)) return (gu == gu) ? 0.0 : gu;              // (NaN bandwidth propagates, as in the reference)
  const double d0 = S0[hi + 1] - S0[lo], d1 = S1[hi + 1] - S1[lo], d2 = S2[hi + 1] - S2[lo];
  return (1.0 - gu * gu) * d0 + 2.0 * gu * d1 - d2;
}

// Windowed recurrence KDE of the staged samples (kde_win.cuh) -- NOT inlined, for the same reason as fu_reweight.
// Merges the 64-sample block summaries into the plan's chunks (scaled units, weights normalised), then phases B and C.
static __device__ __noinline__ void fu_kde_win(const float4* __restrict__ stage, int Ns, int G, double gfirst, double hd, int R,
                                        int LPS, int chunk, int nchunks, double scale, float sf, float koff, float t2,
                                        const float4* __restrict__ sub, float4* __restrict__ summ, int2* __restrict__ win,
                                        float* __restrict__ cr, double* __restrict__ rows, double* __restrict__ dens) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WinPlan wp; wp.R = R; wp.LPS = LPS; wp.chunk = chunk; wp.nchunks = nchunks;
  const float h = (float)hd;
  // the per-warp rows phase C adds to: zeroed HERE, in front of the barrier below -- not by the caller, whose other KDE
  // routines (fu_direct, fu_epan_blocks) write this region without a barrier of their own in front (compute-sanitizer
  // racecheck found exactly that write-after-write hazard when the zeroing stood right after the reweighting)
  for (int i = tid; i < FU_NW * G; i += FU_NT) rows[i] = 0.0;
  if (tid < 16) cr[tid] = exp2f(-(h * h) * (float)(tid * (tid - 1)));
  if (warp == 1 && lane < nchunks) {
    float lo = INFINITY, hi = -INFINITY, lm = -INFINITY, xm = 0.f;
    const int nsub = chunk / FU_SUB, b0 = lane * nsub, b1 = min(b0 + nsub, (Ns + FU_SUB - 1) / FU_SUB);
    for (int i = b0; i < b1; ++i) {
      const float4 v = sub[i];
      lo = fminf(lo, v.x); hi = fmaxf(hi, v.y);
      if (v.z > lm) { lm = v.z; xm = v.w; }
    }
    summ[lane] = make_float4(lo * sf, hi * sf, lm + koff, xm * sf);
    win[lane] = make_int2(G, -1);
  }
  __syncthreads();
  kde_win_BC<FU_NW, false, true>(reinterpret_cast<const float2*>(stage), Ns, G, gfirst, hd, wp, scale, summ, win, cr, rows,
                                 dens, sf, koff, t2);
}

__global__ void __launch_bounds__(FU_NT, CHB_FU_MINB)
numerator_fused_kernel(const NumArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ float FC[CHB_NFC];

  const TableLayout lay = a.mc.lay;
  const int Ns = a.Ns, Nz = a.Nz, Pp = a.P, B = a.binning ? a.num_bins : 0, G = Nz / 2;
  const FusedPlan pl = make_fused_plan(Ns, Nz, B);
  float4* stage = reinterpret_cast<float4*>(smraw + pl.stage);
  double* rows = reinterpret_cast<double*>(smraw + pl.rows);
  double* dens = reinterpret_cast<double*>(smraw + pl.dens);
  double* bc = reinterpret_cast<double*>(smraw + pl.bc);
  double* bs = reinterpret_cast<double*>(smraw + pl.bs);
  float4* bx = reinterpret_cast<float4*>(smraw + pl.bx);
  float4* sub = reinterpret_cast<float4*>(smraw + pl.sub);
  float4* summ = reinterpret_cast<float4*>(smraw + pl.summ);
  int2* win = reinterpret_cast<int2*>(smraw + pl.win);
  float* cr = reinterpret_cast<float*>(smraw + pl.cr);
  double* red = reinterpret_cast<double*>(smraw + pl.red);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* inb = reinterpret_cast<float4*>(smraw + pl.rows + warp * FU_INB);      // (the rows are zeroed after the reweighting)
  const bool pixelated = (a.kind != CHB_PGW_1D);
  const bool has_cat = (a.mc.catalog_kind == 1);
  const bool gauss = (a.kernel == CHB_KERNEL_GAUSS);
  const bool want_lw = gauss && !a.binning;                  // the stage carries log2 w for the Gaussian pair sums

  const long long units = (long long)a.Nev * a.n_hyper;
  const double inv_ns = 1.0 / (double)Ns;
  for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int ev = (int)(unit / a.n_hyper), h = (int)(unit % a.n_hyper);
    const double* tblk = a.tabs + (size_t)h * lay.total() + lay.off_f32();
    // (measured, round 2: double-buffering these constants with cp.async, deferring the unit's final sum/log to the next unit
    //  and two grid points per thread-round in the z-integral removed two barriers per unit but cost registers: C3 +-0,
    //  C1 and the binned path 3-4 % slower -- not adopted)
    __syncthreads();                                         // the previous unit is done with every shared array
    for (int i = tid; i < CHB_NPAR + CHB_NHC + CHB_NFC; i += FU_NT) {      // (any thread count: 128 or 256)
      if (i < CHB_NPAR) P[i] = a.hyper[(size_t)h * CHB_NPAR + i];
      else if (i < CHB_NPAR + CHB_NHC) HC[i - CHB_NPAR] = a.HC[(size_t)h * CHB_NHC + i - CHB_NPAR];
      else FC[i - CHB_NPAR - CHB_NHC] = __ldg(reinterpret_cast<const float*>(tblk + lay.f32_fc()) + i - CHB_NPAR - CHB_NHC);
    }
    __syncthreads();

    // ---- stage 1: reweighting ---------------------------------------------------------------
    {
      const size_t so = (size_t)ev * Ns;
      switch (a.mc.mass_model) {
        case CHB_MASS_TPL: fu_reweight<CHB_MASS_TPL>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, want_lw, stage, sub, red, inb); break;
        case CHB_MASS_BPL: fu_reweight<CHB_MASS_BPL>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, want_lw, stage, sub, red, inb); break;
        default: fu_reweight<CHB_MASS_PLP>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, want_lw, stage, sub, red, inb); break;
      }
    }
    __syncthreads();                                         // publishes stage[], sub[] and the warp partials
#if CHB_FU_TAILPF
    if ((tid & 15) == 0) {
      // one hint per 128-byte line (16 doubles / float2) of the rows of the z-integral
      for (int k = tid; k < Nz; k += FU_NT) {
        fu_prefetch_l1(a.zgrids + (size_t)ev * Nz + k);
        if (a.zterms) fu_prefetch_l1(a.zterms + ((size_t)(h - a.zterms_h0) * a.Nev + ev) * Nz + k);
        if (a.kind != CHB_PGW_1D && a.catA) {
          fu_prefetch_l1(a.catA + (size_t)ev * Nz + k); fu_prefetch_l1(a.catB + (size_t)ev * Nz + k);
          if (has_cat) fu_prefetch_l1(a.P_compl + (size_t)ev * Nz + k);
        }
      }
    } else if (tid < 8) {
      // the next unit's constant rows (hyper-parameters 2 lines, host constants, FC)
      const long long nu = unit + gridDim.x;
      if (nu < units) {
        const int nh = (int)(nu % a.n_hyper);
        if (tid < 3) fu_prefetch_l1(a.hyper + (size_t)nh * CHB_NPAR + (tid - 1) * 16);
        else if (tid < 5) fu_prefetch_l1(a.HC + (size_t)nh * CHB_NHC + (tid - 3) * 16);
        else if (tid == 6) fu_prefetch_l1(reinterpret_cast<const float*>(a.tabs + (size_t)nh * lay.total() + lay.off_f32() + lay.f32_fc()));
      }
    }
#endif
    FuStats st = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < FU_NW; ++i) {
      st.a += red[i * 6 + 0]; st.b += red[i * 6 + 1]; st.c += red[i * 6 + 2]; st.d += red[i * 6 + 3];
      st.mn = fminf(st.mn, (float)red[i * 6 + 4]); st.mx = fmaxf(st.mx, (float)red[i * 6 + 5]);
    }
    const float z0 = (float)red[95];
    const double s1 = st.a, s2 = st.b;
    const double zmn = (double)z0 + (double)st.mn, zmx = (double)z0 + (double)st.mx;
    // (fp64 divisions are ~25 instructions each and every thread repeats this section: 1/Ns is taken once per kernel,
    //  1/bw once per unit, the effective sample size of kde1d re-uses the one of the N_eff gate)
    const double dzmean = st.c * inv_ns;
    const double zstd = sqrt(fmax(st.d * inv_ns - dzmean * dzmean, 0.0));    // one-pass variance about z0
    const double norm = s1 * inv_ns;              // likelihood.py:111
    const double neff = s1 * s1 / s2;             // likelihood.py:112
    const bool ok = (neff >= a.pe_neff);

    double* pout = nullptr;
    if (a.p_gw_out) pout = a.p_gw_out + ((size_t)h * a.Nev + ev) * (size_t)(pixelated ? Pp : 1) * Nz;
    const int npix = pixelated ? a.neff_pix[ev] : 1;
    if (!ok) {
      if (pout) for (int i = tid; i < (pixelated ? Pp : 1) * Nz; i += FU_NT) pout[i] = 0.0;
      if (tid == 0) { a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_fu(0.0); a.like_raw[(size_t)h * a.Nev + ev] = 0.0; }
      continue;
    }

    // ---- effective grid (likelihood.py:115-123): linspace(lb, ub, G), never materialised ---------------
    const double lb = (zmn - a.cut_grid * zstd > 0.0) ? zmn - a.cut_grid * zstd : 1.e-8;
    const double ub = zmx + a.cut_grid * zstd;
    const double step = (ub - lb) / (double)(G - 1);
    auto eg_at = [&](int i) -> double { return (i == G - 1) ? ub : __dadd_rn(__dmul_rn((double)i, step), lb); };

    if (!a.binning) {
      // ---- bandwidth (utils/math.py:62-70), every thread the same value ----------------------------
      const double neff_k = neff;                  // 1 / sum((w/W)^2) (utils/math.py:62) = W^2 / sum(w^2)
      double bw;
      if (a.bw_method == CHB_BW_SCOTT) bw = (double)ex2f_(-0.2f * lg2f_((float)neff_k)) * zstd;
      else if (a.bw_method == CHB_BW_SILVERMAN) bw = (double)ex2f_(-0.2f * lg2f_((float)(neff_k * 0.75))) * zstd;
      else bw = a.bw_value * zstd;
      const double inv_bw = 1.0 / bw;
      const double s = (gauss ? 0.8493218002880191 : 1.0) * inv_bw;      // sqrt(log2(e)/2) / bw
      const float sf = (float)s;
      const double gfirst = (lb - (double)z0) * (double)sf, hd = step * (double)sf;
      const double scale = norm * (gauss ? 0.3989422804014327 : 0.75) * inv_bw;
      WinPlan wp;
      const bool windowed = gauss && a.kde_win_iters > 0 && Nz >= 64 &&
                            win_plan(G, Ns, (float)hd, a.kde_win_iters, 32, wp, FU_SUB, a.win_t2);
      if (windowed) {
        fu_kde_win(stage, Ns, G, gfirst, hd, wp.R, wp.LPS, wp.chunk, wp.nchunks, scale, sf, -lg2f_((float)s1), a.win_t2, sub, summ, win,
                   cr, rows, dens);
      } else if (!gauss && a.epan_blocks) {
        fu_epan_blocks(stage, Ns / 2, G, gfirst, hd, sf, scale / s1, reinterpret_cast<float*>(rows), dens);
      } else {
        const float koff = gauss ? -lg2f_((float)s1) : 0.f;
        fu_direct(stage, Ns / 2, G, gfirst, hd, sf, koff, gauss, gauss ? scale : scale / s1, rows, dens);
      }
    } else {
      // ---- binning1d (utils/math.py:32-46) on the shared-memory stage, then the KDE of the B bin centres -------
      float* binsf = reinterpret_cast<float*>(bs);
      for (int i = tid; i < B; i += FU_NT) binsf[i] = 0.f;
      __syncthreads();
      {
        // the samples are sorted by dL, hence by z: a bin is a contiguous run.  Every thread walks a contiguous block
        // of pairs, sums each run in a register and touches shared memory once per run (native fp32 shared atomics).
        // The bin index is taken in fp32 on {z - z0}: z itself is an fp32 number, so the index is as exact as its input.
        const int npairs = Ns / 2;
        const int per = (npairs + FU_NT - 1) / FU_NT;
        const int ja = min(npairs, tid * per), jb = min(npairs, ja + per);
        const float invB = (float)((double)B / (zmx - zmn)), dlo = st.mn;
        int cur = -1;
        float run = 0.f;
        for (int j = ja; j < jb; ++j) {
          const float4 v = stage[j];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float f = floorf(((q ? v.y : v.x) - dlo) * invB);
            if (f != f) continue;
            const int b = (int)fminf(fmaxf(f, 0.f), (float)(B - 1));
            if (b != cur) {
              if (cur >= 0) atomicAdd(&binsf[cur], run);
              cur = b; run = 0.f;
            }
            run += (q ? v.w : v.z);
          }
        }
        if (cur >= 0) atomicAdd(&binsf[cur], run);
      }
      __syncthreads();
      const EpanBins eb0 = make_epan_bins(zmn, zmx, B, 1.0);
      FuStats u = {0.0, 0.0, 0.0, 0.0, 0.f, 0.f};
      for (int i = tid; i < B; i += FU_NT) { const double bv = (double)binsf[i], cc = epan_bin_centre(eb0, i); u.a += bv; u.b += bv * bv; u.c += cc; u.d += cc * cc; }
      u = fu_block_stats(u, red);                            // (`red` was last read before several barriers)
      const double W = u.a, Q = u.b;
      const double cmean = u.c / B;
      const double dstd = sqrt(fmax(u.d / B - cmean * cmean, 0.0));   // std of the BIN CENTRES (math.py:67 on binning1d's output)
      const double neff_k = 1.0 / (Q / (W * W));
      double bw;
      // (n^-1/5 through MUFU lg2/ex2 like the unbinned branch: ~1e-6 relative on the bandwidth, the fp32 mode's own level,
      //  instead of ~200 fp64 instructions of pow() per thread)
      if (a.bw_method == CHB_BW_SCOTT) bw = (double)ex2f_(-0.2f * lg2f_((float)neff_k)) * dstd;
      else if (a.bw_method == CHB_BW_SILVERMAN) bw = (double)ex2f_(-0.2f * lg2f_((float)(neff_k * 3.0 / 4.0))) * dstd;
      else bw = a.bw_value * dstd;
      if (!gauss) {
        // Epanechnikov (the reference's default kernel): prefix sums over the bins, O(1) per grid point
        const EpanBins eb = make_epan_bins(zmn, zmx, B, bw);
        double* S0 = bc; double* S1 = bc + (B + 2); double* S2 = bc + 2 * (B + 2);
        epan_bins_prefix_cta(eb, binsf, 1.0 / W, S0, S1, S2, red + 6 * FU_NW);      // (red[0 .. 6 FU_NW) may still be read: fu_block_stats)
        __syncthreads();
        const double scale = norm * 0.75 / bw;
        for (int g = tid; g < G; g += FU_NT) dens[g] = epan_bins_sum(eb, eg_at(g), S0, S1, S2) * scale;
      } else {
        const double s = 0.8493218002880191 / bw;
        const double c = 0.5 * (lb + ub);
        // pair layout of the binned data set: {x'_a, x'_b, v_a, v_b}, x' = (centre - c) s, v = log2(w/W)
        const int Bp = (B + 1) / 2;
        for (int i = tid; i < Bp; i += FU_NT) {
          const int ia = 2 * i, ib = 2 * i + 1;
          const float xa = (float)((epan_bin_centre(eb0, ia) - c) * s), xb = (ib < B) ? (float)((epan_bin_centre(eb0, ib) - c) * s) : 0.f;
          const float wa = (float)((double)binsf[ia] / W), wb = (ib < B) ? (float)((double)binsf[ib] / W) : 0.f;
          bx[i] = make_float4(xa, xb, wa > 0.f ? lg2f_(wa) : -INFINITY, wb > 0.f ? lg2f_(wb) : -INFINITY);
        }
        __syncthreads();
        fu_direct(bx, Bp, G, (lb - c) * s, step * s, 1.f, 0.f, true, norm * 0.3989422804014327 / bw, rows, dens);
      }
    }
    __syncthreads();

    // ---- p_gw on the event grid (likelihood.py:139-141) fused with the integrand (likelihood.py:266-292) --------
    const double* zgr = a.zgrids + (size_t)ev * Nz;
    const double inv_step = 1.0 / step;
    auto pgw_at = [&](double x) -> double {
      if (x < lb || x > ub) return 0.0;
      int i = (int)((x - lb) * inv_step);
      i = max(0, min(i, G - 2));
      while (i < G - 2 && x >= eg_at(i + 1)) ++i;
      while (i > 0 && x < eg_at(i)) --i;
      const double x0 = eg_at(i), dx = eg_at(i + 1) - x0, f0 = dens[i], df = dens[i + 1] - f0;
      return (fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((x - x0) / dx) * df;
    };
    const float2* zt = a.zterms ? a.zterms + ((size_t)(h - a.zterms_h0) * a.Nev + ev) * Nz : nullptr;
    auto zterms_at = [&](int k, double z) -> float2 {
      if (zt) return __ldg(zt + k);
      const double zl = (k > 0) ? zgr[k - 1] : z, zr = (k < Nz - 1) ? zgr[k + 1] : z;
      const F32Consts fc = make_f32_consts(a.mc, P, HC, tblk);       // rare path: the precomputed terms did not fit
      return zgrid_terms_f32(fc, make_cosmo_rate_f32(a.mc, P, HC), P, HC, a.mc.cosmo_model, z, 0.5 * (zr - zl));
    };
    const double fR = HC[HC_FR];
    const double* pcompl_ev = has_cat ? a.P_compl + (size_t)ev * Nz : nullptr;
    double like_acc = 0.0;
    // (measured: restricting this loop to the event-grid points inside the effective grid (fu_grid_range, as the
    //  'marginalized' kernel does per unit for all its pixels) costs more than it saves here: C3 +2.5 %, C1 +4.6 %)
    if (a.kind == CHB_PGW_1D) {
      for (int k = tid; k < Nz; k += FU_NT) {
        const double z = zgr[k], pg = pgw_at(z);
        const float2 v = zterms_at(k, z);
        like_acc += pg * (double)v.x * (double)v.y;
        if (pout) pout[k] = pg;
      }
    } else if (a.catA) {
      const double* A = a.catA + (size_t)ev * Nz;
      const double* Bk = a.catB + (size_t)ev * Nz;
      const double* gwp = a.gw_pdf + (size_t)ev * Pp;
      for (int k = tid; k < Nz; k += FU_NT) {
        const double z = zgr[k], pg = pgw_at(z);
        const float2 v = zterms_at(k, z);
        const double dVk = (double)v.x;
        const double pgs = has_cat ? fR * A[k] + (1.0 - pcompl_ev[k]) * dVk * Bk[k] : dVk * Bk[k];
        like_acc += pg * pgs * (double)v.y;
        if (pout) for (int p = 0; p < Pp; ++p) pout[(size_t)p * Nz + k] = pg * gwp[p];
      }
    } else {
      const double* gwp = a.gw_pdf + (size_t)ev * Pp;
      const double* pcat_ev = has_cat ? a.p_cat + (size_t)ev * Pp * Nz : nullptr;
      for (int k = tid; k < Nz; k += FU_NT) {
        const double z = zgr[k], pg = pgw_at(z);
        const float2 v = zterms_at(k, z);
        const double dVk = (double)v.x, ckk = (double)v.y;
        for (int p = 0; p < Pp; ++p) {
          const double pv = pg * gwp[p];
          if (pout) pout[(size_t)p * Nz + k] = pv;
          if (p >= npix) continue;
          const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
          if (pc == -100.0) continue;
          const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dVk : dVk;
          like_acc += pv * pgal * ckk;
        }
      }
    }
    {
      const double ws = warp_sum(like_acc);
      if (lane == 0) red[warp] = ws;
      __syncthreads();
      if (tid == 0) {
        double like = 0.0;
#pragma unroll
        for (int w = 0; w < FU_NW; ++w) like += red[w];
        a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_fu(like); a.like_raw[(size_t)h * a.Nev + ev] = like;
      }
    }
  }
}

#ifndef CHB_FU_VARIANT
// ==================================================================================================================
// 'marginalized' kind with binning (the reference's default options, examples/test1dgalaxies.ipynb): one fused kernel,
// warp per pixel.  likelihood.py:160-205: per pixel the in-pixel samples are binned on [min z (all samples), max z (pixel)]
// (utils/math.py:32-46), the ALWAYS-Epanechnikov kde1d of the B bin centres is evaluated on the effective grid of the
// unmasked statistics, interpolated onto the event grid and integrated against the pixel's catalogue row.
//   * the samples arrive bucketed by pixel (api.cu prepare()), so a pixel is a contiguous range of the stage;
//   * the Epanechnikov kernel has compact support and the bin centres are equally spaced: the sum over the bins inside
//     |g - c_b| <= bw is a difference of prefix sums of {w, w u, w u^2} (fp64, coordinates centred on the pixel and scaled
//     by 1/bw), O(B + Nz) per pixel instead of O(B G) -- and no grid array: the two effective-grid values an event-grid
//     point interpolates between are evaluated on the fly;
//   * no CTA barrier inside the pixel loop (round 1: ~10 barriers per pixel, 15 pixels per unit).
struct MargPlan { int stage, scratch, per_warp, red, total; };
__host__ __device__ inline MargPlan make_marg_plan(int Ns, int B, int G) {
  MargPlan p;
  const int Bp = (B + 32) & ~31;                 // >= B + 1 entries for the inclusive prefix tables
  int o = 0;
  p.stage = o; o += Ns * 8;
  p.scratch = o; p.per_warp = max(Bp * 4 + 3 * Bp * 8 + ((G + 1) & ~1) * 8, FU_INB);   // bins (float) + S0, S1, S2 + dens (double); >= the cp.async landing zone
  o += FU_NW * p.per_warp;
  p.red = o; o += 96 * 8;
  p.total = o;
  return p;
}
size_t numerator_marg_smem_bytes(const NumArgs& a) { return (size_t)make_marg_plan(a.Ns, a.num_bins, a.Nz / 2).total; }
bool numerator_marg_supported(const NumArgs& a) {
  return a.fp_mode == CHB_FP32 && a.kind == CHB_PGW_MARG && a.binning && a.use_cut && (a.Ns % 2 == 0) && a.Nz / 2 >= 2 &&
         a.num_bins >= 2 && a.s4 != nullptr && a.pix_off != nullptr;
}

__global__ void __launch_bounds__(FU_NT, 2)
numerator_marg_kernel(const NumArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ float FC[CHB_NFC];
  const TableLayout lay = a.mc.lay;
  const int Ns = a.Ns, Nz = a.Nz, Pp = a.P, B = a.num_bins, G = Nz / 2;
  const MargPlan pl = make_marg_plan(Ns, B, G);
  float4* stage = reinterpret_cast<float4*>(smraw + pl.stage);
  const float* stf = reinterpret_cast<const float*>(stage);
  double* red = reinterpret_cast<double*>(smraw + pl.red);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Bp = (B + 32) & ~31;
  float* bins = reinterpret_cast<float*>(smraw + pl.scratch + warp * pl.per_warp);
  double* S0 = reinterpret_cast<double*>(bins + Bp);
  double* S1 = S0 + Bp;
  double* S2 = S1 + Bp;
  double* dn = S2 + Bp;                          // the pixel's KDE on the effective grid (G doubles)
  float4* inb = reinterpret_cast<float4*>(bins);  // cp.async landing zone of the reweighting (the scratch is idle then)
  const bool has_cat = (a.mc.catalog_kind == 1);
  // sample j of the stage: pair j/2, slot j&1 of {dz_a, dz_b, w_a, w_b}
  auto dz_of = [&](int j) -> float { return stf[4 * (j >> 1) + (j & 1)]; };
  auto w_of = [&](int j) -> float { return stf[4 * (j >> 1) + 2 + (j & 1)]; };

  const long long units = (long long)a.Nev * a.n_hyper;
  for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int ev = (int)(unit / a.n_hyper), h = (int)(unit % a.n_hyper);
    const double* tblk = a.tabs + (size_t)h * lay.total() + lay.off_f32();
    __syncthreads();
    if (tid < CHB_NPAR) P[tid] = a.hyper[(size_t)h * CHB_NPAR + tid];
    else if (tid >= 64 && tid < 64 + CHB_NHC) HC[tid - 64] = a.HC[(size_t)h * CHB_NHC + tid - 64];
    else if (tid >= 128 && tid < 128 + CHB_NFC) FC[tid - 128] = __ldg(reinterpret_cast<const float*>(tblk + lay.f32_fc()) + tid - 128);
    __syncthreads();
    {
      const size_t so = (size_t)ev * Ns;
      switch (a.mc.mass_model) {
        case CHB_MASS_TPL: fu_reweight<CHB_MASS_TPL>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, 0, stage, nullptr, red, inb); break;
        case CHB_MASS_BPL: fu_reweight<CHB_MASS_BPL>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, 0, stage, nullptr, red, inb); break;
        default: fu_reweight<CHB_MASS_PLP>(tblk, lay.rc, lay.rcs, lay.rms, lay.rm, FC, a.s4 + so, a.l2 + so, Ns, 0, stage, nullptr, red, inb); break;
      }
    }
    __syncthreads();
    FuStats st = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < FU_NW; ++i) {
      st.a += red[i * 6 + 0]; st.b += red[i * 6 + 1]; st.c += red[i * 6 + 2]; st.d += red[i * 6 + 3];
      st.mn = fminf(st.mn, (float)red[i * 6 + 4]); st.mx = fmaxf(st.mx, (float)red[i * 6 + 5]);
    }
    const double z0 = red[95];
    const double s1 = st.a, s2 = st.b;
    const double zmn = z0 + (double)st.mn, zmx = z0 + (double)st.mx;
    const double dzmean = st.c / Ns;
    const double zstd = sqrt(fmax(st.d / Ns - dzmean * dzmean, 0.0));
    const double norm = s1 / Ns;
    const double neff = s1 * s1 / s2;
    double* pout = a.p_gw_out ? a.p_gw_out + ((size_t)h * a.Nev + ev) * (size_t)Pp * Nz : nullptr;
    if (!(neff >= a.pe_neff)) {
      if (pout) for (int i = tid; i < Pp * Nz; i += FU_NT) pout[i] = 0.0;
      if (tid == 0) { a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_fu(0.0); a.like_raw[(size_t)h * a.Nev + ev] = 0.0; }
      continue;
    }
    // effective grid of the UNMASKED statistics (likelihood.py:186-190)
    const double lb = (zmn - a.cut_grid * zstd > 0.0) ? zmn - a.cut_grid * zstd : 1.e-8;
    const double ub = zmx + a.cut_grid * zstd;
    const double step = (ub - lb) / (double)(G - 1), inv_step = 1.0 / step;
    auto eg_at = [&](int i) -> double { return (i == G - 1) ? ub : __dadd_rn(__dmul_rn((double)i, step), lb); };
    const int npix = a.neff_pix[ev];
    const int* off = a.pix_off + (size_t)ev * (Pp + 2);
    const double* gwp = a.gw_pdf + (size_t)ev * Pp;
    const double* zgr = a.zgrids + (size_t)ev * Nz;
    const float2* zt = a.zterms ? a.zterms + ((size_t)(h - a.zterms_h0) * a.Nev + ev) * Nz : nullptr;
    const double fR = HC[HC_FR];
    const double* pcat_ev = has_cat ? a.p_cat + (size_t)ev * Pp * Nz : nullptr;
    const double* pcompl_ev = has_cat ? a.P_compl + (size_t)ev * Nz : nullptr;
    if (pout) {
      for (int i = tid + npix * Nz; i < Pp * Nz; i += FU_NT) pout[i] = 0.0;       // padded pixel rows
    }
    double like_acc = 0.0;
    int k0, k1;
    fu_grid_range(zgr, Nz, lb, ub, k0, k1);
    for (int p = warp; p < npix; p += FU_NW) {
      const int o0 = off[p], o1 = off[p + 1];
      // ---- masked data set of the pixel: max z, then binning1d on [zmn, zmax_in] ------------------------
      float mxf = -INFINITY;
      for (int j = o0 + lane; j < o1; j += 32) mxf = fmaxf(mxf, dz_of(j));
      mxf = warp_max_f32(mxf);
      const double zmax_in = fmax(z0 + (double)mxf, zmn);                 // masked samples sit at min(z)
      for (int i = lane; i < Bp; i += 32) bins[i] = 0.f;
      __syncwarp();
      const double brange = zmax_in - zmn, binv = (double)B / brange;
      for (int j = o0 + lane; j < o1; j += 32) {
        const double f = floor(((z0 + (double)dz_of(j)) - zmn) * binv);
        if (!isnan(f)) atomicAdd(&bins[(int)fmin(fmax(f, 0.0), (double)(B - 1))], w_of(j));
      }
      __syncwarp();
      // ---- kde1d of the bin centres: w/W, neff, bandwidth from the std of the CENTRES (math.py:62-70) -------
      const EpanBins eb0 = make_epan_bins(zmn, zmax_in, B, 1.0);
      double W = 0.0, Q = 0.0, sc = 0.0, sc2 = 0.0;
      for (int i = lane; i < B; i += 32) { const double b = (double)bins[i], c = epan_bin_centre(eb0, i); W += b; Q += b * b; sc += c; sc2 += c * c; }
      W = warp_sum(W); Q = warp_sum(Q); sc = warp_sum(sc); sc2 = warp_sum(sc2);
      const double cmean = sc / B;
      const double dstd = sqrt(fmax(sc2 / B - cmean * cmean, 0.0));
      const double neff_k = 1.0 / (Q / (W * W));
      double bw;
      // (n^-1/5 through MUFU lg2/ex2 like the unbinned branch: ~1e-6 relative on the bandwidth, the fp32 mode's own level,
      //  instead of ~200 fp64 instructions of pow() per thread)
      if (a.bw_method == CHB_BW_SCOTT) bw = (double)ex2f_(-0.2f * lg2f_((float)neff_k)) * dstd;
      else if (a.bw_method == CHB_BW_SILVERMAN) bw = (double)ex2f_(-0.2f * lg2f_((float)(neff_k * 3.0 / 4.0))) * dstd;
      else bw = a.bw_value * dstd;
      // W == 0 (no weight in the pixel) -> w/W = NaN for every sample in the reference
      const double scale = (W != 0.0) ? norm * gwp[p] * 0.75 / bw : nan("");
      const EpanBins eb = make_epan_bins(zmn, zmax_in, B, bw);
      epan_bins_prefix(eb, bins, 1.0 / W, S0, S1, S2);
      __syncwarp();
      for (int g = lane; g < G; g += 32) dn[g] = epan_bins_sum(eb, eg_at(g), S0, S1, S2);
      __syncwarp();
      // z-integral of the pixel: the four rows of an iteration (event grid, catalogue row, z-grid terms, completeness)
      // are requested one iteration ahead -- they do not depend on the KDE, and 28 % of the kernel's warp samples
      // sat on their L2 latency when each iteration loaded them right before use
      auto rows_at = [&](int k, double& x, double& pc, float2& zv, double& pcm) {
        x = zgr[k];
        pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
        zv = zt ? __ldg(zt + k) : make_float2(0.f, 0.f);
        pcm = has_cat ? pcompl_ev[k] : 0.0;
      };
      if (pout) {                                              // p_gw is zero outside the effective grid
        for (int kk = lane; kk < k0; kk += 32) pout[(size_t)p * Nz + kk] = 0.0;
        for (int kk = k1 + lane; kk < Nz; kk += 32) pout[(size_t)p * Nz + kk] = 0.0;
      }
      int k = k0 + lane;
      double x = 0.0, pc = 0.0, pcm = 0.0;
      float2 zv = make_float2(0.f, 0.f);
      if (k < k1) rows_at(k, x, pc, zv, pcm);
      while (k < k1) {
        const int kn = k + 32;
        double xn = 0.0, pcn = 0.0, pcmn = 0.0;
        float2 zvn = make_float2(0.f, 0.f);
        if (kn < k1) rows_at(kn, xn, pcn, zvn, pcmn);
        double v = 0.0;
        if (x >= lb && x <= ub) {
          int i = (int)((x - lb) * inv_step);
          i = max(0, min(i, G - 2));
          while (i < G - 2 && x >= eg_at(i + 1)) ++i;
          while (i > 0 && x < eg_at(i)) --i;
          const double x0 = eg_at(i), x1 = eg_at(i + 1), dx = x1 - x0;
          const double f0 = dn[i], f1 = dn[i + 1];
          v = ((fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((x - x0) / dx) * (f1 - f0)) * scale;
        }
        if (pout) pout[(size_t)p * Nz + k] = v;
        if (pc != -100.0) {
          if (!zt) {
            const double zl = (k > 0) ? zgr[k - 1] : x, zr = (k < Nz - 1) ? zgr[k + 1] : x;
            const F32Consts fc = make_f32_consts(a.mc, P, HC, tblk);
            zv = zgrid_terms_f32(fc, make_cosmo_rate_f32(a.mc, P, HC), P, HC, a.mc.cosmo_model, x, 0.5 * (zr - zl));
          }
          const double pgal = has_cat ? fR * pc + (1.0 - pcm) * (double)zv.x : (double)zv.x;
          like_acc += v * pgal * (double)zv.y;
        }
        k = kn; x = xn; pc = pcn; zv = zvn; pcm = pcmn;
      }
      __syncwarp();
    }
    {
      const double ws = warp_sum(like_acc);
      __syncthreads();                                       // every warp is done reading `red` (statistics)
      if (lane == 0) red[warp] = ws;
      __syncthreads();
      if (tid == 0) {
        double like = 0.0;
#pragma unroll
        for (int w = 0; w < FU_NW; ++w) like += red[w];
        a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_fu(like); a.like_raw[(size_t)h * a.Nev + ev] = like;
      }
    }
  }
}

cudaError_t numerator_marg_configure(size_t optin) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, numerator_marg_kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(numerator_marg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - fa.sharedSizeBytes));
}
int numerator_marg_ctas_per_sm(size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, numerator_marg_kernel, FU_NT, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
cudaError_t launch_numerator_marg(const NumArgs& a, int grid, size_t smem, cudaStream_t s) {
  numerator_marg_kernel<<<grid, FU_NT, smem, s>>>(a);
  return cudaGetLastError();
}

#endif  // CHB_FU_VARIANT

cudaError_t numerator_fused_configure(size_t optin) {      // see numerator_f32_configure
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, numerator_fused_kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(numerator_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - fa.sharedSizeBytes));
}
int numerator_fused_ctas_per_sm(size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, numerator_fused_kernel, FU_NT, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
cudaError_t launch_numerator_fused(const NumArgs& a, int grid, size_t smem, cudaStream_t s) {
  numerator_fused_kernel<<<grid, FU_NT, smem, s>>>(a);
  return cudaGetLastError();
}
