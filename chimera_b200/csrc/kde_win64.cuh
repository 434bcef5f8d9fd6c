// kde_win64.cuh -- the windowed recurrence Gaussian KDE of kde_win.cuh in DOUBLE precision (CHB_FP64 mode).
//
// Same sum as kde1d (utils/math.py:52-81) on the uniform effective grid of likelihood.py:115-123,
//     dens[g] = scale * sum_j w'_j 2^-(g' - x'_j)^2,   x' = (x - c) sqrt(log2(e)/2)/bw,  w' = w/W,
// and the same organisation as the fp32 kernel: chunks of (sorted) samples only visit the grid points where they
// can matter, and along a lane's run of R consecutive grid points the Gaussian advances by the recurrence
//     2^-((d + r h)^2) = 2^-(d^2) q^r c^(r(r-1)/2),   q = 2^-(2 h d + h^2),   c = 2^-(2 h^2)
// -- 2 fp64 exp2 per R pairs instead of one exp per pair, ~1/3 of the pairs.  Both are exact algebra; in fp64 the
// recurrence error is ~R ulp, the window threshold is 2^-64 of the largest term AT THAT GRID POINT (relative error
// < 1e-17 on every grid value, tails included), and the exponent range needs no rescaling.  The bounds of phase B
// are the fp32 ones (they only have to be bounds).  Correctness never depends on the sample order, only the size of
// the windows does.
#pragma once
#include "kde_win.cuh"

// 2^x for finite x or x = -inf, |error| < 2e-16 relative: n = rint(x), f = x - n in [-1/2, 1/2], degree-12 Taylor
// polynomial of 2^f = e^(f ln 2) (remainder (ln2/2)^13/13! = 1.7e-16), exponent added to the result's bits.  No special
// cases beyond underflow to 0 (x < -1021) -- what the pair sums need; ~20 instructions instead of ~50 for exp2().
__device__ __forceinline__ double exp2_fast(double x) {
  if (!(x > -1021.0)) return 0.0;
  x = fmin(x, 1023.0);
  const double n = rint(x), f = x - n;
  double p = 2.5678435993488196e-11;                       // c_k = ln2^k / k!, k = 12 .. 0 (Horner)
  p = fma(p, f, 4.4455382718708101e-10);
  p = fma(p, f, 7.0549116208011209e-09);
  p = fma(p, f, 1.0178086009239696e-07);
  p = fma(p, f, 1.3215486790144305e-06);
  p = fma(p, f, 1.5252733804059838e-05);
  p = fma(p, f, 1.5403530393381606e-04);
  p = fma(p, f, 1.3333558146428441e-03);
  p = fma(p, f, 9.6181291076284769e-03);
  p = fma(p, f, 5.5504108664821576e-02);
  p = fma(p, f, 2.4022650695910069e-01);
  p = fma(p, f, 6.9314718055994529e-01);
  p = fma(p, f, 1.0);
  const int hi = __double2hiint(p) + ((int)n << 20);       // p in [0.70, 1.42]: adding n to the exponent cannot wrap for |n| <= 1023
  return __hiloint2double(hi, __double2loint(p));
}

#define CHB_WIN64_T2 64.0f
#define CHB_WIN64_R 8
#define CHB_WIN64_LPS 4

// One pass: LPS lanes x R grid points per sample, 32/LPS sample sub-streams per warp.  xs = x' (scaled), lws = log2 w'.
template <int R, int LPS>
__device__ __forceinline__ void kde_win64_pass(const double* __restrict__ xs, const double* __restrict__ lws, int cb, int ce,
                                               int gb, int glast, double gfirst, double hd, const double* __restrict__ crd,
                                               double* __restrict__ row) {
  constexpr int S = 32 / LPS;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPS, sub = lane / LPS;
  const int g0 = gb + gl * R;
  const double gp = gfirst + (double)g0 * hd;
  const double m2h = -2.0 * hd, mh2 = -hd * hd;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  for (int j = cb + sub; j < ce; j += S) {
    const double d = gp - xs[j];
    const double e0 = exp2_fast(fma(-d, d, lws[j]));            // w' 2^-(d^2)   (lw = -inf -> 0)
    const double q = exp2_fast(fmin(fma(d, m2h, mh2), 900.0 / (double)R));
    double p = e0;
    acc[0] += p;
#pragma unroll
    for (int r = 1; r < R; ++r) { p *= q; acc[r] += p; }
  }
  // combine the sample sub-streams (lane bits >= LPS)
#pragma unroll
  for (int o = LPS; o < 32; o <<= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  }
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int g = g0 + r;
      if (g <= glast) row[g] += acc[r] * crd[r];
    }
  }
}

// Whole-CTA call (NW warps).  zs/ws: in {z, w} (any order), out {x', log2 w'}.  Scratch: summ (32 float4) + win (32 int2),
// crd (16 doubles), rows (NW*G doubles).  Returns false (nothing touched) when windows cannot pay.
template <int NW>
__device__ __forceinline__ bool kde1d_f64_win(double* __restrict__ zs, double* __restrict__ ws, int n, int G, double lb,
                                              double step, double bw, double W, double scale, float4* __restrict__ summ,
                                              int2* __restrict__ win, double* __restrict__ crd, double* __restrict__ rows,
                                              double* __restrict__ dens) {
  constexpr int R = CHB_WIN64_R, LPS = CHB_WIN64_LPS, Wp = R * LPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double s = 0.8493218002880191 / bw;                     // sqrt(log2(e)/2) / bw
  const double c = lb + 0.5 * step * (double)(G - 1);
  const double gfirst = (lb - c) * s, hd = step * s;
  const float h = (float)hd;
  // the recurrence needs a run to stay inside the fp64 range (R h <= ~15); windows as wide as the grid are fine -- the
  // recurrence alone is 4x fewer exps than the exhaustive sums
  if (!(h > 0.f) || h > 1.8f || n < 512 || G < 32) return false;
  int chunk = ((n + 31) / 32 + 31) / 32 * 32;                  // <= 32 chunks, multiples of 32 samples
  chunk = max(chunk, 64);
  const int nchunks = (n + chunk - 1) / chunk;
  const double lg2invW = -log2(W);
  // ---- phase A: rescale in place + chunk summaries --------------------------------------------
  for (int i = threadIdx.x; i < NW * G; i += NW * 32) rows[i] = 0.0;
  if (threadIdx.x < 16) crd[threadIdx.x] = exp2(-(hd * hd) * (double)(threadIdx.x * (threadIdx.x - 1)));
  for (int ck = warp; ck < nchunks; ck += NW) {
    const int cb = ck * chunk, ce = min(n, cb + chunk);
    float lo = INFINITY, hi = -INFINITY, lm = -INFINITY, xm = 0.f;
    for (int j = cb + lane; j < ce; j += 32) {
      const double x = (zs[j] - c) * s, w = ws[j];
      const bool live = w > 0.0;                                 // zero / NaN weights add exactly 0
      const double lw = live ? log2(w) + lg2invW : -INFINITY;
      zs[j] = x; ws[j] = lw;
      const float xf = (float)x, lwf = (float)lw;
      if (live) {
        lo = fminf(lo, xf); hi = fmaxf(hi, xf);
        if (lwf > lm) { lm = lwf; xm = xf; }
      }
    }
    lo = warp_min_f32(lo); hi = warp_max_f32(hi);
    const float lmw = warp_max_f32(lm);
    const unsigned pick = __ballot_sync(0xffffffffu, lm == lmw && lm > -INFINITY);
    const float xmw = __shfl_sync(0xffffffffu, xm, pick ? (__ffs(pick) - 1) : 0);
    if (lane == 0) {
      // float bounds of double data: widen by an ulp-scale margin so that they stay bounds
      summ[ck] = make_float4(lo - 1e-5f * fabsf(lo) - 1e-6f, hi + 1e-5f * fabsf(hi) + 1e-6f, pick ? lmw + 1e-4f : -INFINITY, xmw);
      win[ck] = make_int2(G, -1);
    }
  }
  __syncthreads();
  // ---- phase B: hull of the grid points every chunk can matter at (kde_win.cuh phase B, threshold 2^-64) ------
  const float lgchunk = lg2f_((float)chunk);
  const float4 my = (lane < nchunks) ? summ[lane] : make_float4(INFINITY, -INFINITY, -INFINITY, 0.f);
  {
    const float myU = my.z + lgchunk, gf = (float)gfirst;
    int gmin = G, gmax = -1;
    for (int g = warp; g < G; g += NW) {
      const float gp = fmaf((float)g, h, gf);
      const float d = gp - my.w;
      const float m = warp_max_f32(fmaf(-d, d, my.z - 2e-4f));      // lower bound of the largest term (the margin undone)
      const float dist = fmaxf(fmaxf(my.x - gp, gp - my.y), 0.f);
      if (fmaf(-dist, dist, myU) >= m - CHB_WIN64_T2 && my.z > -INFINITY) { gmin = min(gmin, g); gmax = g; }
    }
    if (gmax >= 0) { atomicMin(&win[lane].x, gmin); atomicMax(&win[lane].y, gmax); }
  }
  __syncthreads();
  // ---- phase C: the passes of all chunks form one list, cut into NW contiguous slices (deterministic) ----------
  const int2 w2 = (lane < nchunks) ? win[lane] : make_int2(G, -1);
  const int np = (w2.x <= w2.y) ? (w2.y - w2.x + Wp) / Wp : 0;
  int pend = np;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pend, o); if (lane >= o) pend += y; }
  {
    double* row = rows + warp * G;
    const int total = __shfl_sync(0xffffffffu, pend, 31);
    const int i0 = (int)(((long long)warp * total) / NW), i1 = (int)(((long long)(warp + 1) * total) / NW);
    for (int it = i0; it < i1; ++it) {
      const int ck = __popc(__ballot_sync(0xffffffffu, pend <= it));
      const int wx = __shfl_sync(0xffffffffu, w2.x, ck), wy = __shfl_sync(0xffffffffu, w2.y, ck);
      const int first = __shfl_sync(0xffffffffu, pend - np, ck);
      const int gb = wx + (it - first) * Wp;
      const int cb = ck * chunk, ce = min(n, cb + chunk);
      kde_win64_pass<R, LPS>(zs, ws, cb, ce, gb, wy, gfirst, hd, crd, row);
      __syncwarp();
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += NW * 32) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += rows[w * G + g];
    dens[g] = acc * scale;
  }
  return true;
}
