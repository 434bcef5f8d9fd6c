// kde_win.cuh -- windowed Gaussian KDE pair sums for (nearly) SORTED samples on a uniform grid.
//
// Same sum as kde1d (utils/math.py:52-81) on the effective grid of likelihood.py:115-123:
//     dens[g] = scale * sum_j w'_j 2^-(g' - x'_j)^2,   x' = (x - c) sqrt(log2(e)/2)/bw,  w' = w/W
// organised so that every chunk of consecutive samples only visits the grid points where it can matter.
//
//  * The samples of an event are sorted by dL once at upload (api.cu prepare()); z_from_dGW is monotone
//    in dL (cosmo.py:260-264), so the reweighted z's arrive sorted for every hyper-point and a chunk of
//    64..256 consecutive samples spans only a few grid spacings.
//  * Phase A (one pass over the samples): rescale {z, w} -> {x', log2 w'} in place and record per chunk
//    {min x', max x', max log2 w', x' of that heaviest sample}.
//  * Phase B (chunks x grid points, ~6000 cheap tests): a LOWER bound of the largest term at grid point g,
//    M(g) = max_c [lw*_c - (g - x*_c)^2] from the chunks' heaviest samples, and for every chunk the hull of
//    the grid points where its UPPER bound lwmax_c + log2(chunk) - dist(g, chunk)^2 reaches M(g) - T2.
//    Everything outside that hull is below 2^-T2 = 1e-9 of the largest single term AT THAT GRID POINT, so the
//    density keeps its relative accuracy in the far tails and in gaps -- the parts of the KDE that decide
//    log-likelihoods of events sitting at a catalogue edge.  Correctness never depends on the sample order;
//    only the size of the windows does.
//  * Phase C: every warp takes chunks round-robin; a chunk's window is covered by passes of LPS lanes x R
//    consecutive grid points per sample, 32/LPS samples in flight per warp.  Along a lane's run the Gaussian
//    is advanced by the recurrence  2^-((d + r h)^2) = 2^-(d^2) q^r c^(r(r-1)/2),  q = 2^-(2 h d + h^2),
//    with the sample-independent c^(r(r-1)/2) applied once per pass: 2 MUFU.EX2 per R pairs, and with the
//    powers q, q^2, q^3, q^4 the accumulation costs R + R/4 + 2 FP32 instructions per R pairs.
//    Runs that lie left of their chunk are walked right-to-left so that every run STARTS at its point nearest
//    to the samples; when the nearest possible term is below 2^-64 the lane adds an integer exponent offset K
//    so that the start value cannot flush to zero.  Partial sums go to a per-warp row of doubles (exact
//    rescaling by 2^-K, no atomics, bit-reproducible).
//  * What is guaranteed (checked on the CPU by tests/test_window_algorithm.py with a NumPy model of this file, and on
//    the GPU against the fp64 oracle): absolute error < 2e-6 of the peak everywhere, relative error at fp32 level
//    wherever the density exceeds 1e-8 of the peak AND at every grid point outside the span of the samples down to
//    1e-120 of the peak for every tiling (1e-280 on the fine grids of the default bandwidth: a run's far end is
//    carried in fp32 while 2 h d (R-1) stays below ~200 bits).  Not guaranteed: relative accuracy inside a gap BETWEEN samples that is wider than ~12 scaled
//    units (14 bandwidths), where a run that starts > 11.2 units from an isolated sample flushes its start value
//    (terms below 2^-32 of that sample's weight) -- the same floor the full-grid recurrence of kde_f32.cuh has.
#pragma once
#include "kde_f32.cuh"

#ifndef CHB_WIN_T2
#define CHB_WIN_T2 30.0f          // terms below 2^-30 of the largest term at a grid point are dropped (A/B: -DCHB_WIN_T2=24.0f)
#endif
#define CHB_WIN_MAXR 16
#ifndef CHB_WIN_SPAN
#define CHB_WIN_SPAN 5          // grid points allowed for the spread of a chunk when the tiling is chosen
#endif

__device__ __forceinline__ float warp_min_f32(float v) {
  float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {
  float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r;
}

// Blackwell packed-FP32 arithmetic (PTX f32x2 -> SASS FADD2 / FMUL2 / FFMA2): two samples ride in the two halves of a
// 64-bit register pair, so the recurrence of the pair sums issues one instruction for two samples.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct WinPlan { int R, LPS, chunk, nchunks; };

// Tiling: the narrowest tile LPS x R (R in {4,8,12,16} grid points per lane under the run-length bound
// (R-1) h <= 5.5, LPS in {2,4,8} lanes per sample) that covers the typical window of 2 sqrt(T2 + 8)/h + 1 points
// (+ the spread of a chunk) in ONE pass; the widest admissible tile when none does.  Evaluated by one thread per
// unit, so it is a short table walk.  Returns false when windows cannot pay (window ~ whole grid, too few samples,
// grid too coarse for the recurrence).
// `gran`: chunk sizes are rounded up to a multiple of it (the fused kernel summarises 64-sample blocks while it
// reweights them, so its chunks must be unions of such blocks).
// `t2`: the window threshold in bits (terms below 2^-t2 of the largest term at a grid point are dropped).
__device__ __forceinline__ bool win_plan(int G, int n, float h, int iters, int max_chunks, WinPlan& pl, int gran = 1,
                                         float t2 = CHB_WIN_T2) {
  if (!(h > 0.f) || h > 1.8f || iters <= 0) return false;
  const int wn = 2 * (int)ceilf(sqrtf(t2 + 8.f) / h) + CHB_WIN_SPAN;    // half-width sqrt(t2 + log2(chunk)) in scaled units
  if (10 * wn > 7 * G) return false;
  const int rmax = min(CHB_WIN_MAXR, 1 + (int)(5.5f / h));             // (R-1) h <= 5.5
  // tiles by increasing width; at equal width the longer run (fewer MUFU per pair) comes first
  const unsigned char tr[12] = {4, 8, 4, 12, 16, 8, 4, 12, 16, 8, 12, 16};
  const unsigned char tl[12] = {2, 2, 4, 2, 2, 4, 8, 4, 4, 8, 8, 8};
  pl.R = 0; pl.LPS = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int r = tr[i], l = tl[i];
    if (r > rmax) continue;
    if (pl.R == 0 || pl.R * pl.LPS < wn) {                              // still too narrow: take the wider tile
      if (pl.R == 0 || r * l > pl.R * pl.LPS || (r * l == pl.R * pl.LPS && r > pl.R)) { pl.R = r; pl.LPS = l; }
    } else if (r * l == pl.R * pl.LPS && r > pl.R) { pl.R = r; pl.LPS = l; }
  }
  if (pl.R == 0) return false;
  // at most 32 chunks (phase B keeps one chunk per lane), each a multiple of 4 sub-stream rounds
  // Few samples (walker batches with ~1000 samples per event): shorter chunks rather than no windows at all -- the
  // direct sums cost one MUFU per pair on the whole grid, ~4x the windowed recurrence even with 4 loop iterations per pass.
  const int q = 4 * (32 / pl.LPS);
  for (int it = iters;; it >>= 1) {
    pl.chunk = max(it * (32 / pl.LPS), ((n + 31) / 32 + q - 1) / q * q);
    pl.chunk = (pl.chunk + gran - 1) / gran * gran;
    pl.nchunks = (n + pl.chunk - 1) / pl.chunk;
    if (pl.nchunks >= 8 || it <= 8) break;
  }
  (void)max_chunks;
  return pl.nchunks >= 8;
}

// Combine the 32/LPS sample sub-streams of a warp (lane bits >= LPS).  While the number of live values per lane
// is even the exchange is a reduce-scatter (each lane keeps one half and receives the partner's copy of it),
// so both the shuffle count and the values left to flush shrink geometrically; odd counts fall back to a
// butterfly and only the lane with the bit clear stays `owner`.  On return the lane holds rs_final<V,O>()
// sums for run offsets rbase .. rbase + that - 1.
template <int V, int O> struct RsFinal { static constexpr int value = (V % 2 == 0) ? RsFinal<V / 2, O * 2>::value : RsFinal<V, O * 2>::value; };
template <int V> struct RsFinal<V, 32> { static constexpr int value = V; };
template <int V> struct RsFinal<V, 64> { static constexpr int value = V; };
template <int V, int O, int N>
__device__ __forceinline__ void rs_reduce(float (&a)[N], int lane, int& rbase, bool& owner) {
  if constexpr (O < 32) {
    if constexpr (V % 2 == 0) {
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        const float keep = up ? a[i + V / 2] : a[i];
        const float send = up ? a[i] : a[i + V / 2];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      rbase += up ? V / 2 : 0;
      rs_reduce<V / 2, O * 2, N>(a, lane, rbase, owner);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], O);
      owner = owner && ((lane & O) == 0);
      rs_reduce<V, O * 2, N>(a, lane, rbase, owner);
    }
  }
}

// RAW: the stage holds UNSCALED samples {dz_a, dz_b, log2 w_a, log2 w_b}; the scale x' = dz * sf rides in the FFMA2 that
// forms x' - g, the weight normalisation log2(1/W) = koff in the constant added to the exponent -- no rescaling pass
// over the samples and no extra instruction in the loop.
template <int R, int LPS, bool GL, bool RAW = false>
__device__ __forceinline__ void kde_win_pass(const float2* __restrict__ xl, int cb, int ce, int gb, int glast,
                                             float4 sm4, double gfirst, double hd, float h,
                                             const float* __restrict__ cr, double* __restrict__ row,
                                             float sf = 1.f, float koff = 0.f) {
  static_assert(R % 4 == 0 && R <= CHB_WIN_MAXR, "run length");
  constexpr int S = 32 / LPS;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPS, sub = lane / LPS;
  const int g0 = gb + gl * R;
  const float run_lo = (float)(gfirst + (double)g0 * hd), run_hi = run_lo + (float)(R - 1) * h;
  const bool rev = run_hi < sm4.x;                       // run entirely left of the chunk: walk it right-to-left
  const int gs = rev ? g0 + R - 1 : g0;
  const float hs = rev ? -h : h;
  const float gp = (float)(gfirst + (double)gs * hd);
  const float dr = rev ? sm4.x - run_hi : fmaxf(run_lo - sm4.y, 0.f);
  const float top = sm4.z - dr * dr;                     // largest exponent any term of this run can have
  const float Kf = (top > -64.f) ? 0.f : fminf(floorf(100.f - top), 1900.f);
  const float m2h = -2.f * hs, mh2 = -h * h;
  // Samples are stored in pairs {x'_a, x'_b, lw_a, lw_b} (phase A): one LDS.128 brings two samples as two packed
  // operands, and the whole recurrence runs on FADD2 / FMUL2 / FFMA2 -- one issue slot for two samples.
  const float4* __restrict__ xp = reinterpret_cast<const float4*>(xl);
  const f32x2 gpn2 = pk2(-gp, -gp), mone2 = pk2(-1.f, -1.f), Kp2 = pk2(Kf + koff, Kf + koff);
  const f32x2 sf2 = pk2(sf, sf);
  const f32x2 p2h2 = pk2(-m2h, -m2h), mh22 = pk2(mh2, mh2);
  f32x2 acc2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc2[r] = 0ull;
#pragma unroll 2
  for (int jp = (cb >> 1) + sub; jp < (ce >> 1); jp += S) {
    const float4 v = GL ? __ldcg(xp + jp) : xp[jp];                    // GL: samples in global memory, read through L2
    const f32x2 nd = RAW ? fma2(pk2(v.x, v.y), sf2, gpn2) : add2(pk2(v.x, v.y), gpn2);   // x' - g = -d, both samples
    const f32x2 arg = add2(fma2(mul2(nd, nd), mone2, pk2(v.z, v.w)), Kp2);  // lw - d^2 + K
    const f32x2 qa = fma2(nd, p2h2, mh22);                             // -(2 hs d + h^2)
    float a0, a1, b0, b1;
    upk2(arg, a0, a1); upk2(qa, b0, b1);
    const f32x2 e0 = pk2(ex2_ftz(a0), ex2_ftz(a1));
    const f32x2 q = pk2(ex2_ftz(fminf(b0, 31.f)), ex2_ftz(fminf(b1, 31.f)));
    const f32x2 q2 = mul2(q, q), q3 = mul2(q2, q), q4 = mul2(q2, q2);
    f32x2 p = e0;
#pragma unroll
    for (int b = 0; b < R; b += 4) {
      if (b) p = mul2(p, q4);
      acc2[b] = add2(acc2[b], p);
      acc2[b + 1] = fma2(p, q, acc2[b + 1]);
      acc2[b + 2] = fma2(p, q2, acc2[b + 2]);
      acc2[b + 3] = fma2(p, q3, acc2[b + 3]);
    }
  }
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { float lo_, hi_; upk2(acc2[r], lo_, hi_); acc[r] = lo_ + hi_; }
  int rbase = 0;
  bool owner = true;
  rs_reduce<R, LPS, R>(acc, lane, rbase, owner);
  constexpr int VF = RsFinal<R, LPS>::value;
  // exact 2^-K as a double (|K| <= 1900 needs two factors)
  const int K = (int)Kf, K1 = K / 2, K2 = K - K1;
  const double sc = __hiloint2double((1023 - K1) << 20, 0) * __hiloint2double((1023 - K2) << 20, 0);
  const int step = rev ? -1 : 1;
#pragma unroll
  for (int i = 0; i < VF; ++i) {
    const int r = rbase + i;
    const int g = gs + step * r;
    if (owner && g <= glast) row[g] += (double)(acc[i] * cr[r]) * sc;
  }
}

// Phase C driver: the passes of all chunks form one list; lane c of every warp holds chunk c's window and the
// inclusive prefix sum of passes (pend), so a warp finds the chunk of list item `it` with one ballot.  Warp w
// takes the contiguous slice [w T/NW, (w+1) T/NW) of the list: the extra passes of the wide edge chunks are
// spread over the warps and the split is the same on every run (bit-reproducible sums).
template <int R, int LPS, int NW, bool GL, bool RAW = false>
__device__ __forceinline__ void kde_win_chunks(const float2* __restrict__ xl, int n, int G, double gfirst, double hd,
                                               const WinPlan& pl, const float4* __restrict__ summ, int2 w, int np,
                                               int pend, const float* __restrict__ cr, double* __restrict__ rows,
                                               float sf = 1.f, float koff = 0.f) {
  constexpr int W = LPS * R;
  const int warp = threadIdx.x >> 5;
  const float h = (float)hd;
  double* row = rows + warp * G;
  const int total = __shfl_sync(0xffffffffu, pend, 31);
  const int i0 = (int)(((long long)warp * total) / NW), i1 = (int)(((long long)(warp + 1) * total) / NW);
  for (int it = i0; it < i1; ++it) {
    const int c = __popc(__ballot_sync(0xffffffffu, pend <= it));      // first chunk whose prefix exceeds `it`
    const int wx = __shfl_sync(0xffffffffu, w.x, c), wy = __shfl_sync(0xffffffffu, w.y, c);
    const int first = __shfl_sync(0xffffffffu, pend - np, c);
    const int gb = wx + (it - first) * W;
    const int cb = c * pl.chunk, ce = min(n, cb + pl.chunk);
    kde_win_pass<R, LPS, GL, RAW>(xl, cb, ce, gb, wy, summ[c], gfirst, hd, h, cr, row, sf, koff);
    __syncwarp();
  }
}

// Phases B and C for chunk summaries that are already in `summ` (scaled units, weights normalised), rows zeroed and
// `cr` filled; all of them visible to the CTA (a barrier has passed).  Ends with dens[] written (no barrier after).
template <int NW, bool GL, bool RAW>
__device__ __forceinline__ void kde_win_BC(const float2* __restrict__ xw, int n, int G, double gfirst, double hd,
                                           const WinPlan& pl, double scale, const float4* __restrict__ summ,
                                           int2* __restrict__ win, const float* __restrict__ cr,
                                           double* __restrict__ rows, double* __restrict__ dens,
                                           float sf = 1.f, float koff = 0.f, float t2 = CHB_WIN_T2) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float h = (float)hd;
  // ---- phase B: lane c holds chunk c; every warp walks its share of the grid points: M(g) is one warp-wide max,
  // the need test one compare per lane, the hull of the needed points accumulates in registers ------------
  const float lgchunk = lg2f_((float)pl.chunk);
  // (measured: packing two grid points per step into the half-warps for events with <= 16 chunks, with a half-warp
  //  redux.sync, made C5 11 % SLOWER -- the partial-mask redux is not the one-instruction path)
  {
    const float4 my = (lane < pl.nchunks) ? summ[lane] : make_float4(INFINITY, -INFINITY, -INFINITY, 0.f);
    const float myU = my.z + lgchunk, gf = (float)gfirst;
    int gmin = G, gmax = -1;
    for (int g = warp; g < G; g += NW) {
      const float gp = fmaf((float)g, h, gf);
      const float d = gp - my.w;
      const float m = warp_max_f32(fmaf(-d, d, my.z));
      const float dist = fmaxf(fmaxf(my.x - gp, gp - my.y), 0.f);
      if (fmaf(-dist, dist, myU) >= m - t2 && my.z > -INFINITY) { gmin = min(gmin, g); gmax = g; }
    }
    if (gmax >= 0) { atomicMin(&win[lane].x, gmin); atomicMax(&win[lane].y, gmax); }
  }
  __syncthreads();
  // ---- phase C: pair sums ------------------------------------------------------------------------
  const int Wp = pl.LPS * pl.R;
  const int2 w = (lane < pl.nchunks) ? win[lane] : make_int2(G, -1);
  const int np = (w.x <= w.y) ? (w.y - w.x + Wp) / Wp : 0;
  int pend = np;                                                       // inclusive prefix sum of passes per chunk
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pend, o); if (lane >= o) pend += y; }
#define CHB_WIN_CASE(RR, LL) case RR * 16 + LL: kde_win_chunks<RR, LL, NW, GL, RAW>(xw, n, G, gfirst, hd, pl, summ, w, np, pend, cr, rows, sf, koff); break;
  switch (pl.R * 16 + pl.LPS) {
    CHB_WIN_CASE(4, 2) CHB_WIN_CASE(4, 4) CHB_WIN_CASE(4, 8)
    CHB_WIN_CASE(8, 2) CHB_WIN_CASE(8, 4) CHB_WIN_CASE(8, 8)
    CHB_WIN_CASE(12, 2) CHB_WIN_CASE(12, 4) CHB_WIN_CASE(12, 8)
    CHB_WIN_CASE(16, 2) CHB_WIN_CASE(16, 4) CHB_WIN_CASE(16, 8)
    default: break;
  }
#undef CHB_WIN_CASE
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += NW * 32) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += rows[w * G + g];
    dens[g] = acc * scale;
  }
}

// Whole-CTA call (NW warps).  xw: in {x, w} (x sorted or not, n EVEN), out pairs {x'_a, x'_b, log2 w'_a, log2 w'_b}.  Scratch: summ/win hold
// pl.nchunks (<= 32) entries, cr 16 floats, rows NW*G doubles.  dens[g] = scale * sum_j w'_j 2^-(g'_g - x'_j)^2 with
// g'_g = (lb + g step - c) sf, sf = float(s) shared by samples and grid.
template <int NW, bool GL = false>
__device__ __forceinline__ void kde1d_f32_win(float2* __restrict__ xw, int n, int G, double lb, double step, double c,
                                              double s, double W, const WinPlan& pl, double scale,
                                              float4* __restrict__ summ, int2* __restrict__ win, float* __restrict__ cr,
                                              double* __restrict__ rows, double* __restrict__ dens) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float sf = (float)s;
  const double gfirst = (lb - c) * (double)sf, hd = step * (double)sf;
  const float h = (float)hd;
  const float c_hi = (float)c, c_lo = (float)(c - (double)c_hi);
  const float lg2invW = -lg2f_((float)W);
  // ---- phase A: rescale + chunk summaries ----------------------------------------------------
  for (int i = threadIdx.x; i < NW * G; i += NW * 32) rows[i] = 0.0;
  if (threadIdx.x < 16) cr[threadIdx.x] = exp2f(-(h * h) * (float)(threadIdx.x * (threadIdx.x - 1)));
  for (int ck = warp; ck < pl.nchunks; ck += NW) {
    const int cb = ck * pl.chunk, ce = min(n, cb + pl.chunk);
    float lo = INFINITY, hi = -INFINITY, lm = -INFINITY, xm = 0.f;
    float4* __restrict__ xp = reinterpret_cast<float4*>(xw);
    for (int jp = (cb >> 1) + lane; jp < (ce >> 1); jp += 32) {        // two samples {z_a, w_a, z_b, w_b} per lane
      const float4 v = xp[jp];
      const float xa = ((v.x - c_hi) - c_lo) * sf, xb = ((v.z - c_hi) - c_lo) * sf;
      const bool la = v.y > 0.f, lb_ = v.w > 0.f;                      // zero / NaN weights add exactly 0
      const float lwa = la ? lg2f_(v.y) + lg2invW : -INFINITY, lwb = lb_ ? lg2f_(v.w) + lg2invW : -INFINITY;
      xp[jp] = make_float4(xa, xb, lwa, lwb);                          // pair layout of the packed pass loop
      lo = fminf(lo, fminf(la ? xa : INFINITY, lb_ ? xb : INFINITY));
      hi = fmaxf(hi, fmaxf(la ? xa : -INFINITY, lb_ ? xb : -INFINITY));
      if (lwa > lm) { lm = lwa; xm = xa; }
      if (lwb > lm) { lm = lwb; xm = xb; }
    }
    lo = warp_min_f32(lo); hi = warp_max_f32(hi);
    const float lmw = warp_max_f32(lm);
    const unsigned pick = __ballot_sync(0xffffffffu, lm == lmw && lm > -INFINITY);
    const float xmw = __shfl_sync(0xffffffffu, xm, pick ? (__ffs(pick) - 1) : 0);
    if (lane == 0) {
      summ[ck] = make_float4(lo, hi, pick ? lmw : -INFINITY, xmw);
      win[ck] = make_int2(G, -1);
    }
  }
  __syncthreads();
  kde_win_BC<NW, GL, false>(xw, n, G, gfirst, hd, pl, scale, summ, win, cr, rows, dens);
}
