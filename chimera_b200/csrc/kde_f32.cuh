// kde_f32.cuh -- fp32 KDE pair sums shared by the numerator kernels.
#pragma once
#include "models.cuh"

// 2^x on the MUFU pipe, flush-to-zero: no range fix-up code around the instruction (exp2f() adds an
// FSETP and two predicated FMULs per call to keep denormal results, which are irrelevant here).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int R, int NW>
__device__ __forceinline__ void kde1d_f32_pass(const float2* __restrict__ xw, int n, const double* __restrict__ eg,
                                               int G, int g_base, double c, double s, int kernel,
                                               float* __restrict__ part /* [NW][G] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gp[R], acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int g = g_base + r * 32 + lane;
    gp[r] = (g < G) ? (float)((eg[g] - c) * s) : 3.0e18f;   // far away: contributes exactly 0
    acc[r] = 0.f;
  }
  const int per = (n + NW - 1) / NW;
  const int j0 = min(n, warp * per), j1 = min(n, j0 + per);
  if (kernel == CHB_KERNEL_GAUSS) {
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float2 v = xw[j];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d = gp[r] - v.x;
        acc[r] = fmaf(v.y, ex2_ftz(-(d * d)), acc[r]);
      }
    }
  } else {
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float2 v = xw[j];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d = gp[r] - v.x;
        acc[r] = fmaf(v.y, fmaxf(fmaf(-d, d, 1.f), 0.f), acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int g = g_base + r * 32 + lane;
    if (g < G) part[warp * G + g] = acc[r];
  }
}

// Gaussian pair sums on a UNIFORM grid by recurrence.  Each lane owns R CONSECUTIVE grid points
// g0, g0+h, ..., g0+(R-1)h (scaled units).  With d = g0 - x':
//     E_0 = 2^-(d^2),   E_{r+1} = E_r q_r,   q_0 = 2^-(2 h d + h^2),   q_{r+1} = q_r c,   c = 2^-(2 h^2)
// which is the identity 2^-((d+(r+1)h)^2) = 2^-((d+rh)^2) 2^-(2h(d+rh)+h^2).
// Form used here, with 2 FP32 instructions per pair:
//     w E_r = [w E_0] q_0^r c^{r(r-1)/2}
// The constant c^{r(r-1)/2} = 2^-(h^2 r (r-1)) does not depend on the sample, so it is applied ONCE to the
// accumulator after the loop; inside the loop only p_r = p_{r-1} q_0 (FMUL) and acc_r += [w E_0] p_r (FFMA)
// remain.  The weight rides in the exponent: the data set holds {x', log2 w'} and w E_0 = 2^(log2 w' - d^2)
// costs FADD + FFMA + MUFU.  A warp is split into 32/LPS sample streams of LPS lanes; each lane owns R
// consecutive grid points (LPS*R >= G in one pass keeps every lane busy: G=150 -> LPS=16, R=10).
// Bounds: accumulated terms are E_r / c^{r(r-1)/2} <= 2^(h^2 R^2) (caller keeps (R-1) h <= 5.5);
// the exponent of q_0 is clamped to 120/(R-1) so p_r stays finite -- the clamp only acts where
// E_0 < 2^-110, i.e. on terms that are zero to fp32 anyway.
template <int R, int LPS, int NW>
__device__ __forceinline__ void kde1d_f32_rec2_pass(const float2* __restrict__ xl, int n, const double* __restrict__ eg,
                                                    int G, int g_base, double c, double s, float h,
                                                    float* __restrict__ part /* [NW*32/LPS][G] */) {
  constexpr int S = 32 / LPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPS, sub = lane / LPS;
  const int g0 = g_base + gl * R;
  const float gp = (g0 < G) ? (float)((eg[g0] - c) * s) : 3.0e18f;
  const float m2h = -2.f * h, mh2 = -h * h;
  const float qmax = 120.f / (float)(R > 1 ? R - 1 : 1);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const int per = (n + NW - 1) / NW;
  const int j0 = min(n, warp * per), j1 = min(n, j0 + per);
#pragma unroll 4
  for (int j = j0 + sub; j < j1; j += S) {
    const float2 v = xl[j];
    const float d = gp - v.x;
    const float e0 = ex2_ftz(fmaf(-d, d, v.y));                        // w' 2^-(d^2)
    const float q = ex2_ftz(fminf(fmaf(d, m2h, mh2), qmax));
    acc[0] += e0;
    float p = e0;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      p *= q;
      acc[r] += p;
    }
  }
  const float h2 = h * h;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int g = g0 + r;
    if (g < G) part[(warp * S + sub) * G + g] = acc[r] * exp2f(-h2 * (float)(r * (r - 1)));
  }
}

// dens[g] = scale * sum_j w'_j 2^-(g'-x'_j)^2 on a uniform grid; xl = {x', log2 w'}.  Needs room for
// NW*2 partial rows of G floats.  Returns false (nothing done) when no admissible tiling exists.
template <int NW>
__device__ __forceinline__ void kde1d_f32_rec2(const float2* __restrict__ xl, int n, const double* __restrict__ eg, int G,
                                               double c, double s, float h, int R, int LPS, double scale,
                                               float* __restrict__ part, double* __restrict__ dens) {
  const int S = 32 / LPS;
  for (int gb = 0; gb < G; gb += LPS * R) {
    if (LPS == 16) {
      switch (R) {
        case 6: kde1d_f32_rec2_pass<6, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 7: kde1d_f32_rec2_pass<7, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 8: kde1d_f32_rec2_pass<8, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 9: kde1d_f32_rec2_pass<9, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 10: kde1d_f32_rec2_pass<10, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 11: kde1d_f32_rec2_pass<11, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        default: kde1d_f32_rec2_pass<12, 16, NW>(xl, n, eg, G, gb, c, s, h, part); break;
      }
    } else {
      switch (R) {
        case 2: kde1d_f32_rec2_pass<2, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 3: kde1d_f32_rec2_pass<3, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 4: kde1d_f32_rec2_pass<4, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 5: kde1d_f32_rec2_pass<5, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 6: kde1d_f32_rec2_pass<6, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        case 7: kde1d_f32_rec2_pass<7, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
        default: kde1d_f32_rec2_pass<8, 32, NW>(xl, n, eg, G, gb, c, s, h, part); break;
      }
    }
  }
  __syncthreads();
  const int rows = NW * S;
  for (int g = threadIdx.x; g < G; g += NW * 32) {
    double acc = 0.0;
    for (int w = 0; w < rows; ++w) acc += (double)part[w * G + g];
    dens[g] = acc * scale;
  }
}

// tiling choice for kde1d_f32_rec2: minimise instructions per sample, (R-1) h <= 5.5, partial rows fit
__device__ __forceinline__ bool rec2_choose(int G, float h, int part_floats, int NW, int& R, int& LPS) {
  float best = 1e30f;
  R = 0; LPS = 0;
  for (int lps = 16; lps <= 32; lps += 16) {
    const int S = 32 / lps;
    if (NW * S * G > part_floats) continue;
    const int rlo = (lps == 16) ? 6 : 2, rhi = (lps == 16) ? 12 : 8;
    for (int r = rlo; r <= rhi; ++r) {
      if ((float)(r - 1) * h > 5.5f) continue;
      const int passes = (G + lps * r - 1) / (lps * r);
      const float cost = (float)passes * (6.f + 2.f * (float)r) / (float)S;
      if (cost < best) { best = cost; R = r; LPS = lps; }
    }
  }
  return R > 0;
}

// dens[g] = scale * sum_j w'_j K(g' - x'_j).  Must be called by the whole CTA (NW warps).
template <int NW>
__device__ __forceinline__ void kde1d_f32(const float2* __restrict__ xw, int n, const double* __restrict__ eg, int G,
                                          double c, double s, int kernel, double scale, float* __restrict__ part,
                                          double* __restrict__ dens) {
  // register tile height: fewest (passes x R), larger R on ties
  int R = 1, best = 1 << 30;
  for (int r = 8; r >= 1; --r) {
    int cost = ((G + 32 * r - 1) / (32 * r)) * r;
    if (cost < best) { best = cost; R = r; }
  }
  for (int gb = 0; gb < G; gb += 32 * R) {
    switch (R) {
      case 1: kde1d_f32_pass<1, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 2: kde1d_f32_pass<2, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 3: kde1d_f32_pass<3, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 4: kde1d_f32_pass<4, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 5: kde1d_f32_pass<5, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 6: kde1d_f32_pass<6, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      case 7: kde1d_f32_pass<7, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
      default: kde1d_f32_pass<8, NW>(xw, n, eg, G, gb, c, s, kernel, part); break;
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += NW * 32) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += (double)part[w * G + g];
    dens[g] = acc * scale;
  }
}

