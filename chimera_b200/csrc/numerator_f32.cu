// numerator_f32.cu -- fast path of the fused per-(event, hyper-point) numerator (CHB_FP32 mode).
//
// Same stages and semantics as numerator.cu (stage 1 reweighting, stage 2 KDE, stage 3 z-integral;
// likelihood.py:105-301), organised for throughput on B200:
//   * 256-thread CTAs with <= 113 KB of shared memory each, so TWO CTAs share an SM and the
//     latency-bound phases of one (table staging, reweighting, statistics) overlap the MUFU-bound
//     KDE pair sums of the other;
//   * samples staged as float2 {z, w} (8 B) and rescaled in place for the pair sums;
//   * only the tables the reweighting needs are staged (dl4 | cd4 | lut, 42 KB, one TMA bulk copy);
//     the 300 z-grid look-ups read zi4 from L2;
//   * one fused block reduction for all sample statistics (single pass, fp64 accumulation);
//   * z-grid terms (E, dVc/dz, ddL/dz, psi) in fp32 on the MUFU/FP32 pipes, one division per grid
//     point; integration weights folded into ck[k] = psi/(1+z) * trapz_w / jacobian.
// fp64 is kept for: table construction, sample statistics, bandwidth, centring/scaling of samples
// and grid points before the fp32 cast, cross-warp accumulation, interpolation, the z-integral and
// the final reduction.
#include "common.cuh"
#include "models_f32.cuh"
#include "kde_f32.cuh"
#include "kde_win.cuh"
#include "stage.cuh"
#include <algorithm>

#define F_NT 256
#define F_NW (F_NT / 32)
#ifndef F_NT_FULL2
#define F_NT_FULL2 512            // 'full', MODE 2 (one CTA per SM: 120 KB of staged samples): 16 warps instead of 8
#endif
static inline int f32_threads(int kind, int mode) { return (kind == CHB_PGW_FULL && mode == 2) ? F_NT_FULL2 : F_NT; }
// tuning knobs of the split fast path (build.py passes -D overrides from CHB_BUILD_DEFS for A/B runs)
#ifndef CHB_K1_MINB
#define CHB_K1_MINB 3        // co-resident CTAs per SM the reweighting kernel (MODE 1) is compiled for
#endif
#ifndef CHB_K1_U
#define CHB_K1_U 2           // samples in flight per thread in the reweighting loop (x2 with the prefetch)
#endif
#ifndef CHB_K2_MINB
#define CHB_K2_MINB 3        // co-resident CTAs per SM of the KDE/z-integral kernel (MODE 2)
#endif

struct FPlan {
  int tab, zgrid, dV, ck, pgw, eg, dens, bc, bs, xwb, part, red, stage, total;
};
// mode 0: fused kernel; 1: reweighting only (table + reduction scratch); 2: KDE + z-integral on staged samples (no table)
__host__ __device__ inline FPlan make_fplan(int tab_doubles, int Nz, int B, int Ns, int kind, int mode = 0, int nw = F_NW) {
  FPlan p;
  int o = 0;
  p.tab = o; o += (mode >= 2) ? 0 : tab_doubles;
  if (mode == 1) {
    p.zgrid = p.dV = p.ck = p.eg = p.pgw = p.dens = p.bc = p.bs = p.xwb = p.part = o;
    p.red = o; o += 64;
    p.stage = o; p.total = o;
    return p;
  }
  p.zgrid = o; o += Nz;
  p.dV = o; o += Nz;
  p.ck = o; o += Nz;
  p.eg = o; o += Nz;
  p.pgw = o; o += Nz;          // pgw | dens are contiguous: the windowed KDE keeps its chunk tables there
  p.dens = o; o += Nz;
  p.bc = o; o += B;
  p.bs = o; o += B;
  p.xwb = o; o += B;
  p.part = o; o += (nw * Nz + 1) / 2;
  p.red = o; o += (nw > 8) ? 128 : 64;      // block_stats: 6 doubles per warp
  o = (o + 1) & ~1;
  p.stage = o; o += (kind == CHB_PGW_FULL ? 3 : 1) * Ns;     // float2 {z,w} [+ float4 whitened]
  p.total = o;
  return p;
}
static inline int f32_tab_doubles(const TableLayout& lay) { return lay.f32_core() - lay.f32_dl4(); }

size_t numerator_f32_smem_bytes(const NumArgs& a, int mode) {
  return (size_t)make_fplan(f32_tab_doubles(a.mc.lay), a.Nz, a.binning ? a.num_bins : 0, a.Ns, a.kind, mode, f32_threads(a.kind, mode) / 32).total * sizeof(double);
}

__device__ __forceinline__ double nan_to_num_log_f(double like) {
  double l = log(like);                  // likelihood.py:296-297
  if (isnan(l)) return -INFINITY;
  if (isinf(l)) return l > 0 ? CHB_DBL_MAX : -CHB_DBL_MAX;
  return l;
}

// sum of 4 doubles + min/max of a float over the CTA in one barrier pair; result in all threads
struct Stats6 { double a, b, c, d; float mn, mx; };
template <int NW>
__device__ __forceinline__ Stats6 block_stats(Stats6 v, double* red /* >= 6*NW doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.a += __shfl_xor_sync(0xffffffffu, v.a, o);
    v.b += __shfl_xor_sync(0xffffffffu, v.b, o);
    v.c += __shfl_xor_sync(0xffffffffu, v.c, o);
    v.d += __shfl_xor_sync(0xffffffffu, v.d, o);
    v.mn = fminf(v.mn, __shfl_xor_sync(0xffffffffu, v.mn, o));
    v.mx = fmaxf(v.mx, __shfl_xor_sync(0xffffffffu, v.mx, o));
  }
  __syncthreads();
  if (lane == 0) {
    red[w * 6 + 0] = v.a; red[w * 6 + 1] = v.b; red[w * 6 + 2] = v.c; red[w * 6 + 3] = v.d;
    red[w * 6 + 4] = (double)v.mn; red[w * 6 + 5] = (double)v.mx;
  }
  __syncthreads();
  Stats6 r = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    r.a += red[i * 6 + 0]; r.b += red[i * 6 + 1]; r.c += red[i * 6 + 2]; r.d += red[i * 6 + 3];
    r.mn = fminf(r.mn, (float)red[i * 6 + 4]); r.mx = fmaxf(r.mx, (float)red[i * 6 + 5]);
  }
  return r;
}

// rescale a float2 {x, w} data set in place to {(x - c) s, w / W} and run the fp32 pair sums
// `ustep` > 0: eg is a uniform grid with that spacing -> Gaussian sums by recurrence when the scaled
// spacing allows it (kde_f32.cuh).
template <int NT>
__device__ __forceinline__ void kde_inplace(float2* xw, int n, const double* __restrict__ eg, int G, double bw, double W,
                                            int kernel, double scale_pdf, float* part, int part_floats, double* dens,
                                            double ustep = 0.0) {
  const double c = 0.5 * (eg[0] + eg[G - 1]);
  const double s = (kernel == CHB_KERNEL_GAUSS) ? 0.8493218002880191 / bw : 1.0 / bw;   // sqrt(log2(e)/2)
  const double knorm = (kernel == CHB_KERNEL_GAUSS) ? 0.3989422804014327 : 0.75;
  const double invW = 1.0 / W;
  int R = 0, LPS = 0;
  const float h = (float)(ustep * s);
  const bool rec = (kernel == CHB_KERNEL_GAUSS) && ustep > 0.0 && G >= 2 && rec2_choose(G, h, part_floats, NT / 32, R, LPS);
  if (rec) {
    for (int j = threadIdx.x; j < n; j += NT) {
      const float2 v = xw[j];
      xw[j] = make_float2((float)(((double)v.x - c) * s), lg2f_((float)((double)v.y * invW)));
    }
    __syncthreads();
    kde1d_f32_rec2<NT / 32>(xw, n, eg, G, c, s, h, R, LPS, scale_pdf * knorm / bw, part, dens);
  } else {
    for (int j = threadIdx.x; j < n; j += NT) {
      const float2 v = xw[j];
      xw[j] = make_float2((float)(((double)v.x - c) * s), (float)((double)v.y * invW));
    }
    __syncthreads();
    kde1d_f32<NT / 32>(xw, n, eg, G, c, s, kernel, scale_pdf * knorm / bw, part, dens);
  }
}

// One thread per (hyper-point, event, k): full-occupancy evaluation of the z-grid terms, so that the
// persistent numerator CTAs only stream 8 B per grid point instead of running a latency-bound phase.
__global__ void __launch_bounds__(256)
zgrid_terms_kernel(const NumArgs a, int h0, int nh) {
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  const int h = h0 + blockIdx.y;
  if (threadIdx.x < CHB_NPAR) P[threadIdx.x] = a.hyper[(size_t)h * CHB_NPAR + threadIdx.x];
  if (threadIdx.x >= 64 && threadIdx.x < 64 + CHB_NHC) HC[threadIdx.x - 64] = a.HC[(size_t)h * CHB_NHC + threadIdx.x - 64];
  __syncthreads();
  const TableLayout lay = a.mc.lay;
  const double* tblk = a.tabs + (size_t)h * lay.total() + lay.off_f32();
  F32Consts fc = make_f32_consts(a.mc, P, HC, tblk);
  const CosmoRateF32 cr = make_cosmo_rate_f32(a.mc, P, HC);
  const long long n = (long long)a.Nev * a.Nz;
  float2* out = a.zterms_out + (size_t)blockIdx.y * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % a.Nz);
    const double z = a.zgrids[i];
    const double zl = (k > 0) ? a.zgrids[i - 1] : z, zr = (k < a.Nz - 1) ? a.zgrids[i + 1] : z;
    out[i] = zgrid_terms_f32(fc, cr, P, HC, a.mc.cosmo_model, z, 0.5 * (zr - zl));
  }
  (void)nh;
}
cudaError_t launch_zgrid_terms(const NumArgs& a, int h0, int nh, cudaStream_t s) {
  const long long n = (long long)a.Nev * a.Nz;
  // ~16 CTAs per SM over the whole launch: every CTA rebuilds its hyper-point's constants (fp64), so a thread should
  // amortise that over many grid points (C3: 256 hyper-points x 10 CTAs, ~120 points per thread; one point per thread
  // cost 1.06 ms of which ~0.9 ms was the per-CTA set-up)
  int gx = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (148LL * 16 + nh - 1) / nh));
  dim3 grid(gx, nh);
  zgrid_terms_kernel<<<grid, 256, 0, s>>>(a, h0, nh);
  return cudaGetLastError();
}

#define FPHASE(i) do { if (a.prof && tid == 0) { long long _t = clock64(); pacc[i] += (unsigned long long)(_t - tlast); tlast = _t; } } while (0)

// KG: kind group the kernel is compiled for (0: '1d'/'approximate', 1: 'marginalized', 2: 'full') -- one
// instantiation per group keeps the instruction footprint of the hot 1-D path small.
// MODE: 0 = fused (reweighting + KDE + z-integral per unit in one CTA pass); 1 = reweighting + statistics only,
// samples {z, w} and the unit statistics go to the stage buffers in global memory; 2 = KDE + z-integral on the
// staged samples of a unit (one TMA bulk copy into shared memory).  Splitting lets either half run with three
// co-resident CTAs per SM (no 42 KB table block next to the 40 KB sample stage) -- see DESIGN.md section 4.
template <int KG, int MODE, int NT>
// ('full', MODE 2: one CTA per SM whatever the register count -- 120 KB of samples in shared memory -- so the pair loop
// gets the registers to keep several samples in flight)
__global__ void __launch_bounds__(NT, MODE == 0 ? 2 : (MODE == 1 ? CHB_K1_MINB : (KG == 2 ? 1 : CHB_K2_MINB)))
numerator_f32_kernel(const NumArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ double L[8];
  __shared__ float crs[20];
  __shared__ double sh_bw;
  __shared__ WinPlan sh_wp;
  __shared__ int sh_win;

  const TableLayout lay = a.mc.lay;
  const int Ns = a.Ns, Nz = a.Nz, Pp = a.P, B = a.binning ? a.num_bins : 0;
  const int tabd = lay.f32_core() - lay.f32_dl4();
  const FPlan pl = make_fplan(tabd, Nz, B, Ns, a.kind, MODE, NW);
  double* tab = sm + pl.tab;
  double* zgrid = sm + pl.zgrid;
  double* dV = sm + pl.dV;
  double* ck = sm + pl.ck;
  double* pgw = sm + pl.pgw;
  double* eg = sm + pl.eg;
  double* dens = sm + pl.dens;
  double* bc = sm + pl.bc;
  double* bs = sm + pl.bs;
  float2* xwb = reinterpret_cast<float2*>(sm + pl.xwb);
  float* part = reinterpret_cast<float*>(sm + pl.part);
  double* red = sm + pl.red;
  float2* zw_s = reinterpret_cast<float2*>(sm + pl.stage);
  float4* yw = reinterpret_cast<float4*>(sm + pl.stage + Ns);      // 'full' only

  const int cm = a.mc.cosmo_model;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tab_bytes = (uint32_t)(tabd * sizeof(double));
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
    // every thread orders its generic-proxy writes to the staging buffers of the previous unit before the barrier;
    // after it one thread may hand the buffers to the async proxy (TMA bulk copy) again
    fence_proxy_async();
    __syncthreads();
    const double* tblk = a.tabs + (size_t)h * lay.total() + lay.off_f32();
    float2* zw = (MODE == 1) ? a.zw_stage + (size_t)unit * Ns : zw_s;
    if (tid == 0) {
      if (MODE != 2) {
        mbar_expect_tx(&bar, tab_bytes);
        bulk_g2s(tab, tblk + lay.f32_dl4(), tab_bytes, &bar);
      } else {
        mbar_expect_tx(&bar, (uint32_t)(Ns * sizeof(float2)));
        bulk_g2s(zw_s, a.zw_stage + (size_t)unit * Ns, (uint32_t)(Ns * sizeof(float2)), &bar);
      }
    }
    if (tid < CHB_NPAR) P[tid] = a.hyper[(size_t)h * CHB_NPAR + tid];
    if (tid >= 64 && tid < 64 + CHB_NHC) HC[tid - 64] = a.HC[(size_t)h * CHB_NHC + tid - 64];
    if (MODE != 1) {
      const double* zgr = a.zgrids + (size_t)ev * Nz;
      for (int k = tid; k < Nz; k += NT) zgrid[k] = zgr[k];
    }
    __syncthreads();

    // constants of the fast path; table pointers: dl4|cd4|lut in shared memory (global in MODE 2), zi4 in L2
    F32Consts fc = make_f32_consts(a.mc, P, HC, (MODE >= 2) ? tblk : tab - lay.f32_dl4());
    fc.zi4 = reinterpret_cast<const float4*>(tblk + lay.f32_zi4());
    const CosmoRateF32 cr = make_cosmo_rate_f32(a.mc, P, HC);
    const float z_top = (float)P[CHB_P_ZMAX];
    // (fc.cd4_last is filled after the table copy has landed)

    // ---- z-grid terms: precomputed for all (hyper-point, event, k) by zgrid_terms_kernel when the
    // buffer fits (a.zterms), otherwise evaluated here --------------------------------------------
    if (MODE != 1) {
      if (a.zterms) {
        const float2* zt = a.zterms + ((size_t)(h - a.zterms_h0) * a.Nev + ev) * Nz;
        for (int k = tid; k < Nz; k += NT) { const float2 v = __ldg(zt + k); dV[k] = (double)v.x; ck[k] = (double)v.y; }
      } else {
        for (int k = tid; k < Nz; k += NT) {
          const double z = zgrid[k];
          const double zl = (k > 0) ? zgrid[k - 1] : z, zr = (k < Nz - 1) ? zgrid[k + 1] : z;
          const float2 v = zgrid_terms_f32(fc, cr, P, HC, cm, z, 0.5 * (zr - zl));
          dV[k] = (double)v.x; ck[k] = (double)v.y;
        }
      }
    }
    // MODE 2: the unit statistics written by the reweighting kernel are requested BEFORE waiting for the sample
    // copy, so their L2 latency hides behind the TMA instead of sitting in front of the next barrier
    double s1 = 0.0, s2 = 0.0, zmn = 0.0, zmx = 0.0, zstd = 0.0;
    if (MODE >= 2) {
      const double2* us = reinterpret_cast<const double2*>(a.unit_stats + (size_t)unit * 8);
      const double2 u0 = __ldg(us), u1 = __ldg(us + 1);
      s1 = u0.x; s2 = u0.y; zmn = u1.x; zmx = u1.y; zstd = __ldg(a.unit_stats + (size_t)unit * 8 + 4);
    }
    FPHASE(1);
    mbar_wait(&bar, phase);
    phase ^= 1;
    FPHASE(0);

    const size_t so = (size_t)ev * Ns;
    if (MODE < 2) {
      fc.cd4_last = fc.cd4[fc.rm - 1];
      // ---- stage 1: reweighting (pop_wrapper.py:67-80) ------------------------------------------
      Stats6 st = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
      // per-thread partial statistics in fp32 (a thread sees ~Ns/256 samples), z shifted by the redshift of the
      // event's median-dL sample so that the one-pass variance does not cancel; fp64 from the block reduction on
      float fa = 0.f, fb = 0.f, fcs = 0.f, fd = 0.f;
      const float z0 = z_from_dL_f32(fc, __ldg(&a.s4[(size_t)ev * Ns + Ns / 2].x));
      {
        // four samples in flight per thread, identical instruction stream for all of them (branch-free
        // weights, fixed two-step table scan) so the compiler interleaves their dependency chains
        const float4* s4 = a.s4 + so;
        const float2* l2 = a.l2 + so;
        constexpr int U = (MODE == 1) ? CHB_K1_U : 2;
        // software pipeline: the packed samples of the NEXT chunk are requested before the current chunk is
        // evaluated, so the L2 latency of the loads overlaps ~350 instructions of arithmetic
        float4 sv[U], nsv[U]; float2 lv[U], nlv[U];
        // static round-robin of U*32-sample chunks over the warps: every thread sees the same samples on every
        // run, so the fp32 partial statistics (and with them every per-event value) are bit-reproducible
        int cbase = warp * (U * 32);
        if (cbase < Ns) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = min(cbase + lane + u * 32, Ns - 1);   // tail lanes recompute the last sample, never stored
            sv[u] = __ldg(s4 + j);
            lv[u] = __ldg(l2 + j);
          }
        }
        while (cbase < Ns) {
          const int nbase = cbase + NW * (U * 32);
          if (nbase < Ns) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int j = min(nbase + lane + u * 32, Ns - 1);
              nsv[u] = __ldg(s4 + j);
              nlv[u] = __ldg(l2 + j);
            }
          }
          const int jb = cbase + lane;
          float zf[U], wf[U];
          int kk[U]; float4 ee[U]; bool more[U];
#pragma unroll
          for (int u = 0; u < U; ++u) zf[u] = z_lookup2(fc, sv[u].x, z_top, kk[u], ee[u], more[u]);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (more[u]) {                                      // rare: keep scanning
              int k = kk[u]; float4 e = ee[u];
              while (sv[u].x >= e.w && k < fc.rc - 2) { ++k; e = fc.dl4[k]; }
              float z = fmaf(sv[u].x - e.x, e.z, e.y);
              zf[u] = (sv[u].x >= e.w) ? z_top : z;
            }
          }
          float m1s[U], m2s[U], lzs[U];
          bool smooth = false;                                  // does any sample of the chunk sit on the low-mass taper?
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float opz = 1.f + zf[u];
            const float r = rcpf_(opz);
            lzs[u] = lg2f_(opz);
            m1s[u] = sv[u].y * r; m2s[u] = sv[u].z * r;
            smooth |= !(m2s[u] - fc.lo > fc.dm) || !(m1s[u] - fc.lo > fc.dm);
          }
          if (__any_sync(0xffffffffu, smooth)) {
#pragma unroll
            for (int u = 0; u < U; ++u) wf[u] = weight_bf<true>(fc, m1s[u], m2s[u], lv[u].x - lzs[u], lv[u].y - lzs[u], sv[u].w);
          } else {                                               // smoothing == 1 for every lane: skip its 6 MUFU per sample
#pragma unroll
            for (int u = 0; u < U; ++u) wf[u] = weight_bf<false>(fc, m1s[u], m2s[u], lv[u].x - lzs[u], lv[u].y - lzs[u], sv[u].w);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = jb + u * 32;
            if (j < Ns) {
              zw[j] = make_float2(zf[u], wf[u]);
              const float dz = zf[u] - z0;
              fa += wf[u]; fb = fmaf(wf[u], wf[u], fb); fcs += dz; fd = fmaf(dz, dz, fd);
              st.mn = fminf(st.mn, zf[u]); st.mx = fmaxf(st.mx, zf[u]);
            }
          }
          cbase = nbase;
#pragma unroll
          for (int u = 0; u < U; ++u) { sv[u] = nsv[u]; lv[u] = nlv[u]; }
        }
      }
      FPHASE(2);
      st.a = (double)fa; st.b = (double)fb; st.c = (double)fcs; st.d = (double)fd;
      st = block_stats<NW>(st, red);
      s1 = st.a; s2 = st.b;
      zmn = (double)st.mn; zmx = (double)st.mx;
      const double dzmean = st.c / Ns;
      zstd = sqrt(fmax(st.d / Ns - dzmean * dzmean, 0.0));    // one-pass variance about z0
      if (MODE == 1) {
        if (tid == 0) {
          double* us = a.unit_stats + (size_t)unit * 8;
          us[0] = s1; us[1] = s2; us[2] = zmn; us[3] = zmx; us[4] = zstd;
        }
        continue;
      }
    }
    const double norm = s1 / Ns;                  // likelihood.py:111
    const double neff = s1 * s1 / s2;             // likelihood.py:112
    const bool ok = (a.kind == CHB_PGW_FULL) ? !(neff < a.pe_neff) : (neff >= a.pe_neff);

    double* pout = nullptr;
    if (a.p_gw_out) pout = a.p_gw_out + ((size_t)h * a.Nev + ev) * (size_t)(pixelated ? Pp : 1) * Nz;
    const int npix = pixelated ? a.neff_pix[ev] : 1;

    if (!ok) {
      if (pout) for (int i = tid; i < (pixelated ? Pp : 1) * Nz; i += NT) pout[i] = 0.0;
      if (tid == 0) { a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_f(0.0); a.like_raw[(size_t)h * a.Nev + ev] = 0.0; }
      continue;
    }

    // ---- effective grid (likelihood.py:115-123 / 186-190) ------------------------------------
    int G = Nz;
    double ustep = 0.0;       // spacing of the effective grid when it is our own linspace
    if (a.kind != CHB_PGW_FULL) {
      if (a.use_cut) {
        G = Nz / 2;
        const double lb = (zmn - a.cut_grid * zstd > 0.0) ? zmn - a.cut_grid * zstd : 1.e-8;
        const double ub = zmx + a.cut_grid * zstd;
        const double step = (ub - lb) / (double)(G - 1);
        ustep = a.rec_off ? 0.0 : step;
        for (int i = tid; i < G; i += NT) eg[i] = (i == G - 1) ? ub : __dadd_rn(__dmul_rn((double)i, step), lb);
      } else {
        for (int i = tid; i < G; i += NT) eg[i] = zgrid[i];
      }
    }
    if (KG == 0 && tid == 0) {
      // bandwidth (utils/math.py:62-70) and the tiling of the windowed KDE, once per unit
      sh_win = 0;
      if (!a.binning) {
        // neff^(-1/5) on the MUFU pipe (2e-7 relative): the other 255 threads wait on this value
        const double neff_k0 = 1.0 / (s2 / (s1 * s1));
        double bw0;
        if (a.bw_method == CHB_BW_SCOTT) bw0 = (double)ex2f_(-0.2f * lg2f_((float)neff_k0)) * zstd;
        else if (a.bw_method == CHB_BW_SILVERMAN) bw0 = (double)ex2f_(-0.2f * lg2f_((float)(neff_k0 * 3.0 / 4.0))) * zstd;
        else bw0 = a.bw_value * zstd;
        sh_bw = bw0;
        // the chunk tables (32 x {float4, int2} = 768 B) live in the pgw | dens arrays: 2 Nz doubles
        if (a.kernel == CHB_KERNEL_GAUSS && ustep > 0.0 && a.kde_win_iters > 0 && G >= 2 && Nz >= 64 && (Ns & 1) == 0) {
          const double s0 = 0.8493218002880191 / bw0;
          WinPlan wp;
          if (win_plan(G, Ns, (float)(ustep * (double)(float)s0), a.kde_win_iters, 32, wp)) { sh_wp = wp; sh_win = 1; }
        }
      }
    }
    __syncthreads();
    FPHASE(3);

    double like_acc = 0.0;
    const double fR = HC[HC_FR];
    const double* pcat_ev = has_cat ? a.p_cat + (size_t)ev * Pp * Nz : nullptr;
    const double* pcompl_ev = has_cat ? a.P_compl + (size_t)ev * Nz : nullptr;

    if (KG == 0) {
      float2* dxw = zw;
      int dn = Ns;
      double W = s1, Q = s2, dstd = zstd;
      if (a.binning) {                           // utils/math.py:32-46
        const double step = (zmx - zmn) / (double)B;
        for (int i = tid; i < B; i += NT) {
          double e0 = __dadd_rn(__dmul_rn((double)i, step), zmn);
          double e1 = (i + 1 == B) ? zmx : __dadd_rn(__dmul_rn((double)(i + 1), step), zmn);
          bc[i] = (e0 + e1) / 2;
          bs[i] = 0.0;
        }
        __syncthreads();
        if (a.bin_runs) {
          // The samples are sorted by dL, hence by z: a bin is a contiguous run of samples.  Every thread walks a
          // contiguous block, sums each run in a register and touches shared memory once per run (2-3 atomics per
          // thread instead of one contended fp64 atomic per sample).  Correct for any order; only fast when sorted.
          const int per = (Ns + NT - 1) / NT;
          const int ja = min(Ns, tid * per), jb = min(Ns, ja + per);
          const double invB = (double)B / (zmx - zmn);
          int cur = -1;
          double run = 0.0;
          for (int j = ja; j < jb; ++j) {
            const float2 v = zw[j];
            const double f = floor(((double)v.x - zmn) * invB);
            if (isnan(f)) continue;
            const int b = (int)fmin(fmax(f, 0.0), (double)(B - 1));
            if (b != cur) {
              if (cur >= 0) atomicAdd(&bs[cur], run);
              cur = b; run = 0.0;
            }
            run += (double)v.y;
          }
          if (cur >= 0) atomicAdd(&bs[cur], run);
        } else {
          for (int j = tid; j < Ns; j += NT) {
            const float2 v = zw[j];
            double f = floor(((double)v.x - zmn) / (zmx - zmn) * B);
            if (!isnan(f)) atomicAdd(&bs[(int)fmin(fmax(f, 0.0), (double)(B - 1))], (double)v.y);
          }
        }
        __syncthreads();
        Stats6 t = {0.0, 0.0, 0.0, 0.0, 0.f, 0.f};
        for (int i = tid; i < B; i += NT) { t.a += bs[i]; t.b += bs[i] * bs[i]; t.c += bc[i]; t.d += bc[i] * bc[i]; }
        t = block_stats<NW>(t, red);
        W = t.a; Q = t.b;
        const double cmean = t.c / B;
        dstd = sqrt(fmax(t.d / B - cmean * cmean, 0.0));
        for (int i = tid; i < B; i += NT) xwb[i] = make_float2((float)bc[i], (float)bs[i]);
        // (bin centres are O(1) numbers: the float cast before centring costs 6e-8 relative, like z itself)
        dxw = xwb; dn = B;
        __syncthreads();
      }
      double bw;
      if (!a.binning) bw = sh_bw;
      else {
        const double neff_k = 1.0 / (Q / (W * W));
        if (a.bw_method == CHB_BW_SCOTT) bw = pow(neff_k, -0.2) * dstd;
        else if (a.bw_method == CHB_BW_SILVERMAN) bw = pow(neff_k * 3.0 / 4.0, -0.2) * dstd;
        else bw = a.bw_value * dstd;
      }
      const bool windowed = sh_win != 0;
      if (windowed) {
        // windowed recurrence over the sorted samples (kde_win.cuh); chunk tables live in pgw
        const WinPlan wp = sh_wp;
        float4* summ = reinterpret_cast<float4*>(pgw);
        int2* win = reinterpret_cast<int2*>(summ + 32);
        kde1d_f32_win<NW, false>(dxw, dn, G, eg[0], ustep, 0.5 * (eg[0] + eg[G - 1]), 0.8493218002880191 / bw, W, wp,
                            norm * 0.3989422804014327 / bw, summ, win, crs, reinterpret_cast<double*>(part), dens);
      } else {
        kde_inplace<NT>(dxw, dn, eg, G, bw, W, a.kernel, norm, part, NW * Nz, dens, ustep);
      }
      __syncthreads();
      // p_gw on the event grid (likelihood.py:139-141).  When nothing else needs the array, the interpolated value
      // goes straight into the integrand (same k): no staging, no extra barrier.
      const bool fused_tail = !pout && (a.kind == CHB_PGW_1D || a.catA != nullptr);
      const double inv_ustep = (ustep > 0.0) ? 1.0 / ustep : 0.0;
      auto pgw_at = [&](int k) -> double {
        const double x = zgrid[k];
        if (ustep > 0.0) {                        // uniform effective grid: index directly, then settle on the knots
          if (x < eg[0] || x > eg[G - 1]) return 0.0;
          int i = (int)((x - eg[0]) * inv_ustep);
          i = max(0, min(i, G - 2));
          while (i < G - 2 && x >= eg[i + 1]) ++i;
          while (i > 0 && x < eg[i]) --i;
          const double x0 = eg[i], dx = eg[i + 1] - x0, f0 = dens[i], df = dens[i + 1] - f0;
          return (fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((x - x0) / dx) * df;
        }
        return interp_lr(x, eg, dens, G, 0.0, 0.0);
      };
      if (fused_tail) {
        if (a.kind == CHB_PGW_1D) {
          for (int k = tid; k < Nz; k += NT) like_acc += pgw_at(k) * dV[k] * ck[k];
        } else {
          const double* A = a.catA + (size_t)ev * Nz;
          const double* Bk = a.catB + (size_t)ev * Nz;
          for (int k = tid; k < Nz; k += NT) {
            const double pgs = has_cat ? fR * A[k] + (1.0 - pcompl_ev[k]) * dV[k] * Bk[k] : dV[k] * Bk[k];
            like_acc += pgw_at(k) * pgs * ck[k];
          }
        }
      } else {
      for (int k = tid; k < Nz; k += NT) pgw[k] = pgw_at(k);
      __syncthreads();
      if (a.kind == CHB_PGW_1D) {
        for (int k = tid; k < Nz; k += NT) {
          like_acc += pgw[k] * dV[k] * ck[k];
          if (pout) pout[k] = pgw[k];
        }
      } else {
        const double* gwp = a.gw_pdf + (size_t)ev * Pp;
        if (pout) for (int i = tid; i < Pp * Nz; i += NT) pout[i] = pgw[i % Nz] * gwp[i / Nz];
        if (a.catA) {
          const double* A = a.catA + (size_t)ev * Nz;
          const double* Bk = a.catB + (size_t)ev * Nz;
          for (int k = tid; k < Nz; k += NT) {
            const double pgs = has_cat ? fR * A[k] + (1.0 - pcompl_ev[k]) * dV[k] * Bk[k] : dV[k] * Bk[k];
            like_acc += pgw[k] * pgs * ck[k];
          }
        } else {
          for (int i = tid; i < npix * Nz; i += NT) {
            const int p = i / Nz, k = i - p * Nz;
            const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
            if (pc == -100.0) continue;
            const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
            like_acc += (pgw[k] * gwp[p]) * pgal * ck[k];
          }
        }
      }
      }
    } else if (KG == 1) {
      // ---- p_gw3dmarg: per pixel, always Epanechnikov (likelihood.py:160-205) -----------------
      const int* off = a.pix_off + (size_t)ev * (Pp + 2);
      const double* gwp = a.gw_pdf + (size_t)ev * Pp;
      if (pout) for (int i = tid; i < Pp * Nz; i += NT) pout[i] = 0.0;
      for (int p = 0; p < npix; ++p) {
        const int o0 = off[p], o1 = off[p + 1], nin = o1 - o0;
        Stats6 t = {0.0, 0.0, 0.0, 0.0, INFINITY, -INFINITY};
        for (int j = o0 + tid; j < o1; j += NT) {
          const float2 v = zw[j];
          const double z = (double)v.x, w = (double)v.y;
          t.a += w; t.b += w * w; t.c += z; t.d += z * z; t.mx = fmaxf(t.mx, v.x);
        }
        t = block_stats<NW>(t, red);
        double W = t.a, Q = t.b;
        const double zmax_in = fmax((double)t.mx, zmn);
        float2* dxw = zw + o0;
        int dn = nin;
        double dstd;
        if (a.binning) {
          const double step = (zmax_in - zmn) / (double)B;
          for (int i = tid; i < B; i += NT) {
            double e0 = __dadd_rn(__dmul_rn((double)i, step), zmn);
            double e1 = (i + 1 == B) ? zmax_in : __dadd_rn(__dmul_rn((double)(i + 1), step), zmn);
            bc[i] = (e0 + e1) / 2;
            bs[i] = 0.0;
          }
          __syncthreads();
          for (int j = o0 + tid; j < o1; j += NT) {
            const float2 v = zw[j];
            double f = floor(((double)v.x - zmn) / (zmax_in - zmn) * B);
            if (!isnan(f)) atomicAdd(&bs[(int)fmin(fmax(f, 0.0), (double)(B - 1))], (double)v.y);
          }
          __syncthreads();
          Stats6 u = {0.0, 0.0, 0.0, 0.0, 0.f, 0.f};
          for (int i = tid; i < B; i += NT) { u.a += bs[i]; u.b += bs[i] * bs[i]; u.c += bc[i]; u.d += bc[i] * bc[i]; }
          u = block_stats<NW>(u, red);
          W = u.a; Q = u.b;
          const double cmean = u.c / B;
          dstd = sqrt(fmax(u.d / B - cmean * cmean, 0.0));
          for (int i = tid; i < B; i += NT) xwb[i] = make_float2((float)bc[i], (float)bs[i]);
          dxw = xwb; dn = B;
          __syncthreads();
        } else {
          // std of the masked data set: in-pixel samples + (Ns - nin) copies of min(z)
          const double nout = (double)(Ns - nin);
          const double mm_ = (t.c + nout * zmn) / Ns;
          const double ex2_ = (t.d + nout * zmn * zmn) / Ns;
          dstd = sqrt(fmax(ex2_ - mm_ * mm_, 0.0));
        }
        const double neff_k = 1.0 / (Q / (W * W));
        double bw;
        if (a.bw_method == CHB_BW_SCOTT) bw = pow(neff_k, -0.2) * dstd;
        else if (a.bw_method == CHB_BW_SILVERMAN) bw = pow(neff_k * 3.0 / 4.0, -0.2) * dstd;
        else bw = a.bw_value * dstd;
        const double scale = (W != 0.0) ? (norm * gwp[p]) : nan("");
        kde_inplace<NT>(dxw, dn, eg, G, bw, W, CHB_KERNEL_EPAN, 1.0, part, NW * Nz, dens);
        __syncthreads();
        for (int k = tid; k < Nz; k += NT) {
          const double raw = interp_lr(zgrid[k], eg, dens, G, 0.0, 0.0);
          const bool inside = (zgrid[k] >= eg[0] && zgrid[k] <= eg[G - 1]);
          const double v = inside ? raw * scale : 0.0;
          const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
          if (pout) pout[(size_t)p * Nz + k] = v;
          if (pc == -100.0) continue;
          const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
          like_acc += v * pgal * ck[k];
        }
        __syncthreads();
      }
    } else {
      // ---- p_gw3dfull: 3-D whitened Gaussian KDE (likelihood.py:211-260, math.py:154-229) ------
      const double* ra = a.ra + so;
      const double* dec = a.dec + so;
      const double W = s1;
      const double Qn = s2 / (W * W);
      const double neff_k = 1.0 / Qn;
      double factor;
      if (a.bw_method == CHB_BW_SCOTT) factor = pow(neff_k, -1.0 / 7.0);
      else if (a.bw_method == CHB_BW_SILVERMAN) factor = pow(neff_k * 5.0 / 4.0, -1.0 / 7.0);
      else factor = a.bw_value;
      double m0 = 0, m1 = 0, m2 = 0;
      for (int j = tid; j < Ns; j += NT) { const float2 v = zw[j]; double wn = (double)v.y / W; m0 += wn * (double)v.x; m1 += wn * ra[j]; m2 += wn * dec[j]; }
      m0 = block_sum(m0, red); m1 = block_sum(m1, red); m2 = block_sum(m2, red);
      double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
      for (int j = tid; j < Ns; j += NT) {
        const float2 v = zw[j];
        double wn = (double)v.y / W, r0 = (double)v.x - m0, r1 = ra[j] - m1, r2 = dec[j] - m2;
        c00 += wn * r0 * r0; c01 += wn * r0 * r1; c02 += wn * r0 * r2;
        c11 += wn * r1 * r1; c12 += wn * r1 * r2; c22 += wn * r2 * r2;
      }
      c00 = block_sum(c00, red); c01 = block_sum(c01, red); c02 = block_sum(c02, red);
      c11 = block_sum(c11, red); c12 = block_sum(c12, red); c22 = block_sum(c22, red);
      if (tid == 0) {
        const double dn_ = 1.0 - Qn;
        c00 /= dn_; c01 /= dn_; c02 /= dn_; c11 /= dn_; c12 /= dn_; c22 /= dn_;
        const double a00 = c11 * c22 - c12 * c12, a01 = c02 * c12 - c01 * c22, a02 = c01 * c12 - c02 * c11;
        const double a11 = c00 * c22 - c02 * c02, a12 = c01 * c02 - c00 * c12, a22 = c00 * c11 - c01 * c01;
        const double det = c00 * a00 + c01 * a01 + c02 * a02;
        const double f2 = factor * factor;
        const double i00 = a00 / det / f2, i01 = a01 / det / f2, i02 = a02 / det / f2;
        const double i11 = a11 / det / f2, i12 = a12 / det / f2, i22 = a22 / det / f2;
        // Cholesky factor of the inverse covariance with the variables ordered (dec, ra, z): the quadratic form is the
        // same for any order, and with z last the third whitened coordinate y2 = l22 (z - mean z) depends on z alone --
        // monotone in z, so the dL-sorted samples are sorted in y2 and whole blocks of them can be skipped for
        // evaluation points that are far away in z (below)
        const double j00 = i22, j01 = i12, j02 = i02, j11 = i11, j12 = i01, j22 = i00;
        const double l00 = sqrt(j00), l10 = j01 / l00, l20 = j02 / l00;
        const double l11 = sqrt(j11 - l10 * l10), l21 = (j12 - l20 * l10) / l11;
        const double l22 = sqrt(j22 - l20 * l20 - l21 * l21);
        L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
        L[6] = log(l00) + log(l11) + log(l22) - 1.5 * log(2.0 * CHB_PI);
      }
      __syncthreads();
      const double l00 = L[0], l10 = L[1], l11 = L[2], l20 = L[3], l21 = L[4], l22 = L[5], lognorm = L[6];
      const double ps = 0.8493218002880191;            // sqrt(log2(e)/2)
      // whitened coordinates of (z, ra, dec) deviations r0, r1, r2
      auto whiten = [&](double r0, double r1, double r2, float& y0, float& y1, float& y2) {
        y0 = (float)((r2 * l00 + r1 * l10 + r0 * l20) * ps);
        y1 = (float)((r1 * l11 + r0 * l21) * ps);
        y2 = (float)((r0 * l22) * ps);
      };
      for (int j = tid; j < Ns; j += NT) {
        const float2 v = zw[j];
        float y0, y1, y2;
        whiten((double)v.x - m0, ra[j] - m1, dec[j] - m2, y0, y1, y2);
        yw[j] = make_float4(y0, y1, y2, (float)((double)v.y / W));
      }
      if (pout) for (int i = tid; i < Pp * Nz; i += NT) pout[i] = 0.0;
      const double zlo = zmn - a.cut_grid * zstd, zhi = zmx + a.cut_grid * zstd;   // likelihood.py:225
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
      constexpr int FR = 4;                       // evaluation points per lane: two packed pairs (FADD2 / FMUL2 / FFMA2)
      const double enorm = exp(lognorm) * norm;
      // ---- hulls of the 64-sample blocks in y2 and their heaviest weight (one warp per block) ---------------------
      constexpr int SB = 64;
      const int nblk = (Ns + SB - 1) / SB;
      float4* blkh = reinterpret_cast<float4*>(part);          // {min y2, max y2, log2 max w, -}; `part` is idle for this kind
      const bool windows = nblk * 16 <= ((NW * Nz + 1) / 2) * 8 && a.kde_win_iters > 0;
      if (windows) {
        for (int b = warp; b < nblk; b += NW) {
          float lo = INFINITY, hi = -INFINITY, wm = 0.f;
          for (int jj = b * SB + lane; jj < min(Ns, b * SB + SB); jj += 32) {
            const float4 v = yw[jj];
            lo = fminf(lo, v.z); hi = fmaxf(hi, v.z); wm = fmaxf(wm, v.w);
          }
          lo = warp_min_f32(lo); hi = warp_max_f32(hi); wm = warp_max_f32(wm);
          if (lane == 0) blkh[b] = make_float4(lo, hi, (wm > 0.f) ? lg2f_(wm) : -INFINITY, 0.f);
        }
        __syncthreads();
      }
      // Evaluation points ordered pixel-fastest: a tile of 32 FR consecutive points spans only a few z values, i.e. a
      // narrow range of y2.
      const int ntiles = (npts + 32 * FR - 1) / (32 * FR);
      auto tile_points = [&](int t, float (&q0)[FR], float (&q1)[FR], float (&q2)[FR], int (&pk)[FR], float& t2lo, float& t2hi) {
        t2lo = INFINITY; t2hi = -INFINITY;
#pragma unroll
        for (int r = 0; r < FR; ++r) {
          const int i = t * 32 * FR + r * 32 + lane;
          pk[r] = -1;
          q0[r] = q1[r] = q2[r] = 1.0e18f;
          if (i < npts) {
            const int kk = i / npix, p = i - kk * npix, k = kmask[kk];
            pk[r] = p * Nz + k;
            whiten(zgrid[k] - m0, rap[p] - m1, dep[p] - m2, q0[r], q1[r], q2[r]);
            t2lo = fminf(t2lo, q2[r]); t2hi = fmaxf(t2hi, q2[r]);
          }
        }
        t2lo = warp_min_f32(t2lo); t2hi = warp_max_f32(t2hi);
      };
      // Window of a tile: a LOWER bound of the largest single term at each of its points from every 32nd sample (M,
      // minimum over the points), against the UPPER bound of a whole block, 64 x its heaviest weight at the block's
      // nearest approach in y2 alone.  A block is skipped when that bound is below 2^-t2 of M (t2 = 24 bits): what is dropped
      // is below 2^-24 of the largest term AT EVERY POINT of the tile, so far tails keep their relative accuracy.  (M = -inf,
      // every block visited, when a point sees only underflowing terms.)  One pass over the tiles fills Mt[].
      float* Mt = reinterpret_cast<float*>(eg);                // `eg` is idle for this kind
      const bool win3 = windows && ntiles <= 2 * Nz;
      if (win3) {
        for (int t = warp; t < ntiles; t += NW) {
          float q0[FR], q1[FR], q2[FR], mp[FR], t2lo, t2hi;
          int pk[FR];
          tile_points(t, q0, q1, q2, pk, t2lo, t2hi);
#pragma unroll
          for (int r = 0; r < FR; ++r) mp[r] = -INFINITY;
          for (int jj = 0; jj < Ns; jj += 32) {
            const float4 v = yw[jj];
            const float lw = lg2f_(v.w);                         // (w = 0 -> -inf: never the maximum)
#pragma unroll
            for (int r = 0; r < FR; ++r) {
              const float d0 = v.x - q0[r], d1 = v.y - q1[r], d2 = v.z - q2[r];
              mp[r] = fmaxf(mp[r], lw - fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
            }
          }
          float mm = INFINITY;
#pragma unroll
          for (int r = 0; r < FR; ++r) if (pk[r] >= 0) mm = fminf(mm, mp[r]);
          mm = warp_min_f32(mm);
          if (lane == 0) Mt[t] = (mm > -1.0e30f) ? mm : -INFINITY;
        }
        __syncthreads();
      }
      // Work items = (tile) x (slice of the sample blocks): the likelihood is linear in the pair sums, so a warp
      // multiplies its PARTIAL sum of a point by the point's catalogue factor; slices make the item count a multiple of
      // the warp count.  p_gw output wants complete sums per point: one slice then.
      int S = 1;
      while (!pout && (ntiles * S) % NW != 0 && S < 8 && nblk / (2 * S) >= 4) S *= 2;
      const int bper = (nblk + S - 1) / S;
      for (int it = warp; it < ntiles * S; it += NW) {
        const int t = it / S, blo = (it - t * S) * bper, bhi = min(nblk, blo + bper);
        float q0[FR], q1[FR], q2[FR], acc[FR], t2lo, t2hi;
        int pk[FR];
        tile_points(t, q0, q1, q2, pk, t2lo, t2hi);
#pragma unroll
        for (int r = 0; r < FR; ++r) acc[r] = 0.f;
        const float Mtile = win3 ? Mt[t] : -INFINITY;
        // Two points per packed operand: the same d0^2 + d1^2 + d2^2 (same operation order and roundings as the scalar
        // form) in 6 packed instructions per two pairs instead of 12, so the loop is bound by MUFU.EX2, not by issue.
        f32x2 nq0[FR / 2], nq1[FR / 2], nq2[FR / 2], acc2[FR / 2];
#pragma unroll
        for (int r = 0; r < FR / 2; ++r) {
          nq0[r] = pk2(-q0[2 * r], -q0[2 * r + 1]); nq1[r] = pk2(-q1[2 * r], -q1[2 * r + 1]);
          nq2[r] = pk2(-q2[2 * r], -q2[2 * r + 1]); acc2[r] = 0ull;
        }
        for (int b = blo; b < bhi; ++b) {
          if (win3) {
            const float4 hb = blkh[b];
            const float dist = fmaxf(fmaxf(hb.x - t2hi, t2lo - hb.y), 0.f);
            if (hb.z + 6.f - dist * dist < Mtile - a.win_t2) continue;      // (option kde_win_t2, default 24 bits: fp32 carries 24)
          }
          const int jlo = b * SB, jhi = min(Ns, jlo + SB);
#pragma unroll 4
          for (int j = jlo; j < jhi; ++j) {
            const float4 v = yw[j];
            const f32x2 vx = pk2(v.x, v.x), vy = pk2(v.y, v.y), vz = pk2(v.z, v.z), vw = pk2(v.w, v.w);
#pragma unroll
            for (int r = 0; r < FR / 2; ++r) {
              const f32x2 d0 = add2(vx, nq0[r]), d1 = add2(vy, nq1[r]), d2 = add2(vz, nq2[r]);
              const f32x2 e = fma2(d2, d2, fma2(d1, d1, mul2(d0, d0)));
              float e0, e1;
              upk2(e, e0, e1);
              acc2[r] = fma2(vw, pk2(ex2_ftz(-e0), ex2_ftz(-e1)), acc2[r]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < FR / 2; ++r) upk2(acc2[r], acc[2 * r], acc[2 * r + 1]);
#pragma unroll
        for (int r = 0; r < FR; ++r) {
          if (pk[r] < 0) continue;
          const int p = pk[r] / Nz, k = pk[r] - p * Nz;
          const double v = (double)acc[r] * enorm;
          if (pout) pout[(size_t)p * Nz + k] = v;
          const double pc = has_cat ? pcat_ev[(size_t)p * Nz + k] : 0.0;
          if (pc != -100.0) {
            const double pgal = has_cat ? fR * pc + (1.0 - pcompl_ev[k]) * dV[k] : dV[k];
            like_acc += v * pgal * ck[k];
          }
        }
      }
    }

    FPHASE(4);
    {
      // one barrier: warp sums -> thread 0 (`red` is idle here; the barrier at the top of the next unit protects it)
      const double ws = warp_sum(like_acc);
      if (lane == 0) red[warp] = ws;
      __syncthreads();
      if (tid == 0) {
        double like = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) like += red[w];
        a.log_like[(size_t)h * a.Nev + ev] = nan_to_num_log_f(like); a.like_raw[(size_t)h * a.Nev + ev] = like;
      }
    }
    FPHASE(5);
  }
  if (a.prof && tid == 0) for (int i = 0; i < 8; ++i) a.prof[(size_t)blockIdx.x * 8 + i] = pacc[i];
}

static inline int kind_group(int kind) { return kind == CHB_PGW_MARG ? 1 : (kind == CHB_PGW_FULL ? 2 : 0); }
// dispatch over (kind group, mode) -> kernel instantiation
#define CHB_F32_DISPATCH(KG_, MODE_, EXPR)                                                     \
  switch ((KG_) * 3 + (MODE_)) {                                                               \
    case 0: { auto kern = numerator_f32_kernel<0, 0, F_NT>; EXPR; } break;                           \
    case 1: { auto kern = numerator_f32_kernel<0, 1, F_NT>; EXPR; } break;                           \
    case 2: { auto kern = numerator_f32_kernel<0, 2, F_NT>; EXPR; } break;                           \
    case 3: { auto kern = numerator_f32_kernel<1, 0, F_NT>; EXPR; } break;                           \
    case 4: { auto kern = numerator_f32_kernel<1, 1, F_NT>; EXPR; } break;                           \
    case 5: { auto kern = numerator_f32_kernel<1, 2, F_NT>; EXPR; } break;                           \
    case 6: { auto kern = numerator_f32_kernel<2, 0, F_NT>; EXPR; } break;                           \
    case 7: { auto kern = numerator_f32_kernel<2, 1, F_NT>; EXPR; } break;                           \
    default: { auto kern = numerator_f32_kernel<2, 2, F_NT_FULL2>; EXPR; } break;                          \
  }
// `optin`: the device's opt-in maximum of shared memory per block; the attribute is set to that maximum minus the kernel's
// static shared memory (the limit applies to static + dynamic), never to a handle's own footprint.
cudaError_t numerator_f32_configure(int kind, int mode, size_t optin) {
  cudaError_t e = cudaSuccess;
  cudaFuncAttributes fa;
  CHB_F32_DISPATCH(kind_group(kind), mode, {
    e = cudaFuncGetAttributes(&fa, kern);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(optin - fa.sharedSizeBytes));
  });
  return e;
}
int numerator_f32_ctas_per_sm(int kind, int mode, size_t smem) {
  int n = 0;
  cudaError_t e = cudaSuccess;
  CHB_F32_DISPATCH(kind_group(kind), mode, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, f32_threads(kind, mode), smem));
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
cudaError_t launch_numerator_f32(const NumArgs& a, int mode, int grid, size_t smem, cudaStream_t s) {
  CHB_F32_DISPATCH(kind_group(a.kind), mode, (kern<<<grid, f32_threads(a.kind, mode), smem, s>>>(a)));
  return cudaGetLastError();
}
