// selection.cu -- injection-reweighted Monte-Carlo selection function (kernel 1).
//
// For every (injection, hyper-point) pair: dN/dtheta_det of the population at the injection
// (CHIMERA/population/pop_wrapper.py:102-111) divided by p_draw, reduced to the two sums
// selection_function.N_exp needs (CHIMERA/selection_function.py:37-44):
//     S1 = nansum(w),  S2 = sum(w^2),   w = R0 p(m1s,m2s) dVc/dz psi(z)/(1+z) / (|ddL/dz| (1+z)^2 p_draw)
// Grid = (injection tiles, hyper-points).  Each CTA stages its hyper-point's table block in
// shared memory with one TMA bulk copy, streams a contiguous tile of the four injection arrays
// with coalesced 8-byte loads, and writes one (S1,S2) pair; reduce.cu adds the tiles in a fixed
// order, so results are bit-reproducible run to run.  fp64 throughout.
#include "common.cuh"
#include "models_f32.cuh"
#include "stage.cuh"

__global__ void __launch_bounds__(256)
selection_kernel(SelArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ double red[32];
  const TableLayout lay = a.mc.lay;
  const int h = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const uint32_t tab_bytes = (uint32_t)(lay.f64_total() * sizeof(double));

  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, tab_bytes);
    bulk_g2s(sm, a.tabs + (size_t)h * lay.total(), tab_bytes, &bar);
  }
  if (tid < CHB_NPAR) P[tid] = a.hyper[(size_t)h * CHB_NPAR + tid];
  if (tid < CHB_NHC) HC[tid] = a.HC[(size_t)h * CHB_NHC + tid];
  __syncthreads();
  mbar_wait(&bar, 0);

  const double* zg = sm + lay.off_zg();
  const double* dLt = sm + lay.off_dLt();
  const double* mg = sm + lay.off_mg();
  const double* cdf = sm + lay.off_cdf();
  const int rc = lay.rc, rm = lay.rm;
  const int cm = a.mc.cosmo_model, mm = a.mc.mass_model, rmod = a.mc.rate_model;
  const double R0 = P[CHB_P_R0];

  const long long chunk = ((long long)a.Ninj + a.tiles - 1) / a.tiles;
  const long long j0 = (long long)tile * chunk;
  const long long j1 = min((long long)a.Ninj, j0 + chunk);
  double s1 = 0.0, s2 = 0.0;
  // table intervals without binary searches (models.cuh upper_index_lut / upper_index_log): bit-identical interpolants
  const unsigned short* lut = reinterpret_cast<const unsigned short*>(a.tabs + (size_t)h * lay.total() + lay.off_f32() + lay.f32_lut());
  const int lut_b0 = (int)HC[HC_LUT_B0], lut_nb = (int)HC[HC_LUT_NB];
  for (long long j = j0 + tid; j < j1; j += blockDim.x) {
    const double dL = __ldg(a.dL + j), m1d = __ldg(a.m1d + j), m2d = __ldg(a.m2d + j), pd = __ldg(a.p_draw + j);
    const double z = interp_at(dL, dLt, zg, rc, upper_index_lut(dLt, rc, dL, lut, lut_b0, lut_nb));   // z_from_dGW
    const double opz = 1.0 + z;
    const double m1 = m1d / opz, m2 = m2d / opz;                      // theta_det2src
    const double dCt = dL2dCt(cm, P, dL, z);                          // original distances
    const double Ez = E_at_z(P, HC, z);
    const double pz = dVcdz_from(HC, dCt, Ez) * (merger_rate(rmod, P, HC, z) / opz);
    const double dN = R0 * p_m1m2_logidx(mm, P, HC, mg, cdf, rm, m1, m2) * pz;
    const double jac = fabs(ddLdz_from(cm, P, HC, z, dCt, Ez)) * (opz * opz);
    const double w = (dN / jac) / pd;
    if (!isnan(w)) s1 += w;     // nansum for xi (selection_function.py:39)
    s2 += w * w;                // plain sum for the variance (:44)
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (tid == 0) {
    double* out = a.tile_part + ((size_t)h * a.tiles + tile) * 2;
    out[0] = s1;
    out[1] = s2;
  }
}

// fp32 variant (CHB_FP32): same reduction, the per-injection rate evaluated with the fp32 fast path
// (models_f32.cuh); only the dl4 | cd4 | lut part of the fp32 table block is staged (42 KB); sums
// accumulate in fp64.
__global__ void __launch_bounds__(256)
selection_f32_kernel(SelArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double P[CHB_NPAR];
  __shared__ double HC[CHB_NHC];
  __shared__ double red[32];
  const TableLayout lay = a.mc.lay;
  const int h = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const int part_off = lay.f32_dl4();
  const uint32_t tab_bytes = (uint32_t)((lay.f32_core() - part_off) * sizeof(double));
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, tab_bytes);
    bulk_g2s(sm, a.tabs + (size_t)h * lay.total() + lay.off_f32() + part_off, tab_bytes, &bar);
  }
  if (tid < CHB_NPAR) P[tid] = a.hyper[(size_t)h * CHB_NPAR + tid];
  if (tid < CHB_NHC) HC[tid] = a.HC[(size_t)h * CHB_NHC + tid];
  __syncthreads();
  mbar_wait(&bar, 0);
  F32Consts fc = make_f32_consts(a.mc, P, HC, sm - part_off);   // zi4 (not staged) is never touched here
  fc.cd4_last = fc.cd4[fc.rm - 1];
  const CosmoRateF32 cr = make_cosmo_rate_f32(a.mc, P, HC);
  const float z_top = (float)a.tabs[(size_t)h * lay.total() + lay.off_zg() + lay.rc - 1];

  const long long chunk = ((long long)a.Ninj + a.tiles - 1) / a.tiles;
  const long long j0 = (long long)tile * chunk;
  const long long j1 = min((long long)a.Ninj, j0 + chunk);
  double s1 = 0.0, s2 = 0.0;
  // software pipeline: the packed injection of the next iteration is requested before the current one is
  // evaluated (the loads are L2 hits ~600 cycles away; the arithmetic of one injection is ~200 instructions).
  // Every thread runs the same number of iterations (tail lanes re-evaluate the tile's last injection with weight 0), so
  // the warp can vote: the injections are stored sorted by their (estimated) source-frame secondary mass (api.cu), and a
  // warp whose 32 injections are all above m_low + delta_m skips the taper.
  const int iters = (int)((j1 - j0 + blockDim.x - 1) / blockDim.x);
  const long long jlast = (j1 > j0) ? j1 - 1 : j0;
  long long j = j0 + tid;
  float4 sv = make_float4(1.f, 1.f, 1.f, 0.f);
  float2 lv = make_float2(0.f, 0.f);
  if (iters > 0) { sv = __ldg(a.s4 + min(j, jlast)); lv = __ldg(a.l2 + min(j, jlast)); }
  for (int it = 0; it < iters; ++it) {
    const long long jn = j + blockDim.x;
    float4 nsv = sv; float2 nlv = lv;
    if (it + 1 < iters) { nsv = __ldg(a.s4 + min(jn, jlast)); nlv = __ldg(a.l2 + min(jn, jlast)); }
    // z_from_dGW without the zi4 clamp row: same scan, clamp with the staged last knot
    int b = (int)(__float_as_uint(sv.x) >> CHB_LUT_SHIFT) - (int)fc.b0;
    b = max(0, min(b, fc.nb - 1));
    int k = fc.lut[b];
    float4 e = fc.dl4[k];
    while (sv.x >= e.w && k < fc.rc - 2) { ++k; e = fc.dl4[k]; }
    float z = fmaf(sv.x - e.x, e.z, e.y);
    if (sv.x >= e.w) z = z_top;
    if (sv.x <= 0.f) z = 0.f;
    const float opz = 1.f + z;
    const float r = rcpf_(opz), lz = lg2f_(opz);
    const float m1 = sv.y * r, m2 = sv.z * r;
    // (m2 <= m1 inside the support, so the taper test on m2 covers m1 whenever the weight is not already zero)
    const bool taper = !(m2 - fc.lo > fc.dm);
    const float pm = __any_sync(0xffffffffu, taper) ? weight_bf<true>(fc, m1, m2, lv.x - lz, lv.y - lz, sv.w)
                                                    : weight_bf<false>(fc, m1, m2, lv.x - lz, lv.y - lz, sv.w);   // p_m1m2 / p_draw
    const float wf = (j < j1) ? cr.R0 * pm * zterm_inj_f32(cr, z, opz, lz, sv.x) : 0.f;
    const double w = (double)wf;
    if (wf == wf) s1 += w;
    s2 += w * w;
    j = jn; sv = nsv; lv = nlv;
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (tid == 0) {
    double* out = a.tile_part + ((size_t)h * a.tiles + tile) * 2;
    out[0] = s1;
    out[1] = s2;
  }
}

int selection_ctas_per_sm(const ModelCfg& mc, int fp_mode) {
  int n = 0;
  cudaError_t e;
  if (fp_mode == CHB_FP32) {
    const size_t smem = (size_t)(mc.lay.f32_core() - mc.lay.f32_dl4()) * sizeof(double);
    e = cudaFuncSetAttribute(selection_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, selection_f32_kernel, 256, smem);
  } else {
    const size_t smem = (size_t)mc.lay.f64_total() * sizeof(double);
    e = cudaFuncSetAttribute(selection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, selection_kernel, 256, smem);
  }
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

cudaError_t launch_selection(const SelArgs& a, cudaStream_t s) {
  if (a.fp_mode == CHB_FP32) {
    size_t smem32 = (size_t)(a.mc.lay.f32_core() - a.mc.lay.f32_dl4()) * sizeof(double);
    cudaError_t e32 = cudaFuncSetAttribute(selection_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
    if (e32 != cudaSuccess) return e32;
    dim3 grid32(a.tiles, a.n_hyper);
    selection_f32_kernel<<<grid32, 256, smem32, s>>>(a);
    return cudaGetLastError();
  }
  size_t smem = (size_t)a.mc.lay.f64_total() * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(selection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(a.tiles, a.n_hyper);
  selection_kernel<<<grid, 256, smem, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// reduce: per hyper-point, sum log-likelihoods over events and (S1,S2) over injection tiles in a
// fixed order -> partials (n_hyper, 3).  -DBL_MAX entries overflow to -inf when two or more are
// present, exactly like the reference's jnp.sum over nan_to_num'ed values (likelihood.py:297-298).
__global__ void __launch_bounds__(256)
reduce_kernel(int n_hyper, int Nev, int tiles, const double* __restrict__ log_like,
              const double* __restrict__ tile_part, double* __restrict__ partials) {
  __shared__ double red[32];
  const int h = blockIdx.x, tid = threadIdx.x;
  double s = 0.0;
  if (log_like) for (int e = tid; e < Nev; e += blockDim.x) s += log_like[(size_t)h * Nev + e];
  s = block_sum(s, red);
  double s1 = 0.0, s2 = 0.0;
  if (tile_part) for (int t = tid; t < tiles; t += blockDim.x) {
    s1 += tile_part[((size_t)h * tiles + t) * 2];
    s2 += tile_part[((size_t)h * tiles + t) * 2 + 1];
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (tid == 0) {
    partials[(size_t)h * 3 + 0] = s;
    partials[(size_t)h * 3 + 1] = s1;
    partials[(size_t)h * 3 + 2] = s2;
  }
}

cudaError_t launch_reduce(int n_hyper, int Nev, int tiles, const double* d_log_like, const double* d_tile_part,
                          double* d_partials, cudaStream_t s) {
  reduce_kernel<<<n_hyper, 256, 0, s>>>(n_hyper, Nev, tiles, d_log_like, d_tile_part, d_partials);
  return cudaGetLastError();
}
