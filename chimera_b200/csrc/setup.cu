// setup.cu -- the callers on the INPUT side of the likelihood path, on the GPU:
//   * HEALPix RING ang2pix / pix2ang (what the reference gets from healpy: utils/angles.py:32-85, data.py:258);
//   * assignment of every posterior sample to one of its event's sky pixels and the 2-D Gaussian KDE of the
//     localisation at the pixel centres (data.py:316-345, utils/math.py:95-148);
//   * pixelated_catalog.precompute_p_cat (catalog/catalog.py:143-231): bucketing of the galaxies by pixel and,
//     per (event, pixel), the sum of galaxy redshift Gaussians x dVc/dz, each normalised on the event's z grid.
// Everything is fp64 and follows the reference operation by operation; this file is compiled with -fmad=false
// so that the index arithmetic of ang2pix rounds exactly like the host libraries (pixel ids must be bit-exact).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include "../../include/chimera_b200.h"
#include "devguard.cuh"

int chb_fail_global(int code, const char* msg);      // api.cu

#define SPI 3.141592653589793238462643383279502884197
#define S_TWOTHIRD (2.0 / 3.0)
#define S_HALFPI (0.5 * SPI)
#define S_INV_HALFPI (2.0 / SPI)

// ---------------------------------------------------------------------------------------------- HEALPix
// healpix_cxx loc2pix (RING), as restated in chimera_b200/healpix.py
__device__ __forceinline__ long long ang2pix_ring(long long nside, double theta, double phi) {
  const double z = cos(theta);
  const bool have_sth = (theta < 0.01) || (theta > 3.14159 - 0.01);
  const double sth = have_sth ? sin(theta) : 0.0;
  const double za = fabs(z);
  double tt = fmod(phi * S_INV_HALFPI, 4.0);                    // numpy.mod: result takes the sign of the divisor
  if (tt != 0.0 && tt < 0.0) tt += 4.0;
  if (tt >= 4.0) tt = 0.0;
  const long long nl4 = 4 * nside, ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside;
  if (za <= S_TWOTHIRD) {
    const double temp1 = (double)nside * (0.5 + tt);
    const double temp2 = (double)nside * z * 0.75;
    const long long jp = (long long)(temp1 - temp2);
    const long long jm = (long long)(temp1 + temp2);
    const long long ir = nside + 1 + jp - jm;
    const long long kshift = 1 - (ir & 1);
    const long long t1 = jp + jm - nside + kshift + 1 + nl4 + nl4;
    const long long ip = (t1 >> 1) & (nl4 - 1);
    return ncap + (ir - 1) * nl4 + ip;
  }
  const double tp = tt - floor(tt);
  const double tmp = (za < 0.99 || !have_sth) ? (double)nside * sqrt(3.0 * (1.0 - za))
                                              : (double)nside * sth / sqrt((1.0 + za) / 3.0);
  const long long jp = (long long)(tp * tmp);
  const long long jm = (long long)((1.0 - tp) * tmp);
  const long long ir = jp + jm + 1;
  long long ip = (long long)(tt * (double)ir);
  ip = min(ip, 4 * ir - 1);
  return (z > 0) ? 2 * ir * (ir - 1) + ip : npix - 2 * ir * (ir + 1) + ip;
}

__device__ __forceinline__ long long isqrt_ll(long long v) {
  long long r = (long long)floor(sqrt((double)v + 0.5));
  if (r * r > v) --r;
  if ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

// healpix_cxx pix2loc (RING) -> (theta, phi)
__device__ __forceinline__ void pix2ang_ring(long long nside, long long pix, double& theta, double& phi) {
  const long long npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1), nl4 = 4 * nside;
  const double fact2 = 4.0 / (double)npix;
  const double fact1 = (double)(2 * nside) * fact2;
  double z, sth = 0.0;
  bool have_sth = false;
  if (pix < ncap) {
    const long long iring = (1 + isqrt_ll(1 + 2 * pix)) >> 1;
    const long long iphi = (pix + 1) - 2 * iring * (iring - 1);
    const double tmp = (double)(iring * iring) * fact2;
    z = 1.0 - tmp;
    if (z > 0.99) { sth = sqrt(tmp * (2.0 - tmp)); have_sth = true; }
    phi = ((double)iphi - 0.5) * S_HALFPI / (double)iring;
  } else if (pix < npix - ncap) {
    const long long p = pix - ncap;
    const long long tmp = p / nl4;
    const long long iring = tmp + nside;
    const long long iphi = p - nl4 * tmp + 1;
    const double fodd = (((iring + nside) & 1) == 1) ? 1.0 : 0.5;
    z = (double)(2 * nside - iring) * fact1;
    phi = ((double)iphi - fodd) * SPI * 0.75 * fact1;
  } else {
    const long long p = npix - pix;
    const long long iring = (1 + isqrt_ll(2 * p - 1)) >> 1;
    const long long iphi = 4 * iring + 1 - (p - 2 * iring * (iring - 1));
    const double tmp = (double)(iring * iring) * fact2;
    z = tmp - 1.0;
    if (z < -0.99) { sth = sqrt(tmp * (2.0 - tmp)); have_sth = true; }
    phi = ((double)iphi - 0.5) * S_HALFPI / (double)iring;
  }
  theta = have_sth ? atan2(sth, z) : acos(fmin(fmax(z, -1.0), 1.0));
}

__global__ void ang2pix_kernel(long long nside, long long n, const double* __restrict__ theta,
                               const double* __restrict__ phi, long long* __restrict__ pix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pix[i] = ang2pix_ring(nside, theta[i], phi[i]);
}
// find_pix_RAdec (utils/angles.py:32-45): theta = pi/2 - dec, phi = ra
__global__ void radec2pix_kernel(long long nside, long long n, const double* __restrict__ ra,
                                 const double* __restrict__ dec, long long* __restrict__ pix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pix[i] = ang2pix_ring(nside, 0.5 * SPI - dec[i], ra[i]);
}
__global__ void pix2ang_kernel(long long nside, long long n, const long long* __restrict__ pix,
                               double* __restrict__ theta, double* __restrict__ phi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double t, p;
    pix2ang_ring(nside, pix[i], t, p);
    theta[i] = t; phi[i] = p;
  }
}

// ---------------------------------------------------------------------------------------------- pixelisation
// data.py:316-340: a sample whose own pixel is one of the event's pixels keeps it; otherwise it takes the pixel
// whose centre has the smallest angular separation (first minimum; a NaN separation -- arccos of a cosine
// rounded above 1 -- counts as the minimum, like numpy.argmin).  grid = (sample tiles, events).
__global__ void __launch_bounds__(256)
assign_pixels_kernel(int Nev, int Ns, int P, const long long* __restrict__ opt_nsides, const double* __restrict__ ra,
                     const double* __restrict__ dec, const long long* __restrict__ pixels,
                     const double* __restrict__ ra_pix, const double* __restrict__ dec_pix,
                     long long* __restrict__ pe_pix) {
  extern __shared__ double sm[];
  double* sd = sm;              // sin(dec_pix)
  double* cd = sm + P;          // cos(dec_pix)
  double* rp = sm + 2 * P;
  long long* px = reinterpret_cast<long long*>(sm + 3 * P);
  __shared__ int npx_s;
  const int e = blockIdx.y;
  if (threadIdx.x == 0) npx_s = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const long long id = pixels[(size_t)e * P + p];
    px[p] = id;
    if (id != -100) {
      atomicMax(&npx_s, p + 1);
      sd[p] = sin(dec_pix[(size_t)e * P + p]); cd[p] = cos(dec_pix[(size_t)e * P + p]); rp[p] = ra_pix[(size_t)e * P + p];
    }
  }
  __syncthreads();
  const int npx = npx_s;
  const long long nside = opt_nsides[e];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Ns; j += gridDim.x * blockDim.x) {
    const double r = ra[(size_t)e * Ns + j], d = dec[(size_t)e * Ns + j];
    const long long own = ang2pix_ring(nside, 0.5 * SPI - d, r);
    bool valid = false;
    for (int p = 0; p < npx; ++p) valid |= (px[p] == own);
    long long out = own;
    if (!valid && npx > 0) {
      const double s = sin(d), c = cos(d);
      int best = 0;
      double ba = 0.0;
      bool locked = false;
      for (int p = 0; p < npx; ++p) {
        const double ca = s * sd[p] + c * cd[p] * cos(r - rp[p]);     // angular_separation_from_LOS (angles.py:158)
        const double a = acos(ca);
        if (p == 0) { ba = a; locked = isnan(a); }
        else if (!locked && (isnan(a) || a < ba)) { ba = a; best = p; locked = isnan(a); }
      }
      out = px[best];
    }
    pe_pix[(size_t)e * Ns + j] = out;
  }
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// gw_loc2d_pdf = jax_gkde_nd((ra, dec) samples, pixel centres) (data.py:343-345, math.py:95-148): unweighted
// 2-D Gaussian KDE, Scott factor n^(-1/6), covariance / (1 - 1/n), whitening by chol(inv(cov)/factor^2).
// One CTA per event.
__global__ void __launch_bounds__(256)
loc2d_pdf_kernel(int Ns, int P, const double* __restrict__ ra, const double* __restrict__ dec,
                 const double* __restrict__ ra_pix, const double* __restrict__ dec_pix, double* __restrict__ pdf) {
  __shared__ double red[32];
  __shared__ double L[4];
  const int e = blockIdx.x, tid = threadIdx.x;
  const double* r = ra + (size_t)e * Ns;
  const double* d = dec + (size_t)e * Ns;
  const double w = 1.0 / (double)Ns;
  double m0 = 0, m1 = 0;
  for (int j = tid; j < Ns; j += blockDim.x) { m0 += w * r[j]; m1 += w * d[j]; }
  m0 = block_sum_d(m0, red); m1 = block_sum_d(m1, red);
  double c00 = 0, c01 = 0, c11 = 0;
  for (int j = tid; j < Ns; j += blockDim.x) {
    const double a = r[j] - m0, b = d[j] - m1;
    c00 += (a * w) * a; c01 += (a * w) * b; c11 += (b * w) * b;
  }
  c00 = block_sum_d(c00, red); c01 = block_sum_d(c01, red); c11 = block_sum_d(c11, red);
  if (tid == 0) {
    const double dn = 1.0 - (double)Ns * (w * w);
    c00 /= dn; c01 /= dn; c11 /= dn;
    const double neff = 1.0 / ((double)Ns * (w * w));
    const double factor = pow(neff, -1.0 / 6.0);
    const double det = c00 * c11 - c01 * c01;
    const double f2 = factor * factor;
    const double i00 = c11 / det / f2, i01 = -c01 / det / f2, i11 = c00 / det / f2;
    const double l00 = sqrt(i00), l10 = i01 / l00, l11 = sqrt(i11 - l10 * l10);
    L[0] = l00; L[1] = l10; L[2] = l11;
    L[3] = log(l00) + log(l11) - 0.5 * 2.0 * log(2.0 * SPI);
  }
  __syncthreads();
  const double l00 = L[0], l10 = L[1], l11 = L[2], lognorm = L[3];
  for (int p = 0; p < P; ++p) {
    const double rp = ra_pix[(size_t)e * P + p];
    if (rp == -100.0) { if (tid == 0) pdf[(size_t)e * P + p] = -100.0; continue; }
    const double dp = dec_pix[(size_t)e * P + p];
    // whitened coordinates: x L (row vector times lower-triangular L)
    const double q0 = rp * l00 + dp * l10, q1 = dp * l11;
    double acc = 0.0;
    for (int j = tid; j < Ns; j += blockDim.x) {
      const double y0 = r[j] * l00 + d[j] * l10, y1 = d[j] * l11;
      const double a = y0 - q0, b = y1 - q1;
      acc += w * exp(lognorm - 0.5 * (a * a + b * b));
    }
    acc = block_sum_d(acc, red);
    if (tid == 0) pdf[(size_t)e * P + p] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- catalogue
__global__ void gather3_kernel(long long n, const long long* __restrict__ perm, const double* __restrict__ a,
                               const double* __restrict__ b, const double* __restrict__ c, double* __restrict__ oa,
                               double* __restrict__ ob, double* __restrict__ oc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long s = perm[i];
    oa[i] = a[s]; ob[i] = b[s]; oc[i] = c[s];
  }
}

// p_cat[e, i, :] (catalog.py:143-195,209-221) for the events of ONE nside.  grid = (P, events of this nside);
// galaxies sorted by pixel id (keys ascending); every warp takes galaxies of the pixel round-robin, the lanes
// span the z grid; per galaxy: Gaussian x dVc/dz on the grid, trapezoid norm on the grid, accumulate w g / norm.
__global__ void __launch_bounds__(128)
p_cat_kernel(int Nz, int P, int n_ev_sel, const int* __restrict__ ev_sel, const long long* __restrict__ pixels,
             const int* __restrict__ neff_pix, const double* __restrict__ zgrids, const double* __restrict__ dVdz,
             long long ngal, const long long* __restrict__ keys, const double* __restrict__ gz,
             const double* __restrict__ gsig, const double* __restrict__ gw, double* __restrict__ p_cat,
             double* __restrict__ N_gal) {
  extern __shared__ double sm[];
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* zg = sm;                      // Nz
  double* dv = sm + Nz;                 // Nz
  double* val = sm + 2 * Nz + (size_t)warp * Nz;          // per warp
  double* acc = sm + 2 * Nz + (size_t)nw * Nz + (size_t)warp * Nz;
  __shared__ double wsum_s[4];
  __shared__ int cnt_s[4];
  const int e = ev_sel[blockIdx.y], i = blockIdx.x;
  if (i >= neff_pix[e]) return;         // padded slots keep the -100 the caller filled in
  for (int k = threadIdx.x; k < Nz; k += blockDim.x) { zg[k] = zgrids[(size_t)e * Nz + k]; dv[k] = dVdz[(size_t)e * Nz + k]; }
  for (int k = lane; k < Nz; k += 32) acc[k] = 0.0;
  __syncthreads();
  const long long pid = pixels[(size_t)e * P + i];
  // [a, b) = galaxies with key == pid
  long long lo = 0, hi = ngal;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < pid) lo = mid + 1; else hi = mid; }
  const long long a = lo;
  hi = ngal;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < pid + 1) lo = mid + 1; else hi = mid; }
  const long long b = lo;
  const double z0 = zg[0], z1 = zg[Nz - 1];
  double wsum = 0.0;
  int cnt = 0;
  for (long long g = a + warp; g < b; g += nw) {
    const double zgal = gz[g], sg = gsig[g], wg = gw[g];
    if (!(zgal > z0 && zgal < z1)) continue;                          // catalog.py:154-157 (strict)
    wsum += wg; ++cnt;
    const double pref = pow(2.0 * SPI * (sg * sg), -0.5);
    for (int k = lane; k < Nz; k += 32) {
      const double u = (zg[k] - zgal) / sg;
      val[k] = pref * exp(-0.5 * (u * u)) * dv[k];
    }
    __syncwarp();
    double nrm = 0.0;
    for (int k = lane; k < Nz - 1; k += 32) nrm += (zg[k + 1] - zg[k]) * (val[k] + val[k + 1]) / 2.0;
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    for (int k = lane; k < Nz; k += 32) acc[k] += wg * val[k] / nrm;
    __syncwarp();
  }
  if (lane == 0) { wsum_s[warp] = wsum; cnt_s[warp] = cnt; }
  __syncthreads();
  double W = 0.0;
  int C = 0;
  for (int w = 0; w < nw; ++w) { W += wsum_s[w]; C += cnt_s[w]; }
  double* out = p_cat + ((size_t)e * P + i) * Nz;
  const double* acc0 = sm + 2 * Nz + (size_t)nw * Nz;
  for (int k = threadIdx.x; k < Nz; k += blockDim.x) {
    double r = 0.0;
    if (C > 0) {
      for (int w = 0; w < nw; ++w) r += acc0[(size_t)w * Nz + k];
      r = r / W;
      if (!isfinite(r)) r = 0.0;                                        // catalog.py:191
    }
    out[k] = r;
  }
  if (threadIdx.x == 0 && C > 0) atomicAdd(&N_gal[e], (double)C);
}

// ---------------------------------------------------------------------------------------------- C ABI
namespace {
struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 8)); }
  cudaError_t up(const void* src, size_t bytes) {
    cudaError_t e = alloc(bytes);
    return e != cudaSuccess ? e : cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
  }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
int cuda_err(cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return chb_fail_global(CHB_ERR_CUDA, m.c_str());
}
int pick_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return chb_fail_global(CHB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) return chb_fail_global(CHB_ERR_INVALID, "device ordinal out of range");
  cudaError_t e = cudaSetDevice(device);
  return e == cudaSuccess ? CHB_OK : cuda_err(e, "cudaSetDevice");
}
bool pow2(int64_t v) { return v >= 1 && (v & (v - 1)) == 0; }
int grid_for(int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), 148 * 16); }
}  // namespace
#define SCU(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_err(_e, what); } while (0)

extern "C" {

int chb_healpix_ang2pix_ring(int device, int64_t nside, int64_t n, const double* theta, const double* phi, int64_t* pix) {
  if (!pow2(nside) || nside > (1 << 29)) return chb_fail_global(CHB_ERR_INVALID, "nside must be a positive power of 2");
  if (n < 0 || (n > 0 && (!theta || !phi || !pix))) return chb_fail_global(CHB_ERR_INVALID, "bad array arguments");
  for (int64_t i = 0; i < n; ++i)
    if (!(theta[i] >= 0.0 && theta[i] <= SPI)) return chb_fail_global(CHB_ERR_INVALID, "theta out of range [0, pi]");
  DevGuard _dg;
  int rc = pick_device(device);
  if (rc != CHB_OK || n == 0) return rc;
  Buf dt, dp, dx;
  SCU(dt.up(theta, n * sizeof(double)), "upload theta");
  SCU(dp.up(phi, n * sizeof(double)), "upload phi");
  SCU(dx.alloc(n * sizeof(long long)), "alloc pix");
  ang2pix_kernel<<<grid_for(n), 256>>>(nside, n, dt.as<double>(), dp.as<double>(), dx.as<long long>());
  SCU(cudaGetLastError(), "ang2pix launch");
  SCU(cudaMemcpy(pix, dx.p, n * sizeof(long long), cudaMemcpyDeviceToHost), "D2H pix");
  return CHB_OK;
}

int chb_healpix_pix2ang_ring(int device, int64_t nside, int64_t n, const int64_t* pix, double* theta, double* phi) {
  if (!pow2(nside) || nside > (1 << 29)) return chb_fail_global(CHB_ERR_INVALID, "nside must be a positive power of 2");
  if (n < 0 || (n > 0 && (!theta || !phi || !pix))) return chb_fail_global(CHB_ERR_INVALID, "bad array arguments");
  const int64_t npix = 12 * nside * nside;
  for (int64_t i = 0; i < n; ++i)
    if (pix[i] < 0 || pix[i] >= npix) return chb_fail_global(CHB_ERR_INVALID, "pixel index out of range");
  DevGuard _dg;
  int rc = pick_device(device);
  if (rc != CHB_OK || n == 0) return rc;
  Buf dt, dp, dx;
  SCU(dx.up(pix, n * sizeof(long long)), "upload pix");
  SCU(dt.alloc(n * sizeof(double)), "alloc theta");
  SCU(dp.alloc(n * sizeof(double)), "alloc phi");
  pix2ang_kernel<<<grid_for(n), 256>>>(nside, n, dx.as<long long>(), dt.as<double>(), dp.as<double>());
  SCU(cudaGetLastError(), "pix2ang launch");
  SCU(cudaMemcpy(theta, dt.p, n * sizeof(double), cudaMemcpyDeviceToHost), "D2H theta");
  SCU(cudaMemcpy(phi, dp.p, n * sizeof(double), cudaMemcpyDeviceToHost), "D2H phi");
  return CHB_OK;
}

int chb_pixelize_samples(int device, int64_t Nev, int64_t Ns, int64_t P, const int64_t* opt_nsides, const double* ra,
                         const double* dec, const int64_t* pixels_opt_nsides, const double* ra_pix,
                         const double* dec_pix, int64_t* pixels_pe_opt_nside, double* gw_loc2d_pdf) {
  if (Nev < 1 || Ns < 2 || P < 1 || !opt_nsides || !ra || !dec || !pixels_opt_nsides || !ra_pix || !dec_pix)
    return chb_fail_global(CHB_ERR_INVALID, "bad pixelisation arguments");
  if (!pixels_pe_opt_nside && !gw_loc2d_pdf) return chb_fail_global(CHB_ERR_INVALID, "no output requested");
  for (int64_t e = 0; e < Nev; ++e)
    if (!pow2(opt_nsides[e])) return chb_fail_global(CHB_ERR_INVALID, "nside must be a positive power of 2");
  DevGuard _dg;
  int rc = pick_device(device);
  if (rc != CHB_OK) return rc;
  const size_t ns = (size_t)Nev * Ns, np = (size_t)Nev * P;
  Buf dns, dra, ddec, dpx, drp, ddp, dout, dpdf;
  SCU(dns.up(opt_nsides, Nev * sizeof(long long)), "upload opt_nsides");
  SCU(dra.up(ra, ns * sizeof(double)), "upload ra");
  SCU(ddec.up(dec, ns * sizeof(double)), "upload dec");
  SCU(dpx.up(pixels_opt_nsides, np * sizeof(long long)), "upload pixels");
  SCU(drp.up(ra_pix, np * sizeof(double)), "upload ra_pix");
  SCU(ddp.up(dec_pix, np * sizeof(double)), "upload dec_pix");
  if (pixels_pe_opt_nside) {
    SCU(dout.alloc(ns * sizeof(long long)), "alloc pe pixels");
    const size_t smem = (size_t)P * (3 * sizeof(double) + sizeof(long long));
    if (smem > 200 * 1024) return chb_fail_global(CHB_ERR_UNSUPPORTED, "too many pixels per event");
    SCU(cudaFuncSetAttribute(assign_pixels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    dim3 grid((unsigned)std::min<int64_t>((Ns + 255) / 256, 64), (unsigned)Nev);
    assign_pixels_kernel<<<grid, 256, smem>>>((int)Nev, (int)Ns, (int)P, dns.as<long long>(), dra.as<double>(),
                                              ddec.as<double>(), dpx.as<long long>(), drp.as<double>(), ddp.as<double>(),
                                              dout.as<long long>());
    SCU(cudaGetLastError(), "assign_pixels launch");
    SCU(cudaMemcpy(pixels_pe_opt_nside, dout.p, ns * sizeof(long long), cudaMemcpyDeviceToHost), "D2H pe pixels");
  }
  if (gw_loc2d_pdf) {
    SCU(dpdf.alloc(np * sizeof(double)), "alloc pdf");
    loc2d_pdf_kernel<<<(unsigned)Nev, 256>>>((int)Ns, (int)P, dra.as<double>(), ddec.as<double>(), drp.as<double>(),
                                             ddp.as<double>(), dpdf.as<double>());
    SCU(cudaGetLastError(), "loc2d_pdf launch");
    SCU(cudaMemcpy(gw_loc2d_pdf, dpdf.p, np * sizeof(double), cudaMemcpyDeviceToHost), "D2H pdf");
  }
  return CHB_OK;
}

int chb_precompute_p_cat(int device, int64_t Nev, int64_t P, int64_t Nz, const double* z_grids, const double* dVdz,
                         const int64_t* opt_nsides, const int64_t* pixels_opt_nsides, const int32_t* neff_pixels,
                         int64_t Ngal, const double* gal_ra, const double* gal_dec, const double* gal_z,
                         const double* gal_zerr, const double* gal_w, double* p_cat, double* N_gal) {
  if (Nev < 1 || P < 1 || Nz < 2 || !z_grids || !dVdz || !opt_nsides || !pixels_opt_nsides || !neff_pixels || !p_cat)
    return chb_fail_global(CHB_ERR_INVALID, "bad catalogue arguments");
  if (Ngal < 0 || (Ngal > 0 && (!gal_ra || !gal_dec || !gal_z || !gal_zerr || !gal_w)))
    return chb_fail_global(CHB_ERR_INVALID, "bad galaxy arrays");
  std::map<int64_t, std::vector<int>> by_nside;
  for (int64_t e = 0; e < Nev; ++e) {
    if (!pow2(opt_nsides[e])) return chb_fail_global(CHB_ERR_INVALID, "nside must be a positive power of 2");
    if (neff_pixels[e] < 0 || neff_pixels[e] > P) return chb_fail_global(CHB_ERR_INVALID, "neff_pixels out of range");
    by_nside[opt_nsides[e]].push_back((int)e);
  }
  DevGuard _dg;
  int rc = pick_device(device);
  if (rc != CHB_OK) return rc;
  const size_t npz = (size_t)Nev * P * Nz;
  Buf dzg, ddv, dpx, dnp, dra, ddec, dz, dze, dw, dkeys, dperm, dsz, dsze, dsw, dsel, dpc, dng;
  SCU(dzg.up(z_grids, (size_t)Nev * Nz * sizeof(double)), "upload z_grids");
  SCU(ddv.up(dVdz, (size_t)Nev * Nz * sizeof(double)), "upload dVdz");
  SCU(dpx.up(pixels_opt_nsides, (size_t)Nev * P * sizeof(long long)), "upload pixels");
  SCU(dnp.up(neff_pixels, Nev * sizeof(int)), "upload neff_pixels");
  const size_t gb = (size_t)std::max<int64_t>(Ngal, 1) * sizeof(double);
  if (Ngal > 0) {
    SCU(dra.up(gal_ra, gb), "upload gal ra"); SCU(ddec.up(gal_dec, gb), "upload gal dec");
    SCU(dz.up(gal_z, gb), "upload gal z"); SCU(dze.up(gal_zerr, gb), "upload gal z_err"); SCU(dw.up(gal_w, gb), "upload gal w");
  }
  SCU(dkeys.alloc(gb), "alloc keys"); SCU(dperm.alloc(gb), "alloc perm");
  SCU(dsz.alloc(gb), "alloc sorted z"); SCU(dsze.alloc(gb), "alloc sorted z_err"); SCU(dsw.alloc(gb), "alloc sorted w");
  SCU(dpc.alloc(npz * sizeof(double)), "alloc p_cat");
  SCU(dng.alloc(Nev * sizeof(double)), "alloc N_gal");
  SCU(cudaMemset(dng.p, 0, Nev * sizeof(double)), "zero N_gal");
  {
    std::vector<double> pad(npz, -100.0);                              // catalog.py:148 padding
    SCU(cudaMemcpy(dpc.p, pad.data(), npz * sizeof(double), cudaMemcpyHostToDevice), "fill p_cat");
  }
  const size_t smem = (size_t)(2 + 2 * 4) * Nz * sizeof(double);
  if (smem > 200 * 1024) return chb_fail_global(CHB_ERR_UNSUPPORTED, "z grid too long for the p_cat kernel");
  SCU(cudaFuncSetAttribute(p_cat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
  for (auto& kv : by_nside) {
    const int64_t nside = kv.first;
    if (Ngal > 0) {
      radec2pix_kernel<<<grid_for(Ngal), 256>>>(nside, Ngal, dra.as<double>(), ddec.as<double>(), dkeys.as<long long>());
      SCU(cudaGetLastError(), "radec2pix launch");
      try {
        thrust::device_ptr<long long> k(dkeys.as<long long>()), p(dperm.as<long long>());
        thrust::sequence(thrust::device, p, p + Ngal);
        thrust::stable_sort_by_key(thrust::device, k, k + Ngal, p);      // bucketing by pixel, original order kept
      } catch (const std::exception& ex) {
        return chb_fail_global(CHB_ERR_CUDA, ex.what());
      }
      gather3_kernel<<<grid_for(Ngal), 256>>>(Ngal, dperm.as<long long>(), dz.as<double>(), dze.as<double>(), dw.as<double>(),
                                              dsz.as<double>(), dsze.as<double>(), dsw.as<double>());
      SCU(cudaGetLastError(), "gather launch");
    }
    SCU(dsel.up(kv.second.data(), kv.second.size() * sizeof(int)), "upload event list");
    dim3 grid((unsigned)P, (unsigned)kv.second.size());
    p_cat_kernel<<<grid, 128, smem>>>((int)Nz, (int)P, (int)kv.second.size(), dsel.as<int>(), dpx.as<long long>(),
                                      dnp.as<int>(), dzg.as<double>(), ddv.as<double>(), Ngal, dkeys.as<long long>(),
                                      dsz.as<double>(), dsze.as<double>(), dsw.as<double>(), dpc.as<double>(),
                                      dng.as<double>());
    SCU(cudaGetLastError(), "p_cat launch");
    SCU(cudaDeviceSynchronize(), "p_cat kernel");
  }
  SCU(cudaMemcpy(p_cat, dpc.p, npz * sizeof(double), cudaMemcpyDeviceToHost), "D2H p_cat");
  if (N_gal) SCU(cudaMemcpy(N_gal, dng.p, Nev * sizeof(double), cudaMemcpyDeviceToHost), "D2H N_gal");
  return CHB_OK;
}

}  // extern "C"
