// microbench.cu -- measured denominators for the roofline report.
// chb_mufu_peak: sustained MUFU.EX2 throughput of the device (the bound of the Gaussian KDE pair
// sum: one ex2 per pair).  8 independent chains per thread, 2048 threads per SM resident.
#include "common.cuh"
#include <cuda_runtime.h>
#include "../../include/chimera_b200.h"

__global__ void __launch_bounds__(256) mufu_kernel(float* out, int iters) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = 0.001f * (float)(threadIdx.x + 1) + 0.1f * k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[k]) : "f"(-x[k]));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;   // never true; keeps the chains alive
}

extern "C" int chb_mufu_peak(int device, double seconds, double* exp_per_s) {
  if (!exp_per_s) return CHB_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return CHB_ERR_CUDA; }
  DevGuard _dg(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  float* d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return CHB_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * 8, block = 256, iters = 8192;
  mufu_kernel<<<grid, block>>>(d, 64);                       // warm-up
  cudaDeviceSynchronize();
  double best = 0.0, spent = 0.0;
  if (seconds <= 0) seconds = 0.05;
  while (spent < seconds) {
    cudaEventRecord(e0);
    mufu_kernel<<<grid, block>>>(d, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return CHB_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)grid * block * iters * 8.0 / (ms * 1e-3);
    if (rate > best) best = rate;
    spent += ms * 1e-3;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *exp_per_s = best;
  return CHB_OK;
}
