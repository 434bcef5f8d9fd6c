// devguard.cuh -- RAII guard for the calling thread's current CUDA device.
#pragma once
#include <cuda_runtime.h>

// Entry points run on the handle's (or the requested) device and leave the calling thread's current CUDA device as they
// found it: the caller may be a torch process whose current device carries meaning (torch.cuda.current_device()).
struct DevGuard {
  int prev = -1;
  DevGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  explicit DevGuard(int dev) : DevGuard() { cudaSetDevice(dev); }
  // (restores only when the device actually changed: cudaSetDevice would otherwise create a primary context on `prev`)
  ~DevGuard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};
