/*
 * dev_guard.cuh -- entry points run on their handle's device and give the caller's current device back
 * on every return path (the reference is a CPU library: a caller does not expect a table call to move
 * its CUDA context; and the multi-device entry points drive several devices from one process).
 */
#pragma once
#include <cuda_runtime.h>

namespace stb {

struct DeviceGuard {
  int prev;
  cudaError_t err;
  explicit DeviceGuard(int dev) : prev(-1), err(cudaSuccess) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};

}  // namespace stb
