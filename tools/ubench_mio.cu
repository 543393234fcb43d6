// ubench_mio.cu -- issue interval of shared-memory stores and shuffles from ONE warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mio tools/ubench_mio.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, int stride) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31;
  double *p = sm + (threadIdx.x >> 5) * 4096 + lane * stride;
  double v = lane, w = lane + 1;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) p[i * 320] = v;                                                    // STS.64
      if (MODE == 1) *reinterpret_cast<double2 *>(p + i * 320) = make_double2(v, w);   // STS.128
      if (MODE == 2) v = __shfl_up_sync(0xffffffffu, v, 1) + 1.0;                       // 2 SHFL + DADD (dependent)
      if (MODE == 3) { int a = __shfl_up_sync(0xffffffffu, __double2loint(v), 1); w += (double)a; }  // SHFL independent-ish
    }
    v += 1.0;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + w + p[0];
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps, int stride) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 32 * 8);
  const int iters = 20000;
  size_t smem = warps * 4096 * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<148, warps * 32, smem>>>(out, cyc, iters, stride);
  k<MODE><<<148, warps * 32, smem>>>(out, cyc, iters, stride);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-22s warps/SM=%d lane stride %d doubles: %.2f cycles per op per warp (%s)\n", name, warps, stride,
         (double)h / iters / 8.0, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("STS.64", 1, 1); run<0>("STS.64", 1, 5); run<0>("STS.64", 4, 5);
  run<1>("STS.128", 1, 2); run<1>("STS.128", 1, 6); run<1>("STS.128", 1, 4); run<1>("STS.128", 4, 6);
  run<2>("SHFL x2 dependent", 1, 1);
  run<3>("SHFL", 1, 1);
  return 0;
}
