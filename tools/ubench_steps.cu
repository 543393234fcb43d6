// ubench_steps.cu -- cycles per recurrence step of the strip producer's inner loop in isolation
// (no flow control, no consumers): how fast can ONE warp run strip_steps<K>, and how do several
// such warps on different SM sub-partitions add up.  Development aid for fill_strip.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libstb_b200/csrc -o tools/ubench_steps tools/ubench_steps.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fill_strip.cuh"

using namespace stb;

template <int K, bool HAS_V>
__global__ void steps_kernel(long long *cycles, double *sink, int batches, double a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CP = 32 * K, RS = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *xring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * ((RS + 8) * CP + (RS + 8) * 32 + 64);
  double *yring = xring + (RS + 8) * CP;
  double *outx = yring + (RS + 8) * 32;
  double x[K], ma[K], bnd[ST_B];
  for (int k = 0; k < K; k++) { x[k] = 0.0; ma[k] = (double)(1 + lane * K + k) * a; }
  for (int i = 0; i < ST_B; i++) bnd[i] = 0.0;
  double nm1 = (double)(-lane), yin = lane == 0 ? 1.0 : 0.0;
  long long E = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int p = 0; p < batches; p++) {
    const int hi = __double2hiint(x[0]);
    int e = ((hi >> 20) & 0x7ff) - 1023;
    if (x[0] == 0.0) e = 0;
    const double sc = pow2i(-e);
    for (int k = 0; k < K; k++) x[k] *= sc;
    yin *= sc;
    E += e;
    int elow = (int)E;
    int sE = __shfl_up_sync(0xffffffffu, elow, 1);
    if (lane == 0) sE = elow;
    const double scn = pow2i(sE - elow);
    const int slot = p & 1;
    strip_steps<K, HAS_V, false, CP, RS, ST_RB>(x, ma, nm1, yin, scn, bnd, lane == 0, lane == 31,
                                                xring + slot * 8 * CP + lane * K, yring + slot * 8 * 32 + lane, outx);
    strip_steps<K, HAS_V, false, CP, RS, ST_RB>(x, ma, nm1, yin, scn, bnd + ST_RB, lane == 0, lane == 31,
                                                xring + slot * 8 * CP + lane * K, yring + slot * 8 * 32 + lane, outx + ST_RB);
  }
  long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  double s = 0;
  for (int k = 0; k < K; k++) s += x[k];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)E;
}

template <int K, bool HAS_V>
void run(int warps, int batches) {
  long long *cyc;
  double *sink;
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  size_t smem = (size_t)warps * ((16 + 8) * 32 * K + (16 + 8) * 32 + 64) * 8;
  cudaFuncSetAttribute(steps_kernel<K, HAS_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) steps_kernel<K, HAS_V><<<148, warps * 32, smem>>>(cyc, sink, batches, 0.7);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[32];
  cudaMemcpy(h, cyc, warps * sizeof(long long), cudaMemcpyDeviceToHost);
  double per_step = (double)h[0] / batches / (double)ST_B;
  printf("K=%d V=%d warps/SM=%d : %.1f cycles/step/warp  -> %.2f cycles per cell per SM\n", K, (int)HAS_V, warps, per_step,
         per_step / (32.0 * K * warps));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  const int B = 20000;
  run<1, false>(1, B); run<1, false>(4, B); run<1, false>(8, B);
  run<2, false>(1, B); run<2, false>(4, B);
  run<3, false>(1, B); run<3, false>(2, B); run<3, false>(4, B);
  run<5, false>(1, B); run<5, false>(2, B); run<5, false>(4, B);
  run<7, false>(1, B);
  run<5, true>(1, B);
  return 0;
}
