// ubench_steps.cu -- cycles per recurrence step of the strip producer's inner loop in isolation
// (no flow control, no consumers): how fast can ONE warp run strip_steps<K>, and which part of a
// step costs what (variants drop the stores / the neighbour load / the coefficient adds).
// Development aid for fill_strip.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libstb_b200/csrc -o tools/ubench_steps tools/ubench_steps.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fill_strip.cuh"

using namespace stb;

// local variant of the step: MODE bit 0 = no x stores, bit 1 = no neighbour load, bit 2 = no coefficient adds
template <int K, int CP, int MODE>
__device__ __forceinline__ void steps_var(double (&x)[K], const double (&ma)[K], double &nm1, double &yin, const double scn,
                                          unsigned &nb_addr, const unsigned nb_stride, const unsigned xr) {
  double xp[K];
#pragma unroll
  for (int i = 0; i < ST_RB; i++) {
    double nb = 1.0;
    if (!(MODE & 2)) {
      nb = lds_f64(nb_addr);
      nb_addr += nb_stride;
    }
#pragma unroll
    for (int k = K - 1; k >= 1; k--) x[k] = fma((MODE & 4) ? ma[k] : nm1 - ma[k], x[k], x[k - 1]);
    x[0] = fma((MODE & 4) ? ma[0] : nm1 - ma[0], x[0], yin);
    nm1 += 1.0;
    if (!(MODE & 1)) {
      if (MODE & 8) {  // 16-byte stores, K even
#pragma unroll
        for (int k = 0; k < K; k += 2)
          asm volatile("st.volatile.shared.v2.f64 [%0], {%1,%2};" ::"r"(xr + (i * CP + k) * 8), "d"(x[k]), "d"(x[k + 1 < K ? k + 1 : k]));
      } else if (MODE & 16) {  // columns 0..K-2: one 16-byte store per column and PAIR of rows; column K-1 every row
        sts_f64(xr + (i * CP + K - 1) * 8, x[K - 1]);
        if (i & 1) {
#pragma unroll
          for (int k = 0; k < K - 1; k++)
            asm volatile("st.volatile.shared.v2.f64 [%0], {%1,%2};" ::"r"(xr + ((i - 1) * CP + 2 * k) * 8 + 16 * 1024), "d"(xp[k]), "d"(x[k]));
        } else {
#pragma unroll
          for (int k = 0; k < K - 1; k++) xp[k] = x[k];
        }
      } else if (MODE & 32) {  // weak (non-volatile) stores
#pragma unroll
        for (int k = 0; k < K; k++) asm volatile("st.shared.f64 [%0], %1;" ::"r"(xr + (i * CP + k) * 8), "d"(x[k]) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < K; k++) sts_f64(xr + (i * CP + k) * 8, x[k]);
      }
      if (MODE & 64) sts_f64_if(xr + 64 * 1024 + i * 8, x[K - 1], (threadIdx.x & 31) == 31);  // the boundary column's extra store
    }
    yin = nb * scn;
  }
}

template <int K, int MODE>
__global__ void steps_kernel(long long *cycles, double *sink, int batches, double a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CP = 32 * K, RS = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *xring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * ((RS + 8) * CP + (RS + 8) * 32 + 64);
  double *yring = xring + (RS + 8) * CP;
  double *outx = yring + (RS + 8) * 32;
  for (int i = lane; i < (RS + 8) * CP + (RS + 8) * 32 + 64; i += 32) xring[i] = 0.0;
  double x[K], ma[K];
  for (int k = 0; k < K; k++) { x[k] = 0.0; ma[k] = (double)(1 + lane * K + k) * a; }
  double nm1 = (double)(-lane), yin = lane == 0 ? 1.0 : 0.0, nbp = 0.0;
  long long E = 0;
  const unsigned a_xr = smem_u32(xring + lane * K), a_yr = smem_u32(yring + lane), a_out = smem_u32(outx);
  const unsigned nb_stride = lane == 0 ? 8u : (unsigned)(CP * 8);
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
    const double scn = lane == 0 ? 0.0 : pow2i(sE - elow);
    unsigned nb_addr = lane == 0 ? a_out + 8u : a_xr - 8u;
    if (MODE == 0) {  // MODE 128: steps_var with everything on
      strip_steps<K, false, false, CP, RS, true, true>(x, ma, nm1, yin, nbp, lane == 0, lane == 0 ? a_out : a_xr + 15 * CP * 8 - 8u, scn, scn, nb_addr, nb_stride, lane == 31, a_xr, a_yr, a_out);
      strip_steps<K, false, false, CP, RS, false, true>(x, ma, nm1, yin, nbp, lane == 0, 0u, scn, scn, nb_addr, nb_stride, lane == 31, a_xr + 8 * CP * 8,
                                                  a_yr, a_out + 64);
    } else {
      steps_var<K, CP, MODE>(x, ma, nm1, yin, scn, nb_addr, nb_stride, a_xr);
      steps_var<K, CP, MODE>(x, ma, nm1, yin, scn, nb_addr, nb_stride, a_xr + 8 * CP * 8);
    }
  }
  long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  double s = 0;
  for (int k = 0; k < K; k++) s += x[k];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)E + yin;
}

template <int K, int MODE>
void run(int warps, int batches) {
  long long *cyc;
  double *sink;
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  size_t smem = (size_t)warps * ((16 + 8) * 32 * K + (16 + 8) * 32 + 64) * 8 + 96 * 1024;
  cudaFuncSetAttribute(steps_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) steps_kernel<K, MODE><<<148, warps * 32, smem>>>(cyc, sink, batches, 0.7);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[32];
  cudaMemcpy(h, cyc, warps * sizeof(long long), cudaMemcpyDeviceToHost);
  double per_step = (double)h[0] / batches / (double)ST_B;
  printf("K=%d mode=%d warps/SM=%d : %.1f cycles/step/warp  -> %.3f cycles per cell per SM\n", K, MODE, warps, per_step,
         per_step / (32.0 * K * warps));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  const int B = 20000;
  run<1, 0>(1, B); run<1, 0>(4, B);
  run<2, 0>(1, B); run<2, 0>(4, B);
  run<3, 0>(1, B); run<3, 0>(2, B); run<3, 0>(4, B);
  run<5, 0>(1, B); run<5, 0>(2, B); run<5, 0>(4, B);
  run<6, 0>(1, B); run<6, 0>(2, B); run<4, 0>(1, B);
  run<7, 0>(1, B);
  printf("-- variants (K=5, 1 warp): 1 = no stores, 2 = no neighbour load, 4 = no coefficient adds\n");
  run<5, 1>(1, B); run<5, 2>(1, B); run<5, 3>(1, B); run<5, 4>(1, B); run<5, 5>(1, B); run<5, 6>(1, B); run<5, 7>(1, B);
  printf("-- 16-byte stores (mode 8), K even; pair-of-rows 16-byte stores (mode 16)\n");
  run<4, 8>(1, B); run<6, 8>(1, B); run<4, 8>(2, B); run<6, 8>(2, B); run<8, 0>(1, B); run<8, 8>(1, B);
  printf("-- K=5: 0 = plain stores (no boundary store), 32 = weak stores, 64 = with the predicated boundary store\n");
  run<5, 128>(1, B); run<5, 32>(1, B); run<5, 64>(1, B); run<3, 128>(1, B); run<3, 64>(1, B); run<7, 128>(1, B); run<7, 64>(1, B);
  printf("-- variants (K=1, 1 warp)\n");
  run<1, 1>(1, B); run<1, 2>(1, B); run<1, 3>(1, B); run<1, 4>(1, B); run<1, 7>(1, B);
  return 0;
}
