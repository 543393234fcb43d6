// ubench_contend.cu -- what slows the strip producer's step inside the full kernel?  Warp 0 runs
// strip_steps<5> (as in ubench_steps.cu); the other warps of the CTA (on the OTHER sub-partitions
// unless SAME=1) run one kind of background work: FP64 arithmetic, conflict-free shared loads,
// conflicting shared loads, global stores, or sleep-polling.  Development aid for fill_strip.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libstb_b200/csrc -o tools/ubench_contend tools/ubench_contend.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "fill_strip.cuh"
using namespace stb;

template <int BG>
__global__ void __launch_bounds__(512, 1) kern(long long *cycles, double *sink, double *gout, int batches, double a, int nbg, int same) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int K = 5, CP = 32 * K, RS = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *xring = reinterpret_cast<double *>(smem_raw);
  double *yring = xring + (RS + 8) * CP;
  double *outx = yring + (RS + 8) * 32;
  double *bgmem = outx + 64;  // 8192 doubles of background shared memory
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (RS + 8) * CP + (RS + 8) * 32 + 64 + 8192; i += blockDim.x) xring[i] = 1.0 + i;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    double x[K], ma[K];
    for (int k = 0; k < K; k++) { x[k] = 0.0; ma[k] = (double)(1 + lane * K + k) * a; }
    double nm1 = (double)(-lane), yin = lane == 0 ? 1.0 : 0.0;
    long long E = 0;
    const unsigned a_xr = smem_u32(xring + lane * K), a_yr = smem_u32(yring + lane), a_out = smem_u32(outx);
    const unsigned nb_stride = lane == 0 ? 8u : (unsigned)(CP * 8);
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
      strip_steps<K, false, false, CP, RS, true>(x, ma, nm1, yin, lane == 0 ? a_out : a_xr + 15 * CP * 8 - 8u, scn, scn, nb_addr, nb_stride, lane == 31, a_xr, a_yr, a_out);
      strip_steps<K, false, false, CP, RS, false>(x, ma, nm1, yin, 0u, scn, scn, nb_addr, nb_stride, lane == 31, a_xr + 8 * CP * 8,
                                                  a_yr, a_out + 64);
    }
    long long t1 = clock64();
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
    double s = 0;
    for (int k = 0; k < K; k++) s += x[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)E + yin;
    return;
  }
  // background
  if (!same && (warp & 3) == 0) return;
  int rank = 0;
  for (int w = 1; w < warp; w++) if (same || (w & 3) != 0) rank++;
  if (rank >= nbg) return;
  double acc[8];
  for (int i = 0; i < 8; i++) acc[i] = lane + i;
  unsigned h = threadIdx.x * 2654435761u;
  double *g = gout + ((size_t)blockIdx.x * 16 + warp) * (1 << 16);
  size_t gi = lane;
  long long iters = 0;
  while (!stop) {
    if (BG == 1) {  // FP64
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], 0.999, 0.5);
    } else if (BG == 2) {  // conflict-free LDS.64
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += ((volatile double *)bgmem)[(i * 32 + lane + (iters & 7) * 256) & 8191];
    } else if (BG == 3) {  // random 16-byte LDS (conflicts)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        h = h * 1664525u + 1013904223u;
        double vx, vy;
        asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(smem_u32(bgmem) + ((h >> 12) & 255) * 16));
        acc[i] += vx + vy;
      }
    } else if (BG == 4) {  // coalesced global stores
#pragma unroll
      for (int i = 0; i < 8; i++) { g[gi & 65535] = acc[i]; gi += 32; }
    } else if (BG == 5) {  // sleep-poll
      __nanosleep(40);
    } else if (BG == 6) {  // integer ALU
#pragma unroll
      for (int i = 0; i < 32; i++) h = h * 1664525u + 1013904223u;
    }
    iters++;
  }
  double s = h;
  for (int i = 0; i < 8; i++) s += acc[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + iters;
}

template <int BG>
void run(const char *name, int nbg, int same, long long *cyc, double *sink, double *gout) {
  const int batches = 4000;
  size_t smem = ((16 + 8) * 160 + (16 + 8) * 32 + 64 + 8192) * 8;
  cudaFuncSetAttribute(kern<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) kern<BG><<<148, 512, smem>>>(cyc, sink, gout, batches, 0.7, nbg, same);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s bg warps %2d %s: %.1f cycles/step\n", name, nbg, same ? "(any sub-partition)" : "(other sub-partitions)", (double)h / batches / 16.0);
}

int main() {
  long long *cyc; double *sink, *gout;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 8); cudaMalloc(&gout, (size_t)148 * 16 * (1 << 16) * 8);
  run<0>("none", 0, 0, cyc, sink, gout);
  for (int n : {3, 6, 12}) run<1>("FP64", n, 0, cyc, sink, gout);
  for (int n : {3, 6, 12}) run<2>("LDS.64 conflict-free", n, 0, cyc, sink, gout);
  for (int n : {3, 6, 12}) run<3>("LDS.128 random", n, 0, cyc, sink, gout);
  for (int n : {3, 6, 12}) run<4>("STG coalesced", n, 0, cyc, sink, gout);
  for (int n : {3, 12}) run<5>("nanosleep poll", n, 0, cyc, sink, gout);
  for (int n : {3, 12}) run<6>("int ALU", n, 0, cyc, sink, gout);
  run<1>("FP64", 15, 1, cyc, sink, gout);
  run<6>("int ALU", 15, 1, cyc, sink, gout);
  run<5>("nanosleep poll", 15, 1, cyc, sink, gout);
  return 0;
}
