// ubench.cu -- B200 microbenchmarks that set the roofline denominators the table fill is judged
// against and the latencies its wavefront was designed around:
//   FP64 DFMA peak (MEASURED_PEAKS.json has no FP64 entry), DFMA / SHFL / LDS dependent latency,
//   HBM write-only bandwidth (the fill writes and never reads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_latency(double *out, long long *cycles, double a, double b, int iters) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; i++) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void shfl_latency(int *out, long long *cycles, int iters) {
  int x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; i++) x = __shfl_up_sync(0xffffffffu, x, 1) + 1;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void lds_latency(int *out, long long *cycles, int iters) {
  __shared__ int buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; i++) x = buf[x];
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

// 8 independent DFMA chains per thread
__global__ void dfma_throughput(double *out, double a, double b, int iters) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void write_bw(double2 *dst, size_t n16, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double2 val = make_double2(v, v + 1.0);
  for (; i < n16; i += stride) dst[i] = val;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d", p.name, p.multiProcessorCount, clk_khz / 1000);
  double *dout; long long *dcyc; int *iout;
  CK(cudaMalloc(&dout, 1 << 24)); CK(cudaMalloc(&dcyc, 64)); CK(cudaMalloc(&iout, 1 << 16));
  long long cyc;
  const int it = 1 << 16;
  for (int rep = 0; rep < 2; rep++) { dfma_latency<<<1, 32>>>(dout, dcyc, 1.0000001, 0.9999999, it); }
  CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  printf(", \"dfma_latency_cycles\": %.2f", (double)cyc / it);
  for (int rep = 0; rep < 2; rep++) shfl_latency<<<1, 32>>>(iout, dcyc, it);
  CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  printf(", \"shfl_plus_iadd_latency_cycles\": %.2f", (double)cyc / it);
  for (int rep = 0; rep < 2; rep++) lds_latency<<<1, 32>>>(iout, dcyc, it);
  CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  printf(", \"lds_latency_cycles\": %.2f", (double)cyc / it);

  // FP64 throughput: all SMs, 1024 threads/SM x 2 CTAs
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int blocks = p.multiProcessorCount * 4, threads = 512, iters = 1 << 15;
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    dfma_throughput<<<blocks, threads>>>(dout, 1.0000001, 0.9999999, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double fmas = (double)blocks * threads * iters * 8.0;
  printf(", \"dfma_per_s\": %.4g, \"fp64_tflops\": %.3f", fmas / (best * 1e-3), 2 * fmas / (best * 1e-3) / 1e12);
  // sustained (2 s)
  {
    int n = 0; CK(cudaEventRecord(e0));
    float ms = 0;
    do { for (int k = 0; k < 10; k++) dfma_throughput<<<blocks, threads>>>(dout, 1.0000001, 0.9999999, iters); n += 10;
         CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000);
    printf(", \"fp64_tflops_sustained\": %.3f", 2 * fmas * n / (ms * 1e-3) / 1e12);
  }
  // HBM write-only bandwidth, 16 GiB buffer
  size_t bytes = (size_t)16 << 30;
  double2 *big; CK(cudaMalloc(&big, bytes));
  best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    write_bw<<<p.multiProcessorCount * 8, 512>>>(big, bytes / 16, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  printf(", \"hbm_write_gbs\": %.1f", bytes / (best * 1e-3) / 1e9);
  best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    CK(cudaMemsetAsync(big, 0, bytes));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  printf(", \"memset_gbs\": %.1f}\n", bytes / (best * 1e-3) / 1e9);
  return 0;
}
