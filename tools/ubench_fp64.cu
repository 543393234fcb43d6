// ubench_fp64.cu -- FP64 issue interval of ONE warp (independent DFMA / DADD streams) and how it adds up
// over warps of the same SM sub-partition.  Development aid for the producer of fill_strip.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, double a, double b, int iters) {
  double x[16];
  unsigned y[16];
  for (int i = 0; i < 16; i++) x[i] = a + i, y[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (MODE == 0) x[i] = fma(x[i], b, a);
      if (MODE == 1) x[i] = x[i] + b;
      if (MODE == 2) x[i] = (i & 1) ? fma(x[i], b, a) : x[i] + b;
      if (MODE == 5) { x[i] = fma(x[i], b, a); y[i] = y[i] * 1664525u + 1013904223u; }  // one DFMA + one IMAD, independent
      if (MODE == 6) { y[i] = y[i] * 1664525u + 1013904223u; }                            // the IMAD alone
      if (MODE == 3) x[i] = (double)(__double2loint(x[i]) + it + i);  // I2F.F64.S32 (integer -> double conversion)
      if (MODE == 4) x[i] = __hiloint2double(0x43300000, __double2loint(x[i]) + it + i) - 4503601774854144.0;  // magic-number conversion: IADD + MOV + DADD
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += x[i] + (double)y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 32 * 8);
  const int iters = 20000;
  k<MODE><<<148, warps * 32>>>(out, cyc, 1.0, 0.999999, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, 1.0, 0.999999, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-10s warps/SM=%2d (%d per sub-partition): %.2f cycles per FP64 instruction per warp, %.1f lanes/clk/SM\n", name,
         warps, (warps + 3) / 4, (double)h / iters / 16.0, 32.0 * 16 * iters * warps / (double)h);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) run<0>("DFMA", w);
  for (int w : {1, 4, 8, 16}) run<1>("DADD", w);
  for (int w : {1, 4, 16}) run<2>("DFMA+DADD", w);
  for (int w : {1, 4, 16}) run<5>("DFMA+IMAD pair", w);
  for (int w : {1, 4, 16}) run<6>("IMAD alone", w);
  for (int w : {1, 4, 16}) run<3>("I2F.F64", w);
  for (int w : {1, 4, 16}) run<4>("magic i2d", w);
  return 0;
}
