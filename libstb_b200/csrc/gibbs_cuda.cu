/*
 * gibbs_cuda.cu -- a device-side consumer of the V table: the table-indicator Gibbs step of a
 * Pitman-Yor / Dirichlet-process mixture, many restaurants per launch.
 *
 * What it replaces: the per-token loop of the reference's demo (test/demo.c:405-434), whose inner
 * operation is one scalar S_V look-up per token on the host.  Restaurant j holds n[j][i] customers
 * eating dish i at t[j][i] tables (1 <= t <= n), T[j] = sum_i t[j][i]; a sweep visits the restaurant's
 * tokens in order and, for a token of dish i with n > 1,
 *   - removes its table indicator with probability (t - 1) / (n - 1)      (only drawn when t > 1),
 *   - adds one with probability one / (one + 1),
 *         one = H_i (b + T_j a) t / (n - t + 1) V^n_{t+1}                  (rounded to float, as there).
 * Restaurants are independent given (a, b, H): one thread per restaurant, each on its own 48-bit
 * stream, reading V straight from the table's slab in HBM / L2 -- the table never crosses PCIe.
 * shared_stream != 0 is the reference's schedule instead: ONE thread visits the restaurants in order
 * on one stream, which reproduces demo.c's draws uniform for uniform (the parity mode).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dev_guard.cuh"
#include "rng48.h"
#include "stb_cuda.h"

extern "C" void stb_cuda_set_error(const char *what, int code);

#define GCK(call)                          \
  do {                                     \
    cudaError_t e_ = (call);               \
    if (e_ != cudaSuccess) {               \
      stb_cuda_set_error(#call, (int)e_);  \
      rc = (int)e_ ? (int)e_ : -1;         \
      goto done;                           \
    }                                      \
  } while (0)

/* S_V inside the filled table (lib/stable.c:927-928: 0 for m < 2 or n < m); 0 beyond it, like a look-up beyond
 * the table's maximum extent (:921) -- the C layer grows the table to cover the counts before the launch */
template <typename T>
__device__ __forceinline__ double v_cell(const T *__restrict__ tab, size_t ld, unsigned usedN, unsigned usedM, unsigned n,
                                         unsigned m) {
  if (m < 2 || n < m || n > usedN || m > usedM) return 0.0;
  return (double)tab[(size_t)(n - 1) * ld + (m - 1)];
}

template <typename T>
__device__ void ti_restaurant(const T *__restrict__ tab, size_t ld, unsigned usedN, unsigned usedM, double apar, double bpar,
                              const uint32_t *__restrict__ dish, uint32_t ntok, const float *__restrict__ H,
                              const uint32_t *__restrict__ n, uint16_t *__restrict__ t, uint32_t &Tj, stb_rng48 &r) {
  for (uint32_t c = 0; c < ntok; c++) {
    const uint32_t i = dish[c];
    const uint32_t ni = n[i];
    uint32_t ti = t[i];
    if (ni <= 1) continue;  // a single customer: the indicator is always 1
    if (ti > 1 && (double)(ni - 1) * stb_rng48_unit(&r) < (double)(ti - 1)) {
      ti--;
      Tj--;
    }
    // the arithmetic of test/demo.c:427-429, operation by operation: double products left to right, the
    // result rounded to float, the acceptance ratio formed in double from the float
    const double v = v_cell(tab, ld, usedN, usedM, ni, ti + 1);
    const float one = (float)__dmul_rn(
        __ddiv_rn(__dmul_rn(__dmul_rn((double)H[i], __dadd_rn(bpar, __dmul_rn((double)Tj, apar))), (double)ti),
                  (double)(ni - ti + 1)),
        v);
    if (stb_rng48_unit(&r) < __ddiv_rn((double)one, __dadd_rn((double)one, 1.0))) {
      ti++;
      Tj++;
    }
    t[i] = (uint16_t)ti;
  }
}

template <typename T>
__global__ void ti_gibbs_kernel(const T *__restrict__ tab, size_t ld, unsigned usedN, unsigned usedM, double apar, double bpar,
                                uint32_t R, const uint32_t *__restrict__ tok_off, const uint32_t *__restrict__ tok_dish,
                                const float *__restrict__ H, uint32_t D, const uint32_t *__restrict__ n,
                                uint16_t *__restrict__ t, uint32_t *__restrict__ Tsum, unsigned long long *__restrict__ rng,
                                int shared_stream, int sweeps) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (shared_stream) {
    if (j != 0) return;
    stb_rng48 r;
    r.x = rng[0];
    for (int s = 0; s < sweeps; s++)
      for (uint32_t q = 0; q < R; q++) {
        uint32_t Tq = Tsum[q];
        ti_restaurant(tab, ld, usedN, usedM, apar, bpar, tok_dish + tok_off[q], tok_off[q + 1] - tok_off[q], H,
                      n + (size_t)q * D, t + (size_t)q * D, Tq, r);
        Tsum[q] = Tq;
      }
    rng[0] = r.x;
    return;
  }
  if (j >= R) return;
  stb_rng48 r;
  r.x = rng[j];
  uint32_t Tj = Tsum[j];
  for (int s = 0; s < sweeps; s++)
    ti_restaurant(tab, ld, usedN, usedM, apar, bpar, tok_dish + tok_off[j], tok_off[j + 1] - tok_off[j], H, n + (size_t)j * D,
                  t + (size_t)j * D, Tj, r);
  Tsum[j] = Tj;
  rng[j] = r.x;
}

extern "C" int stb_cuda_ti_gibbs(stb_dev_t *d, unsigned usedN, unsigned usedM, double apar, double bpar, size_t R,
                                 const uint32_t *tok_off, const uint32_t *tok_dish, const float *H, uint32_t D,
                                 const uint32_t *n, uint16_t *t, uint32_t *T, uint64_t *rng, int shared_stream, int sweeps,
                                 float *ms) {
  const void *tabV = stb_cuda_table_ptr(d, STB_TAB_V);
  const size_t ld = stb_cuda_table_ld(d);
  const int is_float = stb_cuda_table_is_float(d);
  stb::DeviceGuard guard(stb_cuda_table_device(d));
  int rc = 0;
  const size_t ntok = tok_off[R], nstreams = shared_stream ? 1 : R;
  uint32_t *d_off = NULL, *d_dish = NULL, *d_n = NULL, *d_T = NULL;
  uint16_t *d_t = NULL;
  float *d_H = NULL;
  unsigned long long *d_rng = NULL;
  cudaEvent_t ev0 = NULL, ev1 = NULL;
  if (guard.err != cudaSuccess) {
    stb_cuda_set_error("cudaSetDevice", (int)guard.err);
    return (int)guard.err;
  }
  if (!tabV) {
    stb_cuda_set_error("stb_cuda_ti_gibbs: the table has no V slab", 0);
    return -1;
  }
  GCK(cudaMalloc(&d_off, (R + 1) * sizeof(uint32_t)));
  GCK(cudaMalloc(&d_dish, (ntok ? ntok : 1) * sizeof(uint32_t)));
  GCK(cudaMalloc(&d_H, (size_t)D * sizeof(float)));
  GCK(cudaMalloc(&d_n, R * (size_t)D * sizeof(uint32_t)));
  GCK(cudaMalloc(&d_t, R * (size_t)D * sizeof(uint16_t)));
  GCK(cudaMalloc(&d_T, R * sizeof(uint32_t)));
  GCK(cudaMalloc(&d_rng, nstreams * sizeof(unsigned long long)));
  GCK(cudaMemcpy(d_off, tok_off, (R + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_dish, tok_dish, ntok * sizeof(uint32_t), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_H, H, (size_t)D * sizeof(float), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_n, n, R * (size_t)D * sizeof(uint32_t), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_t, t, R * (size_t)D * sizeof(uint16_t), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_T, T, R * sizeof(uint32_t), cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(d_rng, rng, nstreams * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  GCK(cudaEventCreate(&ev0));
  GCK(cudaEventCreate(&ev1));
  GCK(cudaEventRecord(ev0, 0));
  {
    // restaurants differ in length: small blocks spread the long ones over the SMs
    const unsigned threads = shared_stream ? 32 : 64, blocks = shared_stream ? 1 : (unsigned)((R + threads - 1) / threads);
    if (is_float)
      ti_gibbs_kernel<float><<<blocks, threads>>>((const float *)tabV, ld, usedN, usedM, apar, bpar, (uint32_t)R, d_off, d_dish,
                                                 d_H, D, d_n, d_t, d_T, d_rng, shared_stream, sweeps);
    else
      ti_gibbs_kernel<double><<<blocks, threads>>>((const double *)tabV, ld, usedN, usedM, apar, bpar, (uint32_t)R, d_off,
                                                   d_dish, d_H, D, d_n, d_t, d_T, d_rng, shared_stream, sweeps);
  }
  GCK(cudaGetLastError());
  GCK(cudaEventRecord(ev1, 0));
  GCK(cudaMemcpy(t, d_t, R * (size_t)D * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  GCK(cudaMemcpy(T, d_T, R * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  GCK(cudaMemcpy(rng, d_rng, nstreams * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (ms) GCK(cudaEventElapsedTime(ms, ev0, ev1));
done:
  cudaFree(d_off);
  cudaFree(d_dish);
  cudaFree(d_H);
  cudaFree(d_n);
  cudaFree(d_t);
  cudaFree(d_T);
  cudaFree(d_rng);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  return rc;
}
