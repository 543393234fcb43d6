/*
 * psample_cuda.cu -- device side of the batched samplers (psample.h, batched section).
 *
 * What runs here, per lock-step round over many chains:
 *   - the lgamma part of samplea's log-posterior (lib/samplea.c:65-67): for chain x,
 *       sum_i [ T_i log x + lgamma(T_i + b_i/x) - lgamma(b_i/x) ]
 *   - sampleb's log-posterior (lib/sampleb.c:33-41) and the digamma sum of its warm-up bmax
 *     (:59-62), both I-term reductions per chain;
 *   - sampleb's auxiliary variables (lib/sampleb.c:90-100): Q = 1/scale - sum_i log q_i with
 *     q_i ~ Beta(b, N_i) drawn from the chain's own 48-bit stream -- inherently serial per chain
 *     (rejection loops consume a data-dependent number of draws), so one thread per chain.
 * lgamma is CUDA's FP64 lgamma; digamma / trigamma are specfun.h.  Reductions are fixed-order trees.
 * The Stirling-table part of samplea's log-posterior is the discount sweep in stb_cuda.cu.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rng48.h"
#include "specfun.h"
#include "stb_cuda.h"
#include "dev_guard.cuh"

#include <mutex>

extern "C" const stb_zig_tables *stb_zig_tables_get(void);  // psample.c
extern "C" const char *stb_cuda_last_error(void);
extern "C" void stb_cuda_set_error(const char *what, int code);

struct stb_pstat_dev {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  int I;
  uint32_t *dT, *dN;     // [I]
  double *dbpar;         // [I] or [C][I]
  size_t bpar_elems, bpar_cap;
  stb_zig_tables *dzig;
  // per-call staging: x, aux1, aux2, out (doubles), chain (ints), rng (u64)
  size_t cap;
  double *dx, *da1, *da2, *dout;
  int *dchain;
  unsigned long long *drng;
  double *hbuf;  // pinned, 4*cap doubles
};

#define P_ON_DEVICE(dev)                                   \
  stb::DeviceGuard dev_guard_(dev);                        \
  if (dev_guard_.err != cudaSuccess) {                     \
    stb_cuda_set_error("cudaSetDevice", (int)dev_guard_.err); \
    return (int)dev_guard_.err;                            \
  }
#define PCK(call)                                        \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) {                             \
      stb_cuda_set_error(#call, (int)e_);                \
      return (int)e_ ? (int)e_ : -1;                     \
    }                                                    \
  } while (0)

__device__ __forceinline__ double block_sum_256(double v, double *red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  return red[0];
}

// one block per evaluation j: chain[j] selects the bpar row when bpar is per chain
__global__ void k_aterms_lg(const uint32_t *__restrict__ T, const double *__restrict__ bpar, int bpar_stride, int I,
                            const double *__restrict__ x, const int *__restrict__ chain, double *__restrict__ out) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  const double xv = x[j], lx = log(xv);
  const double *bp = bpar + (size_t)bpar_stride * (size_t)chain[j];
  double v = 0.0;
  for (int i = threadIdx.x; i < I; i += 256) {
    const double Ti = (double)T[i], bx = bp[i] / xv;
    v += Ti * lx + lgamma(Ti + bx) - lgamma(bx);
  }
  const double s = block_sum_256(v, red);
  if (threadIdx.x == 0) out[j] = s;
}

// sampleb's log-posterior: -Q x + (shape-1) log x + sum_i [lgamma(T_i + x/a) - lgamma(x/a)]
__global__ void k_bterms(const uint32_t *__restrict__ T, int I, const double *__restrict__ x,
                         const double *__restrict__ Q, const double *__restrict__ apar, double shape,
                         double *__restrict__ out) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  const double xa = x[j] / apar[j];
  const double lg = lgamma(xa);
  double v = 0.0;
  for (int i = threadIdx.x; i < I; i += 256) v += lgamma((double)T[i] + xa) - lg;
  const double s = block_sum_256(v, red);
  if (threadIdx.x == 0) out[j] = -Q[j] * x[j] + (shape - 1) * log(x[j]) + s;
}

// sum_i digamma(T_i + x/a)   (the fixed-point step of bmax, lib/sampleb.c:60-62)
__global__ void k_bdigamma(const uint32_t *__restrict__ T, int I, const double *__restrict__ x,
                           const double *__restrict__ apar, double *__restrict__ out) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  const double xa = x[j] / apar[j];
  double v = 0.0;
  for (int i = threadIdx.x; i < I; i += 256) v += stb_digamma((double)T[i] + xa);
  const double s = block_sum_256(v, red);
  if (threadIdx.x == 0) out[j] = s;
}

// Q_c = 1/scale - sum_i log Beta(b_c, N_i), serial per chain on the chain's own stream
__global__ void k_betaQ(const uint32_t *__restrict__ N, int I, const double *__restrict__ b_in,
                        unsigned long long *__restrict__ rng, double scale, const stb_zig_tables *__restrict__ zt,
                        double *__restrict__ Q, int C) {
  // ONE chain per warp, on lane 0: the draws are rejection loops of data-dependent length, and 32
  // chains in one warp would each pay for the longest of them at every step
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C || (threadIdx.x & 31)) return;
  stb_rng48 r;
  r.x = rng[c];
  double q = 1.0 / scale;
  const double b = b_in[c];
  bool bad = false;
  for (int i = 0; i < I; i++) {
    if (N[i] == 0) continue;
    const double v = stb_beta(&r, zt, b, (double)(int)N[i]);
    if (!(v > 0)) bad = true;
    q -= log(v);
  }
  rng[c] = r.x;
  Q[c] = bad ? nan("") : q;
}

static void pstat_free(stb_pstat_dev_t *p) {
  if (!p) return;
  stb::DeviceGuard dev_guard_(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  cudaFree(p->dT);
  cudaFree(p->dN);
  cudaFree(p->dbpar);
  cudaFree(p->dzig);
  cudaFree(p->dx);
  cudaFree(p->da1);
  cudaFree(p->da2);
  cudaFree(p->dout);
  cudaFree(p->dchain);
  cudaFree(p->drng);
  if (p->hbuf) cudaFreeHost(p->hbuf);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->stream) cudaStreamDestroy(p->stream);
  free(p);
}

/*
 * One context per device is parked between calls (the larger one when two compete): an MCMC run calls
 * samplea / sampleb once per sweep with the same shapes, and a dozen cudaMalloc / cudaFree plus a pinned
 * allocation per call cost about as much as sampleb's evaluations.  stb_cuda_pstat_purge() frees them
 * (stb_release_caches).  The parking slots are guarded by a mutex: the multi-device entry points run one
 * sampler per device concurrently (a context in use is never in a slot).
 */
#define STB_MAX_DEVICES 64
static stb_pstat_dev_t *g_pstat_parked[STB_MAX_DEVICES];
static std::mutex g_pstat_mutex;

static stb_pstat_dev_t *pstat_park_take(int dev) {
  if (dev < 0 || dev >= STB_MAX_DEVICES) return NULL;
  std::lock_guard<std::mutex> lk(g_pstat_mutex);
  stb_pstat_dev_t *q = g_pstat_parked[dev];
  g_pstat_parked[dev] = NULL;
  return q;
}

extern "C" void stb_cuda_pstat_purge(void) {
  for (int dev = 0; dev < STB_MAX_DEVICES; dev++) pstat_free(pstat_park_take(dev));
}

extern "C" void stb_cuda_pstat_destroy(stb_pstat_dev_t *p) {
  if (!p) return;
  if (p->stream) {
    stb::DeviceGuard dev_guard_(p->device);
    cudaStreamSynchronize(p->stream);
  }
  stb_pstat_dev_t *drop = p;
  if (p->device >= 0 && p->device < STB_MAX_DEVICES) {
    std::lock_guard<std::mutex> lk(g_pstat_mutex);
    stb_pstat_dev_t *&slot = g_pstat_parked[p->device];
    if (!slot || slot->cap < p->cap) {
      drop = slot;
      slot = p;
    }
  }
  pstat_free(drop);
}

static int pstat_upload_u32(uint32_t **dst, const uint32_t *src, int I) {
  if (!src) return 0;
  if (!*dst) PCK(cudaMalloc(dst, (size_t)I * sizeof(uint32_t)));
  PCK(cudaMemcpy(*dst, src, (size_t)I * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" stb_pstat_dev_t *stb_cuda_pstat_create(int I, const uint32_t *T, const uint32_t *N, const double *bpar,
                                                  size_t bpar_elems, size_t max_evals) {
  if (stb_cuda_device_count() <= 0) return NULL;
  {
    /* this device's parked context, when it fits: only the statistics are uploaded again */
    int dev = -1;
    stb_pstat_dev_t *q = cudaGetDevice(&dev) == cudaSuccess ? pstat_park_take(dev) : NULL;
    if (q && q->I == I && q->cap >= (max_evals ? max_evals : 1)) {
      int bad = pstat_upload_u32(&q->dT, T, I) || pstat_upload_u32(&q->dN, N, I);
      if (!bad && bpar && bpar_elems) {
        if (bpar_elems > q->bpar_cap) {
          cudaFree(q->dbpar);
          q->dbpar = NULL;
          q->bpar_cap = 0;
          bad = cudaMalloc(&q->dbpar, bpar_elems * sizeof(double)) != cudaSuccess;
          if (!bad) q->bpar_cap = bpar_elems;
        }
        if (!bad) bad = cudaMemcpy(q->dbpar, bpar, bpar_elems * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess;
      }
      if (!bad) {
        q->bpar_elems = bpar_elems;
        return q;
      }
      cudaGetLastError();
    }
    pstat_free(q); /* does not fit (or failed): build a fresh one below */
  }
  stb_pstat_dev_t *p = (stb_pstat_dev_t *)calloc(1, sizeof *p);
  if (!p) return NULL;
  p->I = I;
  p->cap = max_evals ? max_evals : 1;
  p->bpar_elems = bpar_elems;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
  if (e == cudaSuccess && T) {
    e = cudaMalloc(&p->dT, (size_t)I * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(p->dT, T, (size_t)I * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && N) {
    e = cudaMalloc(&p->dN, (size_t)I * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(p->dN, N, (size_t)I * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && bpar && bpar_elems) {
    e = cudaMalloc(&p->dbpar, bpar_elems * sizeof(double));
    if (e == cudaSuccess) p->bpar_cap = bpar_elems;
    if (e == cudaSuccess) e = cudaMemcpy(p->dbpar, bpar, bpar_elems * sizeof(double), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    e = cudaMalloc(&p->dzig, sizeof(stb_zig_tables));
    if (e == cudaSuccess) e = cudaMemcpy(p->dzig, stb_zig_tables_get(), sizeof(stb_zig_tables), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&p->dx, p->cap * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->da1, p->cap * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->da2, p->cap * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->dout, p->cap * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->dchain, p->cap * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&p->drng, p->cap * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaHostAlloc(&p->hbuf, 4 * p->cap * sizeof(double), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    stb_cuda_set_error("stb_cuda_pstat_create", (int)e);
    pstat_free(p);
    return NULL;
  }
  return p;
}

static int stage_in(stb_pstat_dev_t *p, double *dst, const double *src, size_t cnt) {
  if (!src) return 0;
  PCK(cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  return 0;
}

extern "C" int stb_cuda_pstat_aterms_lg(stb_pstat_dev_t *p, const double *x, const int *chain, size_t cnt,
                                        int bpar_per_chain, double *out, float *ms) {
  P_ON_DEVICE(p->device);
  if (cnt > p->cap || !p->dT || !p->dbpar) {
    stb_cuda_set_error("stb_cuda_pstat_aterms_lg: bad call", -1);
    return -1;
  }
  if (!cnt) return 0;
  if (stage_in(p, p->dx, x, cnt)) return -1;
  PCK(cudaMemcpyAsync(p->dchain, chain, cnt * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  PCK(cudaEventRecord(p->ev0, p->stream));
  k_aterms_lg<<<(unsigned)cnt, 256, 0, p->stream>>>(p->dT, p->dbpar, bpar_per_chain ? p->I : 0, p->I, p->dx, p->dchain,
                                                    p->dout);
  PCK(cudaGetLastError());
  PCK(cudaEventRecord(p->ev1, p->stream));
  PCK(cudaMemcpyAsync(out, p->dout, cnt * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  PCK(cudaStreamSynchronize(p->stream));
  if (ms) PCK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return 0;
}

extern "C" int stb_cuda_pstat_bterms(stb_pstat_dev_t *p, const double *x, const double *Q, const double *apar,
                                     double shape, size_t cnt, int digamma_sum, double *out, float *ms) {
  P_ON_DEVICE(p->device);
  if (cnt > p->cap || !p->dT) {
    stb_cuda_set_error("stb_cuda_pstat_bterms: bad call", -1);
    return -1;
  }
  if (!cnt) return 0;
  if (stage_in(p, p->dx, x, cnt) || stage_in(p, p->da1, Q, cnt) || stage_in(p, p->da2, apar, cnt)) return -1;
  PCK(cudaEventRecord(p->ev0, p->stream));
  if (digamma_sum)
    k_bdigamma<<<(unsigned)cnt, 256, 0, p->stream>>>(p->dT, p->I, p->dx, p->da2, p->dout);
  else
    k_bterms<<<(unsigned)cnt, 256, 0, p->stream>>>(p->dT, p->I, p->dx, p->da1, p->da2, shape, p->dout);
  PCK(cudaGetLastError());
  PCK(cudaEventRecord(p->ev1, p->stream));
  PCK(cudaMemcpyAsync(out, p->dout, cnt * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  PCK(cudaStreamSynchronize(p->stream));
  if (ms) PCK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return 0;
}

extern "C" int stb_cuda_pstat_betaQ(stb_pstat_dev_t *p, const double *b_in, uint64_t *rng, size_t C, double scale,
                                    double *Q, float *ms) {
  P_ON_DEVICE(p->device);
  if (C > p->cap || !p->dN) {
    stb_cuda_set_error("stb_cuda_pstat_betaQ: bad call", -1);
    return -1;
  }
  if (!C) return 0;
  if (stage_in(p, p->dx, b_in, C)) return -1;
  PCK(cudaMemcpyAsync(p->drng, rng, C * sizeof(uint64_t), cudaMemcpyHostToDevice, p->stream));
  PCK(cudaEventRecord(p->ev0, p->stream));
  k_betaQ<<<(unsigned)((C + 1) / 2), 64, 0, p->stream>>>(p->dN, p->I, p->dx, p->drng, scale, p->dzig, p->dout, (int)C);
  PCK(cudaGetLastError());
  PCK(cudaEventRecord(p->ev1, p->stream));
  PCK(cudaMemcpyAsync(Q, p->dout, C * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  PCK(cudaMemcpyAsync(rng, p->drng, C * sizeof(uint64_t), cudaMemcpyDeviceToHost, p->stream));
  PCK(cudaStreamSynchronize(p->stream));
  if (ms) PCK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return 0;
}
