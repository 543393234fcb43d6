/*
 * stb_cuda.cu -- device-side table object and kernel launchers behind stb_cuda.h.
 *
 * Replaces the allocation part of S_make (lib/stable.c:155-304: ragged row pointers, one
 * malloc per row) by ONE dense row-major slab per table in HBM, and S_remake_part's two
 * double loops (lib/stable.c:356-388, 451-482) by the kernels in fill_strip.cuh /
 * fill_mirror.cuh.  sm_100a only; no CPU path.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "stb_cuda.h"
#include "dev_guard.cuh"
#include "fill_mirror.cuh"
#include "fill_strip.cuh"

/* last error text, per thread: S_THREADS callers and the per-device worker threads of the multi-device
 * entry points (multi.c) fail independently */
static thread_local char g_err[512] = "";

static int fail(cudaError_t e, const char *what) {
  snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return (int)e ? (int)e : -1;
}
#define ON_DEVICE(dev)              \
  stb::DeviceGuard dev_guard_(dev); \
  if (dev_guard_.err != cudaSuccess) return fail(dev_guard_.err, "cudaSetDevice")
#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) return fail(e_, #call); \
  } while (0)

struct stb_dev {
  int device;
  int num_sms;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  int has_S, has_V, is_float;
  unsigned capN, capM;  // rows / columns the slabs can hold
  size_t ld;            // elements per row
  void *S, *V;          // [capN][ld] of double or float
  double *s1;           // device copy of column 1 / S1 (capN doubles)
  double *scratch;      // mirror kernel live rows
  size_t scratch_elems;
  stb::StripState strip;  // hand-off rings, counters and tables array of the strip kernel
  float last_ms;
  // staging for host-pointer gathers
  uint32_t *g_n, *g_m;
  double *g_out;
  size_t g_cap;
  // staging of the partition sampler, and the device time of its most recent kernel
  char *p_buf;
  size_t p_cap;
  float last_part_ms;
};

extern "C" int stb_cuda_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(e, "cudaGetDeviceCount");
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char *stb_cuda_last_error(void) { return g_err; }

extern "C" int stb_cuda_current_device(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return dev;
}

extern "C" int stb_cuda_use_device(int dev) {
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return fail(e, "cudaSetDevice");
  return 0;
}
extern "C" void stb_cuda_set_error(const char *what, int code) {
  if (code > 0)
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString((cudaError_t)code));
  else
    snprintf(g_err, sizeof g_err, "%s", what);
}

extern "C" stb_dev_t *stb_cuda_table_create(int want_S, int want_V, int is_float) {
  if (stb_cuda_device_count() <= 0) {
    if (!g_err[0]) snprintf(g_err, sizeof g_err, "no CUDA device");
    return NULL;
  }
  stb_dev_t *d = (stb_dev_t *)calloc(1, sizeof *d);
  if (!d) return NULL;
  d->has_S = want_S != 0;
  d->has_V = want_V != 0;
  d->is_float = is_float != 0;
  if (cudaGetDevice(&d->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, d->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&d->ev0) != cudaSuccess || cudaEventCreate(&d->ev1) != cudaSuccess) {
    fail(cudaGetLastError(), "stb_cuda_table_create");
    free(d);
    return NULL;
  }
  return d;
}

extern "C" void stb_cuda_table_destroy(stb_dev_t *d) {
  if (!d) return;
  stb::DeviceGuard dev_guard_(d->device);
  cudaStreamSynchronize(d->stream);
  cudaFree(d->S);
  cudaFree(d->V);
  cudaFree(d->s1);
  cudaFree(d->scratch);
  cudaFree(d->g_n);
  cudaFree(d->g_m);
  cudaFree(d->g_out);
  cudaFree(d->p_buf);
  stb::strip_state_free(&d->strip);
  cudaEventDestroy(d->ev0);
  cudaEventDestroy(d->ev1);
  cudaStreamDestroy(d->stream);
  free(d);
}

extern "C" int stb_cuda_table_device(const stb_dev_t *d) { return d->device; }
extern "C" int stb_cuda_table_is_float(const stb_dev_t *d) { return d->is_float; }

static size_t elem_size(const stb_dev_t *d) { return d->is_float ? sizeof(float) : sizeof(double); }

extern "C" size_t stb_cuda_table_ld(const stb_dev_t *d) { return d->ld; }

extern "C" size_t stb_cuda_table_bytes(const stb_dev_t *d) {
  size_t slab = (size_t)d->capN * d->ld * elem_size(d);
  return slab * (size_t)(d->has_S + d->has_V) + (size_t)d->capN * sizeof(double) +
         d->scratch_elems * sizeof(double) + stb::strip_state_bytes(&d->strip);
}

static void table_release(stb_dev_t *d) {
  cudaFree(d->S);
  cudaFree(d->V);
  cudaFree(d->s1);
  d->S = d->V = NULL;
  d->s1 = NULL;
  d->capN = d->capM = 0;
  d->ld = 0;
}

/* fresh slabs for newN x newM (the old ones are gone by now); on failure the table holds nothing */
static cudaError_t table_alloc(stb_dev_t *d, unsigned newN, unsigned newM) {
  const size_t newld = ((size_t)newM + 31) / 32 * 32, es = elem_size(d);
  cudaError_t e = cudaSuccess;
  if (d->has_S) e = cudaMalloc(&d->S, (size_t)newN * newld * es);
  if (e == cudaSuccess && d->has_V) e = cudaMalloc(&d->V, (size_t)newN * newld * es);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d->s1, (size_t)newN * sizeof(double));
  if (e != cudaSuccess) {
    table_release(d);
    cudaGetLastError();
    return e;
  }
  d->capN = newN;
  d->capM = newM;
  d->ld = newld;
  return cudaSuccess;
}

extern "C" int stb_cuda_table_reserve(stb_dev_t *d, unsigned N, unsigned M, int keep) {
  ON_DEVICE(d->device);
  if (N <= d->capN && M <= d->capM) return 0;
  const unsigned newN = N > d->capN ? N : d->capN, newM = M > d->capM ? M : d->capM;
  if (!keep || !d->capN) {  // nothing worth keeping: give the old slabs back first (never two generations at once)
    CK(cudaStreamSynchronize(d->stream));
    table_release(d);
    cudaError_t e = table_alloc(d, newN, newM);
    return e == cudaSuccess ? 0 : fail(e, "cudaMalloc(table slab)");
  }
  // keep the filled cells: new slabs beside the old ones, a pitched copy, then the old ones go
  stb_dev_t old = *d;
  d->S = d->V = NULL;
  d->s1 = NULL;
  cudaError_t e = table_alloc(d, newN, newM);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  const size_t es = elem_size(d);
  if (e == cudaSuccess && d->has_S)
    e = cudaMemcpy2D(d->S, d->ld * es, old.S, old.ld * es, (size_t)old.capM * es, old.capN, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess && d->has_V)
    e = cudaMemcpy2D(d->V, d->ld * es, old.V, old.ld * es, (size_t)old.capM * es, old.capN, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d->s1, old.s1, (size_t)old.capN * sizeof(double), cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) {  // back to the old generation, nothing leaks
    cudaFree(d->S);
    cudaFree(d->V);
    cudaFree(d->s1);
    d->S = old.S;
    d->V = old.V;
    d->s1 = old.s1;
    d->capN = old.capN;
    d->capM = old.capM;
    d->ld = old.ld;
    return fail(e, "stb_cuda_table_reserve(keep)");
  }
  cudaFree(old.S);
  cudaFree(old.V);
  cudaFree(old.s1);
  return 0;
}

/*
 * Room for a table that GROWS to N x M and will be refilled as a whole (S_extend): nothing is copied, the
 * old slabs are freed before the new ones are allocated, and the capacity at least doubles in the
 * dimension that ran out (capped by the table's maxima) -- the reference's growth policy is +10 % per
 * step (lib/stable.c:590-601): ~41 steps from 10 000 to 500 000 rows become 6 allocations.
 */
extern "C" int stb_cuda_table_grow(stb_dev_t *d, unsigned N, unsigned M, unsigned maxN, unsigned maxM) {
  ON_DEVICE(d->device);
  if (N <= d->capN && M <= d->capM) return 0;
  unsigned newN = d->capN, newM = d->capM;
  if (N > d->capN) {
    const unsigned long long twice = 2ull * d->capN;
    newN = (unsigned)(twice > N ? (twice < maxN ? twice : maxN) : N);
    if (newN < N) newN = N;
  }
  if (M > d->capM) {
    const unsigned long long twice = 2ull * d->capM;
    newM = (unsigned)(twice > M ? (twice < maxM ? twice : maxM) : M);
    if (newM < M) newM = M;
  }
  if (newM > newN) newM = newN > M ? newN : M;
  CK(cudaStreamSynchronize(d->stream));
  table_release(d);
  cudaError_t e = table_alloc(d, newN, newM);
  if (e != cudaSuccess && (newN > N || newM > M)) e = table_alloc(d, N, M);  // no room for the head start: the exact extent
  return e == cudaSuccess ? 0 : fail(e, "cudaMalloc(table slab)");
}

extern "C" void *stb_cuda_table_ptr(stb_dev_t *d, int which) { return which == STB_TAB_S ? d->S : d->V; }

extern "C" float stb_cuda_last_fill_ms(const stb_dev_t *d) { return d->last_ms; }
extern "C" float stb_cuda_last_partition_ms(const stb_dev_t *d) { return d->last_part_ms; }

static int fill_mirror(stb_dev_t *d, double a, unsigned N, unsigned M, double *s1_host) {
  size_t need = 4 * ((size_t)M + 2);
  if (need > d->scratch_elems) {
    cudaFree(d->scratch);
    d->scratch = NULL;
    d->scratch_elems = 0;
    CK(cudaMalloc(&d->scratch, need * sizeof(double)));
    d->scratch_elems = need;
  }
  if (!s1_host) {
    snprintf(g_err, sizeof g_err, "mirror fill needs the host S1 running sum");
    return -1;
  }
  CK(cudaMemcpyAsync(d->s1, s1_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, d->stream));
  CK(cudaEventRecord(d->ev0, d->stream));
  if (d->is_float)
    stb::fill_mirror_kernel<float><<<1, 1024, 0, d->stream>>>((float *)d->S, (float *)d->V, d->s1,
                                                             d->scratch, d->ld, N, M, a);
  else
    stb::fill_mirror_kernel<double><<<1, 1024, 0, d->stream>>>((double *)d->S, (double *)d->V, d->s1,
                                                               d->scratch, d->ld, N, M, a);
  CK(cudaGetLastError());
  CK(cudaEventRecord(d->ev1, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  CK(cudaEventElapsedTime(&d->last_ms, d->ev0, d->ev1));
  return 0;
}

extern "C" int stb_cuda_fill(stb_dev_t *d, double a, unsigned startN, unsigned startM, unsigned N,
                             unsigned M, int algo, double *s1_host) {
  ON_DEVICE(d->device);
  if (N > d->capN || M > d->capM || N < 1 || M < 1) {
    snprintf(g_err, sizeof g_err, "stb_cuda_fill: extent %ux%u exceeds reserved %ux%u", N, M, d->capN,
             d->capM);
    return -1;
  }
  if (algo == STB_FILL_MIRROR) return fill_mirror(d, a, N, M, s1_host);
  CK(cudaEventRecord(d->ev0, d->stream));
  int rc;
  {
    // linear-domain strip pipeline; the whole extent is refilled (startN/startM are an
    // optimisation the reference has and this path does not need: a refill costs milliseconds)
    stb::StripTable tb;
    tb.tabS = d->has_S ? d->S : NULL;
    tb.tabV = d->has_V ? d->V : NULL;
    tb.s1 = d->s1;
    tb.a = a;
    stb::StripFillArgs args;
    args.tables = &tb;
    args.ntables = 1;
    args.has_S = d->has_S;
    args.has_V = d->has_V;
    args.is_float = d->is_float;
    args.ld = d->ld;
    args.N = N;
    args.M = M;
    args.num_sms = d->num_sms;
    args.async_flag = NULL;
    rc = stb::strip_fill(&d->strip, args, d->stream, d->ev1, g_err, sizeof g_err);
  }
  if (rc) return rc;
  if (s1_host && d->has_S)
    CK(cudaMemcpyAsync(s1_host, d->s1, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  CK(cudaEventElapsedTime(&d->last_ms, d->ev0, d->ev1));
  return 0;
}

/* log S^n_1, n = 1..N, of the most recent fill (the strip kernel leaves the column in d->s1) */
extern "C" int stb_cuda_read_s1(stb_dev_t *d, unsigned N, double *dst) {
  ON_DEVICE(d->device);
  if (!d->s1 || N > d->capN) {
    snprintf(g_err, sizeof g_err, "stb_cuda_read_s1: bad range");
    return -1;
  }
  CK(cudaMemcpyAsync(dst, d->s1, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}

extern "C" int stb_cuda_read_rows(stb_dev_t *d, int which, unsigned row0, unsigned nrows, double *dst) {
  ON_DEVICE(d->device);
  const void *tab = which == STB_TAB_S ? d->S : d->V;
  if (!tab || row0 + nrows > d->capN) {
    snprintf(g_err, sizeof g_err, "stb_cuda_read_rows: bad range");
    return -1;
  }
  size_t cnt = (size_t)nrows * d->ld;
  if (!d->is_float) {
    CK(cudaMemcpyAsync(dst, (const double *)tab + (size_t)row0 * d->ld, cnt * sizeof(double),
                       cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
  } else {
    // copy the floats into the tail of dst, then widen front to back
    float *tmp = (float *)dst + cnt;
    CK(cudaMemcpyAsync(tmp, (const float *)tab + (size_t)row0 * d->ld, cnt * sizeof(float),
                       cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    for (size_t i = 0; i < cnt; i++) dst[i] = (double)tmp[i];
  }
  return 0;
}

/*
 * Batched look-up with the scalar API's in-range conventions (lib/stable.c:941-949, 927-928):
 *   S: n==m -> 0, m==0 or n<m -> -inf, else the cell (m==1 is column 0 of the slab);
 *   V: m<2 or n<m -> 0, else the cell.
 * Anything outside the filled extent answers like "beyond bounds" (-inf / 0).
 */
/* what: STB_TAB_S, STB_TAB_V, or the ratios derived from V: STB_GATHER_U = n - m a + 1/V (m == 1: n - a,
 * lib/stable.c:875-883), STB_GATHER_UV = (n - m a) V + 1 (m == 1: -inf, m == n+1: 1, m == n: (n+1)/(n-1),
 * lib/stable.c:885-897); m == 0 answers NaN for U (the scalar call exits) */
template <typename T>
__global__ void gather_kernel(const T *__restrict__ tab, const double *__restrict__ s1, size_t ld, int what, double a,
                              unsigned usedN, unsigned usedM,
                              const uint32_t *__restrict__ n, const uint32_t *__restrict__ m,
                              double *__restrict__ out, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const unsigned nn = n[i], mm = m[i];
  double v;
  if (what == STB_TAB_S) {
    if (nn == mm)
      v = 0.0;
    else if (mm == 0 || nn < mm || nn > usedN || mm > usedM)
      v = -HUGE_VAL;
    else if (mm == 1)
      v = s1[nn - 1];  // S_S(n,1) is S_S1(n): FP64 also when the table stores floats (lib/stable.c:946-947)
    else
      v = (double)tab[(size_t)(nn - 1) * ld + (mm - 1)];
  } else {
    const double V = (mm < 2 || nn < mm || nn > usedN || mm > usedM) ? 0.0 : (double)tab[(size_t)(nn - 1) * ld + (mm - 1)];
    if (what == STB_TAB_V)
      v = V;
    else {
      // the scalar calls' arithmetic, operation by operation (no FMA contraction): bit-identical to S_U / S_UV
      const double nd = (double)nn, nma = __dsub_rn(nd, __dmul_rn((double)mm, a));
      if (what == STB_GATHER_U)
        v = mm == 1 ? __dsub_rn(nd, a) : (mm == 0 ? nan("") : __dadd_rn(nma, __ddiv_rn(1.0, V)));
      else
        v = mm == 1 ? -HUGE_VAL
                    : (mm == nn + 1 ? 1.0
                                    : (mm == nn ? __ddiv_rn(__dadd_rn(nd, 1.0), __dsub_rn(nd, 1.0))
                                                : __dadd_rn(__dmul_rn(nma, V), 1.0)));
    }
  }
  out[i] = v;
}

extern "C" int stb_cuda_gather(stb_dev_t *d, int which, double a, unsigned usedN, unsigned usedM, const uint32_t *n,
                               const uint32_t *m, double *out, size_t count, int on_device) {
  ON_DEVICE(d->device);
  const void *tab = which == STB_TAB_S ? d->S : d->V;
  if (!tab) {
    snprintf(g_err, sizeof g_err, "stb_cuda_gather: table not held");
    return -1;
  }
  if (count == 0) return 0;
  const uint32_t *dn = n, *dm = m;
  double *dout = out;
  if (!on_device) {
    if (count > d->g_cap) {
      cudaFree(d->g_n);
      cudaFree(d->g_m);
      cudaFree(d->g_out);
      d->g_n = d->g_m = NULL;
      d->g_out = NULL;
      d->g_cap = 0;
      CK(cudaMalloc(&d->g_n, count * sizeof(uint32_t)));
      CK(cudaMalloc(&d->g_m, count * sizeof(uint32_t)));
      CK(cudaMalloc(&d->g_out, count * sizeof(double)));
      d->g_cap = count;
    }
    CK(cudaMemcpyAsync(d->g_n, n, count * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
    CK(cudaMemcpyAsync(d->g_m, m, count * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
    dn = d->g_n;
    dm = d->g_m;
    dout = d->g_out;
  }
  unsigned blocks = (unsigned)((count + 255) / 256);
  if (d->is_float)
    gather_kernel<float><<<blocks, 256, 0, d->stream>>>((const float *)tab, d->s1, d->ld, which, a, usedN, usedM, dn, dm,
                                                        dout, count);
  else
    gather_kernel<double><<<blocks, 256, 0, d->stream>>>((const double *)tab, d->s1, d->ld, which, a, usedN, usedM, dn,
                                                         dm, dout, count);
  CK(cudaGetLastError());
  if (!on_device)
    CK(cudaMemcpyAsync(out, d->g_out, count * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// seat-partition sampler of samplea2 (lib/samplea.c:290-321)
// ---------------------------------------------------------------------------------------------
/* log(e^x - e^y), lib/samplea.c:233-239 */
__device__ __forceinline__ double logminus_dev(double x, double y) {
  if (y >= x) return -HUGE_VAL;
  const double d = __dsub_rn(y, x);
  if (d < -80.0) return __dsub_rn(x, exp(d));
  return __dadd_rn(x, log(__dsub_rn(1.0, exp(d))));
}

/*
 * One thread per node (n customers at t tables, 1 < t < n).  Round M = t-1 .. 1 draws the size l of
 * one more table from the remaining N customers by walking l = 1, 2, ... and subtracting each
 * term's mass from the running remainder.
 *   exact == 0: the reference's loop, operation by operation (lib/samplea.c:293-314): ONE uniform per
 *     node, rem = ptot + log u with ptot = S(n,t) never renewed, growth factor (l - a)(N-l+1)/(l-1).
 *     Because ptot (a large log Stirling number) is added to the remainder while the terms are
 *     normalised by it, no term reaches the remainder unless S(n,t) is tiny (n of a few units):
 *     the walk runs to its end and the "sample" is one table of N-M customers and singletons.
 *     Mirrored for parity.
 *   exact != 0: the distribution the loop is after, P(l | N, M+1) = C(N-1,l-1) (1-a)_{l-1} S^{N-l}_M /
 *     S^N_{M+1} (it sums to one by the recurrence of the generalised Stirling numbers): a fresh uniform
 *     per round (logu[off + M-1]), rem = log u, the normaliser S(N, M+1) of the CURRENT N, growth
 *     factor (l-1-a)(N-l+1)/(l-1).
 * The walk is sequential, but the rounds of one node visit about n cells in total, so a node costs
 * O(n) table reads; the table (a few MB for the samplers' sizes) stays in L2.
 */
template <typename T>
__global__ void partition_kernel(const T *__restrict__ tab, const double *__restrict__ s1, size_t ld, double a,
                                 const uint32_t *__restrict__ n, const uint16_t *__restrict__ t,
                                 const double *__restrict__ logu, const uint32_t *__restrict__ off,
                                 const uint32_t *__restrict__ order, uint16_t *__restrict__ m, size_t count,
                                 int exact) {
  const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= count) return;
  // nodes are taken longest first: the 32 walks of a warp have about the same length, and the
  // longest walks of all (the critical path of the launch) start at once
  const size_t i = order[slot];
  int N = (int)n[i];
  const int t0 = (int)t[i];
  auto S = [&](int nn, int mm) -> double {  // S_S, lib/stable.c:941-974, inside the table
    if (nn == mm) return 0.0;
    if (mm == 1) return s1[nn - 1];
    return (double)tab[(size_t)(nn - 1) * ld + (mm - 1)];
  };
  double ptot = S(N, t0);
  double rem = exact ? logu[off[i] + t0 - 2] : __dadd_rn(ptot, logu[i]);
  const double shift = exact ? 1.0 : 0.0;
  uint16_t *mp = m + off[i];
  // The reference's two nested loops (rounds M, walk l) as ONE loop that advances the walk by a step
  // per trip and changes round in a short predicated tail: lanes whose walks end at different l do
  // not wait for each other at every round boundary, only at the end of the node.
  int M = t0 - 1, l = 1;
  double fact = 0.0;
  // the table cell of the NEXT step is fetched a step ahead: its address does not depend on the
  // remainder, its L2 latency (the longest link of a step) then overlaps the exp/log chain
  double s_next = S(N - 1, M);
  while (M >= 1) {
    const double s_cur = s_next;
    s_next = S(N - (l + 1 <= N - M ? l + 1 : l), M);
    if (l > 1)
      fact = __dadd_rn(fact, log(__ddiv_rn(__dmul_rn(__dsub_rn((double)l - shift, a), (double)(N - l + 1)),
                                           (double)(l - 1))));
    const double term = __dsub_rn(__dadd_rn(fact, s_cur), ptot);
    bool done = term >= rem;
    if (!done) {
      rem = logminus_dev(rem, term);
      l++;
      if (l > N - M) {  // walked off the end: the last admissible size
        l = N - M;
        done = true;
      }
    }
    if (done) {
      mp[M - 1] = (uint16_t)l;
      N -= l;
      M--;
      l = 1;
      fact = 0.0;
      if (M >= 1) {
        s_next = S(N - 1, M);
        if (exact) {
          ptot = S(N, M + 1);
          rem = logu[off[i] + M - 1];
        }
      }
    }
  }
}

extern "C" int stb_cuda_partition(stb_dev_t *d, double a, const uint32_t *n, const uint16_t *t, const double *logu,
                                  const uint32_t *off, size_t count, uint16_t *m_out, size_t n_m, int exact) {
  ON_DEVICE(d->device);
  if (!d->S || !d->s1) {
    snprintf(g_err, sizeof g_err, "stb_cuda_partition: S table not held");
    return -1;
  }
  if (count == 0 || n_m == 0) return 0;
  // the nodes by decreasing n (counting sort; n <= capN was checked by the caller)
  std::vector<uint32_t> order, start;
  try {
    order.resize(count);
    start.assign((size_t)d->capN + 2, 0);
  } catch (...) {  // no exception crosses the C ABI
    snprintf(g_err, sizeof g_err, "stb_cuda_partition: out of host memory");
    return -1;
  }
  for (size_t i = 0; i < count; i++) {
    if (n[i] > d->capN) {
      snprintf(g_err, sizeof g_err, "stb_cuda_partition: n beyond the table");
      return -1;
    }
    start[d->capN - n[i] + 1]++;
  }
  for (size_t v = 1; v < start.size(); v++) start[v] += start[v - 1];
  for (size_t i = 0; i < count; i++) order[start[d->capN - n[i]]++] = (uint32_t)i;
  // one staging block: logu (f64), n, off, order (u32), t (u16), m (u16)
  const size_t b_n = count * sizeof(uint32_t), b_u = (exact ? n_m : count) * sizeof(double), b_t = count * sizeof(uint16_t),
               b_m = n_m * sizeof(uint16_t);
  const size_t o_u = 0, o_n = o_u + b_u, o_off = o_n + b_n, o_ord = o_off + b_n, o_t = o_ord + b_n,
               o_m = (o_t + b_t + 7) & ~(size_t)7;
  if (o_m + b_m > d->p_cap) {
    cudaFree(d->p_buf);
    d->p_buf = NULL;
    d->p_cap = 0;
    CK(cudaMalloc(&d->p_buf, o_m + b_m));
    d->p_cap = o_m + b_m;
  }
  char *buf = d->p_buf;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + o_ord, order.data(), b_n, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + o_u, logu, b_u, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + o_n, n, b_n, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + o_off, off, b_n, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + o_t, t, b_t, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaEventRecord(d->ev0, d->stream);
  if (e == cudaSuccess) {
    // 64 threads per block: nodes differ in length by orders of magnitude, small blocks spread the long ones
    const unsigned blocks = (unsigned)((count + 63) / 64);
    if (d->is_float)
      partition_kernel<float><<<blocks, 64, 0, d->stream>>>((const float *)d->S, d->s1, d->ld, a,
                                                            (const uint32_t *)(buf + o_n), (const uint16_t *)(buf + o_t),
                                                            (const double *)(buf + o_u), (const uint32_t *)(buf + o_off),
                                                            (const uint32_t *)(buf + o_ord), (uint16_t *)(buf + o_m), count,
                                                            exact);
    else
      partition_kernel<double><<<blocks, 64, 0, d->stream>>>((const double *)d->S, d->s1, d->ld, a,
                                                             (const uint32_t *)(buf + o_n), (const uint16_t *)(buf + o_t),
                                                             (const double *)(buf + o_u), (const uint32_t *)(buf + o_off),
                                                             (const uint32_t *)(buf + o_ord), (uint16_t *)(buf + o_m), count,
                                                             exact);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventRecord(d->ev1, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m_out, buf + o_m, b_m, cudaMemcpyDeviceToHost, d->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  if (e == cudaSuccess) e = cudaEventElapsedTime(&d->last_part_ms, d->ev0, d->ev1);
  if (e != cudaSuccess) return fail(e, "stb_cuda_partition");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// discount sweep
// ---------------------------------------------------------------------------------------------
struct stb_sweep_dev {
  int device, num_sms;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  unsigned N, M;
  int is_float;
  size_t ld;
  int T;               // tables a launch fills side by side (CTA groups)
  int slabs;           // resident slabs = tables per launch: T groups x rounds (the kernel's rounds, fill_strip.cuh)
  void *slab;          // [slabs][N][ld]
  double *s1;          // [slabs][N]
  stb::StripState strip;
  uint32_t *d_n, *d_m;
  size_t npairs, pairs_cap, gather_cap;
  double *d_gather;    // [T][npairs], allocated when a gather is first asked for
  double *d_partial;   // [T][nblk]
  // the DISTINCT pairs in cell order with their multiplicities (built on the device by set_pairs when
  // every pair names a stored cell): what the sums-only runs gather -- fewer, address-ordered reads
  unsigned *d_cnt;     // [N][ld] multiplicity per cell (scratch of the build)
  unsigned *d_blkoff;  // per 256-cell block: distinct pairs before it; [nblocks] = total
  uint32_t *d_un, *d_um;
  double *d_uw;
  size_t nuniq, uniq_cap;
  double *d_sum;       // [stage_cap]: one sum per table of a run
  double *h_stage;     // pinned staging for the sums of a whole run
  int *h_flags;        // pinned: watchdog flag of every wave of a run
  size_t stage_cap, flags_cap;
  // the batched seat-partition step of samplea2 (stb_cuda_sweep_set_nodes / _partition / _hist_eval)
  uint32_t *d_pn, *d_pdraw, *d_porder;  // [pcount]: customers, first draw index, nodes longest first
  uint16_t *d_pt;                       // [pcount]: tables
  size_t pcount, pnodes_cap;
  unsigned hbins;                       // histogram bins per chain: arguments j = size - 1 = 0 .. hbins-1
  uint32_t *d_hist, *d_hbase;           // [hist_cap][hbins] per-chain histograms of the sampled sizes; [hbins] what every chain starts from
  size_t hist_cap;
  unsigned long long *d_x0;             // [slabs] stream states of a wave's chains
  double *d_ex, *d_eout;                // evaluation points / results
  int *d_echain;
  size_t eval_cap;
};

/* S_S conventions (lib/stable.c:941-949) on a dense slab; partial sums per block in a fixed order */
template <typename T>
__global__ void sweep_gather_kernel(const T *__restrict__ slabs, const double *__restrict__ s1s, size_t slab_elems,
                                    size_t ld, unsigned N, unsigned M,
                                    const uint32_t *__restrict__ n, const uint32_t *__restrict__ m,
                                    const double *__restrict__ wt, size_t npairs, double *__restrict__ gather,
                                    double *__restrict__ partial) {
  __shared__ double red[256];
  const int tb = blockIdx.y;
  const T *tab = slabs + (size_t)tb * slab_elems;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < npairs) {
    const unsigned nn = n[i], mm = m[i];
    if (nn == mm)
      v = 0.0;
    else if (mm == 0 || nn < mm || nn > N || mm > M)
      v = -HUGE_VAL;
    else if (mm == 1)
      v = s1s[(size_t)tb * N + (nn - 1)];  // S_S(n,1) = S_S1(n), FP64 whatever the table stores
    else
      v = (double)tab[(size_t)(nn - 1) * ld + (mm - 1)];
    if (gather) gather[(size_t)tb * npairs + i] = v;
    if (wt) v *= wt[i];  // a distinct pair stands for wt of the caller's
  }
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)tb * gridDim.x + blockIdx.x] = red[0];
}

/* ---- distinct pairs with multiplicities, in cell order (deterministic: no atomics decide an order) ---- */
__global__ void pair_count_kernel(const uint32_t *__restrict__ n, const uint32_t *__restrict__ m, size_t npairs, size_t ld,
                                  unsigned *__restrict__ cnt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const unsigned nn = n[i], mm = m[i];
  if (nn != mm) atomicAdd(&cnt[(size_t)(nn - 1) * ld + (mm - 1)], 1u);  // S^n_n = 1 adds nothing to a sum
}
__global__ void cell_count_kernel(const unsigned *__restrict__ cnt, size_t cells, unsigned *__restrict__ blkcnt) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const int c = __syncthreads_count(i < cells && cnt[i] != 0);
  if (threadIdx.x == 0) blkcnt[blockIdx.x] = (unsigned)c;
}
/* exclusive scan of nb block counts in place, total at [nb]; one block of 1024 threads */
__global__ void blk_scan_kernel(unsigned *__restrict__ blk, unsigned nb) {
  __shared__ unsigned part[1024];
  const unsigned per = (nb + 1023u) / 1024u, b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
  unsigned s = 0;
  for (unsigned b = b0; b < b1; b++) s += blk[b];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int t = 0; t < 1024; t++) {
      const unsigned v = part[t];
      part[t] = run;
      run += v;
    }
    blk[nb] = run;
  }
  __syncthreads();
  unsigned run = part[threadIdx.x];
  for (unsigned b = b0; b < b1; b++) {
    const unsigned v = blk[b];
    blk[b] = run;
    run += v;
  }
}
__global__ void cell_compact_kernel(const unsigned *__restrict__ cnt, size_t cells, size_t ld,
                                    const unsigned *__restrict__ blkoff, uint32_t *__restrict__ un,
                                    uint32_t *__restrict__ um, double *__restrict__ uw) {
  __shared__ unsigned wbase[8];
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const unsigned c = i < cells ? cnt[i] : 0u;
  const unsigned ball = __ballot_sync(0xffffffffu, c != 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) wbase[warp] = __popc(ball);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int w = 0; w < 8; w++) {
      const unsigned v = wbase[w];
      wbase[w] = run;
      run += v;
    }
  }
  __syncthreads();
  if (c != 0) {
    const unsigned pos = blkoff[blockIdx.x] + wbase[warp] + __popc(ball & ((1u << lane) - 1u));
    un[pos] = (uint32_t)(i / ld) + 1u;
    um[pos] = (uint32_t)(i % ld) + 1u;
    uw[pos] = (double)c;
  }
}

__global__ void sweep_sum_kernel(const double *__restrict__ partial, int nblk, double *__restrict__ sum) {
  __shared__ double red[256];
  const int tb = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) v += partial[(size_t)tb * nblk + b];
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) sum[tb] = red[0];
}

extern "C" void stb_cuda_sweep_destroy(stb_sweep_dev_t *w) {
  if (!w) return;
  stb::DeviceGuard dev_guard_(w->device);
  if (w->stream) cudaStreamSynchronize(w->stream);
  cudaFree(w->slab);
  cudaFree(w->s1);
  cudaFree(w->d_n);
  cudaFree(w->d_m);
  cudaFree(w->d_cnt);
  cudaFree(w->d_blkoff);
  cudaFree(w->d_un);
  cudaFree(w->d_um);
  cudaFree(w->d_uw);
  cudaFree(w->d_gather);
  cudaFree(w->d_partial);
  cudaFree(w->d_sum);
  if (w->h_stage) cudaFreeHost(w->h_stage);
  if (w->h_flags) cudaFreeHost(w->h_flags);
  cudaFree(w->d_pn);
  cudaFree(w->d_pdraw);
  cudaFree(w->d_porder);
  cudaFree(w->d_pt);
  cudaFree(w->d_hist);
  cudaFree(w->d_hbase);
  cudaFree(w->d_x0);
  cudaFree(w->d_ex);
  cudaFree(w->d_eout);
  cudaFree(w->d_echain);
  stb::strip_state_free(&w->strip);
  if (w->ev0) cudaEventDestroy(w->ev0);
  if (w->ev1) cudaEventDestroy(w->ev1);
  if (w->stream) cudaStreamDestroy(w->stream);
  free(w);
}

extern "C" stb_sweep_dev_t *stb_cuda_sweep_create(unsigned N, unsigned M, int is_float) {
  if (stb_cuda_device_count() <= 0) {
    if (!g_err[0]) snprintf(g_err, sizeof g_err, "no CUDA device");
    return NULL;
  }
  if (M < 1 || M > N) {
    snprintf(g_err, sizeof g_err, "stb_cuda_sweep_create: bad extent %ux%u", N, M);
    return NULL;
  }
  stb_sweep_dev_t *w = (stb_sweep_dev_t *)calloc(1, sizeof *w);
  if (!w) return NULL;
  w->N = N;
  w->M = M;
  w->is_float = is_float != 0;
  w->ld = ((size_t)M + 31) / 32 * 32;
  cudaError_t e = cudaGetDevice(&w->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&w->num_sms, cudaDevAttrMultiProcessorCount, w->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&w->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&w->ev1);
  if (e == cudaSuccess) {
    w->T = stb::strip_tables_per_launch(M, w->num_sms);
    // leave room on the device: never more than a quarter of its memory in slabs.  Within that, as many ROUNDS
    // of T tables as fit (at most 16 rounds, at most 64 GB or a third of what is free): a launch then fills rounds x T tables back to back, and the
    // launch gap, the ring resets and the pipeline's fill and drain are paid once per launch, not once per T tables
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t per = (size_t)N * w->ld * (is_float ? 4 : 8);
    while (w->T > 1 && (size_t)w->T * per > free_b / 4) w->T--;
    size_t budget = free_b / 3 < ((size_t)64 << 30) ? free_b / 3 : ((size_t)64 << 30);
    int rounds = (int)(budget / ((size_t)w->T * per));
    if (rounds > 16) rounds = 16;
    if (const char *sv = getenv("STB_SWEEP_ROUNDS")) rounds = atoi(sv);  // development: 1 = one wave per launch (round-1 behaviour)
    if (rounds < 1) rounds = 1;
    w->slabs = w->T * rounds;
    e = cudaMalloc(&w->slab, (size_t)w->slabs * per);
  }
  if (e == cudaSuccess) e = cudaMalloc(&w->s1, (size_t)w->slabs * N * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->d_sum, (size_t)w->slabs * sizeof(double));
  if (e == cudaSuccess) e = cudaHostAlloc(&w->h_stage, (size_t)w->slabs * sizeof(double), cudaHostAllocDefault);
  if (e == cudaSuccess) w->stage_cap = (size_t)w->slabs;
  if (e != cudaSuccess) {
    fail(e, "stb_cuda_sweep_create");
    stb_cuda_sweep_destroy(w);
    return NULL;
  }
  return w;
}

extern "C" int stb_cuda_sweep_tables_in_flight(const stb_sweep_dev_t *w) { return w->T; }
extern "C" int stb_cuda_sweep_tables_per_launch(const stb_sweep_dev_t *w) { return w->slabs; }

extern "C" int stb_cuda_sweep_set_pairs(stb_sweep_dev_t *w, const uint32_t *n, const uint32_t *m, size_t npairs) {
  ON_DEVICE(w->device);
  w->npairs = npairs;
  if (!npairs) return 0;
  if (npairs > w->pairs_cap) {  // buffers are kept across calls (a cached handle sees many pair sets)
    cudaFree(w->d_n);
    cudaFree(w->d_m);
    cudaFree(w->d_partial);
    w->d_n = w->d_m = NULL;
    w->d_partial = NULL;
    w->pairs_cap = 0;
    const size_t nblk = (npairs + 255) / 256;
    CK(cudaMalloc(&w->d_n, npairs * sizeof(uint32_t)));
    CK(cudaMalloc(&w->d_m, npairs * sizeof(uint32_t)));
    CK(cudaMalloc(&w->d_partial, (size_t)w->slabs * nblk * sizeof(double)));
    w->pairs_cap = npairs;
  }
  CK(cudaMemcpyAsync(w->d_n, n, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
  CK(cudaMemcpyAsync(w->d_m, m, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
  // Sums-only runs gather the DISTINCT pairs, in cell order, each weighted by how often it occurs
  // (sampler statistics repeat: config 4 has 94 000 pairs on 58 752 cells).  Only when every pair
  // names a stored cell; a pair outside the table keeps the plain list (its look-up answers -inf).
  w->nuniq = 0;
  bool all_in = true;
  for (size_t i = 0; i < npairs && all_in; i++) all_in = !(m[i] == 0 || n[i] < m[i] || n[i] > w->N || m[i] > w->M);
  if (all_in) {
    const size_t cells = (size_t)w->N * w->ld;
    const unsigned nb = (unsigned)((cells + 255) / 256);
    if (!w->d_cnt) {
      CK(cudaMalloc(&w->d_cnt, cells * sizeof(unsigned)));
      CK(cudaMalloc(&w->d_blkoff, ((size_t)nb + 1) * sizeof(unsigned)));
    }
    if (npairs > w->uniq_cap) {
      cudaFree(w->d_un);
      cudaFree(w->d_um);
      cudaFree(w->d_uw);
      w->d_un = w->d_um = NULL;
      w->d_uw = NULL;
      w->uniq_cap = 0;
      CK(cudaMalloc(&w->d_un, npairs * sizeof(uint32_t)));
      CK(cudaMalloc(&w->d_um, npairs * sizeof(uint32_t)));
      CK(cudaMalloc(&w->d_uw, npairs * sizeof(double)));
      w->uniq_cap = npairs;
    }
    CK(cudaMemsetAsync(w->d_cnt, 0, cells * sizeof(unsigned), w->stream));
    pair_count_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, w->stream>>>(w->d_n, w->d_m, npairs, w->ld, w->d_cnt);
    cell_count_kernel<<<nb, 256, 0, w->stream>>>(w->d_cnt, cells, w->d_blkoff);
    blk_scan_kernel<<<1, 1024, 0, w->stream>>>(w->d_blkoff, nb);
    cell_compact_kernel<<<nb, 256, 0, w->stream>>>(w->d_cnt, cells, w->ld, w->d_blkoff, w->d_un, w->d_um, w->d_uw);
    CK(cudaGetLastError());
    unsigned total = 0;
    CK(cudaMemcpyAsync(&total, w->d_blkoff + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    w->nuniq = total;  // 0 (every pair on the diagonal): the plain list is used
  }
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

extern "C" int stb_cuda_sweep_run(stb_sweep_dev_t *w, const double *a, size_t na, double *gather_out, double *sum_out,
                                  double *lastrow_out, float *fill_ms) {
  return stb_cuda_sweep_run_dealt(w, a, na, 0, 1, gather_out, sum_out, lastrow_out, fill_ms);
}

/*
 * The same for the share of a sweep that was dealt to this device: unit k of this run (k < na) is the
 * caller's unit u(k) = first + k * stride -- its discount is a[u(k)], its results go to row u(k) of the
 * caller's arrays.  (first, stride) = (0, 1) is the whole sweep on one device.
 */
extern "C" int stb_cuda_sweep_run_dealt(stb_sweep_dev_t *w, const double *a, size_t na, size_t first, size_t stride,
                                        double *gather_out, double *sum_out, double *lastrow_out, float *fill_ms) {
  ON_DEVICE(w->device);
  auto unit = [first, stride](size_t k) { return first + k * stride; };
  const size_t es = w->is_float ? 4 : 8;
  const size_t slab_elems = (size_t)w->N * w->ld;
  int nblk_run = 0;
  if ((gather_out || sum_out) && !w->npairs) {
    snprintf(g_err, sizeof g_err, "stb_cuda_sweep_run: no look-up pairs set");
    return -1;
  }
  if (fill_ms) *fill_ms = 0.f;
  if (!na) return 0;
  // tables per launch: every resident slab; with per-pair outputs (a [tables][npairs] staging array) one wave of T
  const size_t chunk = gather_out ? (size_t)w->T : (size_t)w->slabs;
  if (gather_out && w->npairs > w->gather_cap) {
    cudaFree(w->d_gather);
    w->d_gather = NULL;
    w->gather_cap = 0;
    CK(cudaMalloc(&w->d_gather, (size_t)w->T * w->npairs * sizeof(double)));
    w->gather_cap = w->npairs;
  }
  // The waves (T tables each, streamed through the same slabs) are queued back to back: stream order
  // keeps fill -> reduce -> next fill apart, the per-table sums of the whole run collect in d_sum and
  // come back with ONE copy, the watchdog flags of all waves in pinned memory, and the host waits once.
  // (A wave whose results go to pageable caller memory -- gather_out, lastrow_out -- still waits for
  // its own copies.)
  const size_t nwaves = (na + chunk - 1) / chunk;
  if (sum_out && na > w->stage_cap) {
    cudaFree(w->d_sum);
    cudaFreeHost(w->h_stage);
    w->d_sum = NULL;
    w->h_stage = NULL;
    w->stage_cap = 0;
    CK(cudaMalloc(&w->d_sum, na * sizeof(double)));
    CK(cudaHostAlloc(&w->h_stage, na * sizeof(double), cudaHostAllocDefault));
    w->stage_cap = na;
  }
  if (nwaves > w->flags_cap) {
    if (w->h_flags) cudaFreeHost(w->h_flags);
    w->h_flags = NULL;
    w->flags_cap = 0;
    CK(cudaHostAlloc(&w->h_flags, nwaves * sizeof(int), cudaHostAllocDefault));
    w->flags_cap = nwaves;
  }
  memset(w->h_flags, 0, nwaves * sizeof(int));
  std::vector<stb::StripTable> tabs(chunk);
  CK(cudaEventRecord(w->ev0, w->stream));
  size_t wave = 0;
  for (size_t j0 = 0; j0 < na; j0 += chunk, ++wave) {
    const int nt = (int)((na - j0 < chunk) ? na - j0 : chunk);
    for (int t = 0; t < nt; t++) {
      tabs[t].tabS = (char *)w->slab + (size_t)t * slab_elems * es;
      tabs[t].tabV = NULL;
      tabs[t].s1 = w->s1 + (size_t)t * w->N;
      tabs[t].a = a[unit(j0 + t)];
    }
    stb::StripFillArgs args;
    args.tables = tabs.data();
    args.ntables = nt;
    args.has_S = 1;
    args.has_V = 0;
    args.is_float = w->is_float;
    args.ld = w->ld;
    args.N = w->N;
    args.M = w->M;
    args.num_sms = w->num_sms;
    args.async_flag = w->h_flags + wave;
    int rc = stb::strip_fill(&w->strip, args, w->stream, w->ev1, g_err, sizeof g_err);
    if (rc) return rc;
    if (gather_out || sum_out) {
      // sums only: the distinct pairs with their multiplicities
      const bool uq = !gather_out && w->nuniq > 0 && !getenv("STB_SWEEP_PLAIN_PAIRS");
      const size_t np = uq ? w->nuniq : w->npairs;
      const uint32_t *pn = uq ? w->d_un : w->d_n, *pm = uq ? w->d_um : w->d_m;
      const double *pw = uq ? w->d_uw : NULL;
      nblk_run = (int)((np + 255) / 256);
      dim3 grid((unsigned)nblk_run, (unsigned)nt);
      if (w->is_float)
        sweep_gather_kernel<float><<<grid, 256, 0, w->stream>>>((const float *)w->slab, w->s1, slab_elems, w->ld, w->N, w->M,
                                                                  pn, pm, pw, np, gather_out ? w->d_gather : NULL,
                                                                  w->d_partial);
      else
        sweep_gather_kernel<double><<<grid, 256, 0, w->stream>>>((const double *)w->slab, w->s1, slab_elems, w->ld, w->N,
                                                                   w->M, pn, pm, pw, np, gather_out ? w->d_gather : NULL,
                                                                   w->d_partial);
      CK(cudaGetLastError());
      if (sum_out) {
        sweep_sum_kernel<<<nt, 256, 0, w->stream>>>(w->d_partial, nblk_run, w->d_sum + j0);
        CK(cudaGetLastError());
      }
      if (gather_out) {
        if (stride == 1)
          CK(cudaMemcpyAsync(gather_out + unit(j0) * w->npairs, w->d_gather, (size_t)nt * w->npairs * sizeof(double),
                             cudaMemcpyDeviceToHost, w->stream));
        else
          CK(cudaMemcpy2DAsync(gather_out + unit(j0) * w->npairs, stride * w->npairs * sizeof(double), w->d_gather,
                               w->npairs * sizeof(double), w->npairs * sizeof(double), (size_t)nt, cudaMemcpyDeviceToHost,
                               w->stream));
        CK(cudaStreamSynchronize(w->stream));
      }
    }
    if (lastrow_out) {
      for (int t = 0; t < nt; t++) {
        const char *row = (const char *)w->slab + ((size_t)t * slab_elems + (size_t)(w->N - 1) * w->ld) * es;
        if (!w->is_float) {
          CK(cudaMemcpyAsync(lastrow_out + unit(j0 + t) * w->M, row, (size_t)w->M * sizeof(double), cudaMemcpyDeviceToHost,
                             w->stream));
        } else {
          std::vector<float> tmp(w->M);
          CK(cudaMemcpyAsync(tmp.data(), row, (size_t)w->M * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
          CK(cudaStreamSynchronize(w->stream));
          for (unsigned c = 0; c < w->M; c++) lastrow_out[unit(j0 + t) * w->M + c] = (double)tmp[c];
        }
      }
      CK(cudaStreamSynchronize(w->stream));
    }
  }
  CK(cudaEventRecord(w->ev1, w->stream));
  if (sum_out) CK(cudaMemcpyAsync(w->h_stage, w->d_sum, na * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  for (size_t i = 0; i < nwaves; i++)
    if (w->h_flags[i]) {
      snprintf(g_err, sizeof g_err, "stb_cuda_sweep_run: pipeline watchdog fired in wave %zu of %zu", i, nwaves);
      return -2;
    }
  if (sum_out)
    for (size_t j = 0; j < na; j++) sum_out[unit(j)] = w->h_stage[j];
  if (fill_ms) CK(cudaEventElapsedTime(fill_ms, w->ev0, w->ev1));
  return 0;
}

extern "C" void *stb_cuda_host_alloc(size_t bytes) {
  void *p = NULL;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return NULL;
  }
  return p;
}

extern "C" void stb_cuda_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------
// samplea2 over many chains: the seat-partition step against one table per chain (lib/samplea.c:290-321)
// ---------------------------------------------------------------------------------------------
/* glibc's 48-bit generator k steps on: X -> A^k X + C (A^k - 1)/(A - 1) mod 2^48 by repeated squaring */
__host__ __device__ static inline unsigned long long lcg48_jump(unsigned long long x, unsigned long long k) {
  unsigned long long am = 1, ap = 0, cm = 0x5DEECE66DULL, cp = 0xBULL;
  while (k) {
    if (k & 1) {
      am = am * cm;
      ap = ap * cm + cp;
    }
    cp = (cm + 1) * cp;
    cm = cm * cm;
    k >>= 1;
  }
  return (am * x + ap) & 0xFFFFFFFFFFFFULL;
}
extern "C" uint64_t stb_cuda_lcg48_jump(uint64_t x, uint64_t k) { return lcg48_jump(x, k); }

/*
 * partition_kernel's walk for the nodes of MANY chains at once: blockIdx.y is the chain's table of the wave
 * (its slab, its discount, its stream), the sizes go into the chain's histogram of likelihood arguments
 * j = size - 1 (integer counts: the order the threads arrive in does not matter) instead of a list, and the
 * log-uniforms are made here -- the chain's stream is glibc's 48-bit generator, draw number q is the state
 * q + 1 steps on, reached by a jump and then stepped round by round.  (log is CUDA's here and glibc's in the
 * scalar samplea2: a walk would have to stop within an ulp of a log-uniform for the two to differ.)
 */
template <typename T>
__global__ void partition_hist_kernel(const T *__restrict__ slabs, const double *__restrict__ s1s, size_t slab_elems, size_t ld,
                                      unsigned Nrows, const stb::StripTable *__restrict__ tables,
                                      const unsigned long long *__restrict__ x0, const uint32_t *__restrict__ n,
                                      const uint16_t *__restrict__ t, const uint32_t *__restrict__ draw,
                                      const uint32_t *__restrict__ order, size_t count, int exact, uint32_t *__restrict__ hist,
                                      unsigned hbins) {
  const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= count) return;
  const int tb = blockIdx.y;
  const T *tab = slabs + (size_t)tb * slab_elems;
  const double *s1 = s1s + (size_t)tb * Nrows;
  const double a = tables[tb].a;
  uint32_t *h = hist + (size_t)tb * hbins;
  const size_t i = order[slot];
  int N = (int)n[i];
  const int t0 = (int)t[i];
  auto S = [&](int nn, int mm) -> double {
    if (nn == mm) return 0.0;
    if (mm == 1) return s1[nn - 1];
    return (double)tab[(size_t)(nn - 1) * ld + (mm - 1)];
  };
  unsigned long long xs = lcg48_jump(x0[tb], (unsigned long long)draw[i] + 1);
  auto logu = [&]() -> double { return log((double)xs * (1.0 / 281474976710656.0)); };
  double ptot = S(N, t0);
  double rem = exact ? logu() : __dadd_rn(ptot, logu());
  const double shift = exact ? 1.0 : 0.0;
  int M = t0 - 1, l = 1;
  double fact = 0.0;
  double s_next = S(N - 1, M);
  while (M >= 1) {
    const double s_cur = s_next;
    s_next = S(N - (l + 1 <= N - M ? l + 1 : l), M);
    if (l > 1)
      fact = __dadd_rn(fact, log(__ddiv_rn(__dmul_rn(__dsub_rn((double)l - shift, a), (double)(N - l + 1)),
                                           (double)(l - 1))));
    const double term = __dsub_rn(__dadd_rn(fact, s_cur), ptot);
    bool done = term >= rem;
    if (!done) {
      rem = logminus_dev(rem, term);
      l++;
      if (l > N - M) {
        l = N - M;
        done = true;
      }
    }
    if (done) {
      if (l > 1) atomicAdd(&h[l - 1], 1u);
      N -= l;
      M--;
      l = 1;
      fact = 0.0;
      if (M >= 1) {
        s_next = S(N - 1, M);
        if (exact) {
          ptot = S(N, M + 1);
          xs = (0x5DEECE66DULL * xs + 0xBULL) & 0xFFFFFFFFFFFFULL;
          rem = logu();
        }
      }
    }
  }
  if (N > 1) atomicAdd(&h[N - 1], 1u);  // what is left of the node's customers sits at its last table
}

__global__ void hist_init_kernel(uint32_t *__restrict__ hist, const uint32_t *__restrict__ base, unsigned hbins, size_t rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * hbins) hist[i] = base[i % hbins];
}

/* sum_j hist[chain][j] (lgamma(j + 1 - x) - lgamma(1 - x)): one block per point, partial sums in a fixed order */
__global__ void hist_eval_kernel(const uint32_t *__restrict__ hist, unsigned hbins, const double *__restrict__ x,
                                 const int *__restrict__ chain, double *__restrict__ out) {
  __shared__ double red[256];
  const size_t e = blockIdx.x;
  const double xv = x[e];
  const uint32_t *h = hist + (size_t)chain[e] * hbins;
  const double lg1 = lgamma(1.0 - xv);
  double acc = 0.0;
  for (unsigned j = 1 + threadIdx.x; j < hbins; j += blockDim.x) {
    const uint32_t c = h[j];
    if (c) acc += (double)c * (lgamma((double)j + 1.0 - xv) - lg1);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[e] = red[0];
}

/*
 * The nodes of a statistics set that take part in the partition step (1 < t < n), in the order their uniforms
 * come off a chain's stream; draw[j]: the number of the node's first draw (reference mode: j; exact mode: the
 * t - 1 draws of the nodes before it); hbase[hbins]: the arguments every chain's histogram starts from (the nodes
 * with one table).  Sorted longest first here, like stb_cuda_partition.
 */
extern "C" int stb_cuda_sweep_set_nodes(stb_sweep_dev_t *w, const uint32_t *n, const uint16_t *t, const uint32_t *draw,
                                        size_t count, const uint32_t *hbase, unsigned hbins) {
  ON_DEVICE(w->device);
  std::vector<uint32_t> order, start;
  try {
    order.resize(count ? count : 1);
    start.assign((size_t)w->N + 2, 0);
  } catch (...) {
    snprintf(g_err, sizeof g_err, "stb_cuda_sweep_set_nodes: out of host memory");
    return -1;
  }
  for (size_t i = 0; i < count; i++) {
    if (n[i] > w->N || t[i] > w->M || t[i] < 2 || t[i] >= n[i]) {
      snprintf(g_err, sizeof g_err, "stb_cuda_sweep_set_nodes: node %zu (n=%u, t=%u) outside the tables or not 1 < t < n", i,
               n[i], (unsigned)t[i]);
      return -1;
    }
    start[w->N - n[i] + 1]++;
  }
  for (size_t v = 1; v < start.size(); v++) start[v] += start[v - 1];
  for (size_t i = 0; i < count; i++) order[start[w->N - n[i]]++] = (uint32_t)i;
  if (count > w->pnodes_cap) {
    cudaFree(w->d_pn);
    cudaFree(w->d_pdraw);
    cudaFree(w->d_porder);
    cudaFree(w->d_pt);
    w->d_pn = w->d_pdraw = w->d_porder = NULL;
    w->d_pt = NULL;
    w->pnodes_cap = 0;
    CK(cudaMalloc(&w->d_pn, count * sizeof(uint32_t)));
    CK(cudaMalloc(&w->d_pdraw, count * sizeof(uint32_t)));
    CK(cudaMalloc(&w->d_porder, count * sizeof(uint32_t)));
    CK(cudaMalloc(&w->d_pt, count * sizeof(uint16_t)));
    w->pnodes_cap = count;
  }
  if (hbins != w->hbins || !w->d_hbase) {
    cudaFree(w->d_hbase);
    cudaFree(w->d_hist);
    w->d_hbase = w->d_hist = NULL;
    w->hist_cap = 0;
    w->hbins = 0;
    CK(cudaMalloc(&w->d_hbase, (size_t)(hbins ? hbins : 1) * sizeof(uint32_t)));
    w->hbins = hbins;
  }
  if (!w->d_x0) CK(cudaMalloc(&w->d_x0, (size_t)w->slabs * sizeof(unsigned long long)));
  w->pcount = count;
  if (count) {
    CK(cudaMemcpyAsync(w->d_pn, n, count * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(w->d_pdraw, draw, count * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(w->d_porder, order.data(), count * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(w->d_pt, t, count * sizeof(uint16_t), cudaMemcpyHostToDevice, w->stream));
  }
  CK(cudaMemcpyAsync(w->d_hbase, hbase, (size_t)hbins * sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

/*
 * One table per chain at its discount a[c] (waves of the sweep's resident slabs), the partition step of every
 * node against it, chain c drawing from the stream state x0[c].  The histograms stay on the device for
 * stb_cuda_sweep_hist_eval; hist_out (host, [na][hbins], may be NULL) receives a copy.  *ms: device time.
 */
extern "C" int stb_cuda_sweep_partition(stb_sweep_dev_t *w, const double *a, size_t na, const uint64_t *x0, int exact,
                                        uint32_t *hist_out, float *ms) {
  ON_DEVICE(w->device);
  const size_t es = w->is_float ? 4 : 8;
  const size_t slab_elems = (size_t)w->N * w->ld;
  if (ms) *ms = 0.f;
  if (!na) return 0;
  if (!w->d_hbase || !w->hbins) {
    snprintf(g_err, sizeof g_err, "stb_cuda_sweep_partition: no nodes set");
    return -1;
  }
  if (na > w->hist_cap) {
    cudaFree(w->d_hist);
    w->d_hist = NULL;
    w->hist_cap = 0;
    CK(cudaMalloc(&w->d_hist, na * (size_t)w->hbins * sizeof(uint32_t)));
    w->hist_cap = na;
  }
  const size_t chunk = (size_t)w->slabs;
  const size_t nwaves = (na + chunk - 1) / chunk;
  if (nwaves > w->flags_cap) {
    if (w->h_flags) cudaFreeHost(w->h_flags);
    w->h_flags = NULL;
    w->flags_cap = 0;
    CK(cudaHostAlloc(&w->h_flags, nwaves * sizeof(int), cudaHostAllocDefault));
    w->flags_cap = nwaves;
  }
  memset(w->h_flags, 0, nwaves * sizeof(int));
  std::vector<stb::StripTable> tabs(chunk);
  CK(cudaEventRecord(w->ev0, w->stream));
  {
    const size_t tot = na * (size_t)w->hbins;
    hist_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, w->stream>>>(w->d_hist, w->d_hbase, w->hbins, na);
    CK(cudaGetLastError());
  }
  cudaEvent_t ev_fill_end = w->ev1;
  size_t wave = 0;
  for (size_t j0 = 0; j0 < na; j0 += chunk, ++wave) {
    const int nt = (int)((na - j0 < chunk) ? na - j0 : chunk);
    for (int t = 0; t < nt; t++) {
      tabs[t].tabS = (char *)w->slab + (size_t)t * slab_elems * es;
      tabs[t].tabV = NULL;
      tabs[t].s1 = w->s1 + (size_t)t * w->N;
      tabs[t].a = a[j0 + t];
    }
    stb::StripFillArgs args;
    args.tables = tabs.data();
    args.ntables = nt;
    args.has_S = 1;
    args.has_V = 0;
    args.is_float = w->is_float;
    args.ld = w->ld;
    args.N = w->N;
    args.M = w->M;
    args.num_sms = w->num_sms;
    args.async_flag = w->h_flags + wave;
    int rc = stb::strip_fill(&w->strip, args, w->stream, ev_fill_end, g_err, sizeof g_err);
    if (rc) return rc;
    if (w->pcount) {
      CK(cudaMemcpyAsync(w->d_x0, x0 + j0, (size_t)nt * sizeof(unsigned long long), cudaMemcpyHostToDevice, w->stream));
      dim3 grid((unsigned)((w->pcount + 63) / 64), (unsigned)nt);
      uint32_t *hist = w->d_hist + j0 * (size_t)w->hbins;
      if (w->is_float)
        partition_hist_kernel<float><<<grid, 64, 0, w->stream>>>((const float *)w->slab, w->s1, slab_elems, w->ld, w->N,
                                                                  w->strip.tables, w->d_x0, w->d_pn, w->d_pt, w->d_pdraw,
                                                                  w->d_porder, w->pcount, exact, hist, w->hbins);
      else
        partition_hist_kernel<double><<<grid, 64, 0, w->stream>>>((const double *)w->slab, w->s1, slab_elems, w->ld, w->N,
                                                                   w->strip.tables, w->d_x0, w->d_pn, w->d_pt, w->d_pdraw,
                                                                   w->d_porder, w->pcount, exact, hist, w->hbins);
      CK(cudaGetLastError());
    }
    // (the next wave's table list overwrites strip.tables: stream order keeps it behind this wave's kernel;
    // the host copy `tabs` is consumed by strip_fill's cudaMemcpyAsync from pageable memory before it returns)
  }
  CK(cudaEventRecord(w->ev1, w->stream));
  if (hist_out)
    CK(cudaMemcpyAsync(hist_out, w->d_hist, na * (size_t)w->hbins * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  for (size_t v = 0; v < nwaves; v++)
    if (w->h_flags[v]) {
      snprintf(g_err, sizeof g_err, "stb_cuda_sweep_partition: pipeline watchdog fired in wave %zu", v);
      return -2;
    }
  if (ms) CK(cudaEventElapsedTime(ms, w->ev0, w->ev1));
  return 0;
}

/* out[e] = sum_j hist[chain[e]][j] (lgamma(j + 1 - x[e]) - lgamma(1 - x[e])) over the histograms of the last partition */
extern "C" int stb_cuda_sweep_hist_eval(stb_sweep_dev_t *w, const double *x, const int *chain, size_t cnt, double *out,
                                        float *ms) {
  ON_DEVICE(w->device);
  if (ms) *ms = 0.f;
  if (!cnt) return 0;
  if (!w->d_hist) {
    snprintf(g_err, sizeof g_err, "stb_cuda_sweep_hist_eval: no histograms");
    return -1;
  }
  for (size_t e = 0; e < cnt; e++)
    if (chain[e] < 0 || (size_t)chain[e] >= w->hist_cap) {
      snprintf(g_err, sizeof g_err, "stb_cuda_sweep_hist_eval: chain index out of range");
      return -1;
    }
  if (cnt > w->eval_cap) {
    cudaFree(w->d_ex);
    cudaFree(w->d_eout);
    cudaFree(w->d_echain);
    w->d_ex = w->d_eout = NULL;
    w->d_echain = NULL;
    w->eval_cap = 0;
    CK(cudaMalloc(&w->d_ex, cnt * sizeof(double)));
    CK(cudaMalloc(&w->d_eout, cnt * sizeof(double)));
    CK(cudaMalloc(&w->d_echain, cnt * sizeof(int)));
    w->eval_cap = cnt;
  }
  CK(cudaMemcpyAsync(w->d_ex, x, cnt * sizeof(double), cudaMemcpyHostToDevice, w->stream));
  CK(cudaMemcpyAsync(w->d_echain, chain, cnt * sizeof(int), cudaMemcpyHostToDevice, w->stream));
  CK(cudaEventRecord(w->ev0, w->stream));
  hist_eval_kernel<<<(unsigned)cnt, 256, 0, w->stream>>>(w->d_hist, w->hbins, w->d_ex, w->d_echain, w->d_eout);
  CK(cudaGetLastError());
  CK(cudaEventRecord(w->ev1, w->stream));
  CK(cudaMemcpyAsync(out, w->d_eout, cnt * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  if (ms) CK(cudaEventElapsedTime(ms, w->ev0, w->ev1));
  return 0;
}
