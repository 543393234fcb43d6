/*
 * fill_common.cuh -- device helpers shared by the table-fill kernels (sm_100a).
 *
 * Scaled representation: a table value S is carried as x * 2^E (x a double, E an integer), so
 * the recurrence  S^n_m = (n-1-m a) S^{n-1}_m + S^{n-1}_{m-1}  (lib/stable.c:380-388 restated in
 * the linear domain) is one DFMA per cell and the logarithm is taken once per stored cell:
 * log S = log(x) + E ln2.  Scaling by powers of two is exact, so results do not depend on when
 * or how often a value is renormalised.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace stb {

constexpr int LOGTAB_N = 257;  // c_i = 1 + i/256, i = 0..256
constexpr long long FILL_WATCHDOG = 6000000000LL;  // cycles a wait may last before the fill aborts

struct __align__(16) LogTabEntry {
  double inv_c;  // 1/c_i rounded
  double log_c;  // -log(inv_c) in double
};

__device__ __forceinline__ int ld_vol(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void st_vol(int *p, int v) { *(volatile int *)p = v; }

__device__ __forceinline__ int ld_relaxed_gpu(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double pow2i(int e) {  // 2^e, |e| <= 1022
  return __hiloint2double((e + 1023) << 20, 0);
}

__device__ __forceinline__ double shfl_up_d(double v) {
  int lo = __shfl_up_sync(0xffffffffu, __double2loint(v), 1);
  int hi = __shfl_up_sync(0xffffffffu, __double2hiint(v), 1);
  return __hiloint2double(hi, lo);
}

/*
 * log(x * 2^E) for x > 0 finite normal; Eoff = (double)E - (2^52 + 2^31).
 * Table-driven: x = 2^k * mant, mant in [1,2); c = 1 + i/256 nearest to mant; r = mant/c - 1
 * (|r| <= 2^-9, one FMA); log(mant) = log1p(r) + log(c); result = (E+k) ln2 + log(c) + log1p(r)
 * with E+k formed exactly.  mant == 1 gives exactly (E+k) ln2, so S^n_n comes out as +0.0.
 */
constexpr double LOG_EBIAS = 4503601774854144.0;  // 2^52 + 2^31

__device__ __forceinline__ double log_scaled(double x, double Eoff, const LogTabEntry *tab) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int k = (hi >> 20) - 1023;
  const int frac = hi & 0xFFFFF;
  const int idx = (frac + 0x800) >> 12;
  const double mant = __hiloint2double(frac | 0x3FF00000, lo);
  const double2 tb = *reinterpret_cast<const double2 *>(tab + idx);
  const double r = fma(mant, tb.x, -1.0);
  double t = fma(r, 0.2, -0.25);
  t = fma(r, t, 1.0 / 3.0);
  t = fma(r, t, -0.5);
  const double p = fma(r * r, t, r);
  // (2^52 + 2^31 + k) + (E - 2^52 - 2^31) == E + k exactly
  const double Ek = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)k)) + Eoff;
  return fma(Ek, 0.693147180559945309417232, tb.y + p);
}

/* x / d for normal positive operands: MUFU seed + two Newton steps + one correction, no branch */
__device__ __forceinline__ double div_pos(double x, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  double q = x * r;
  return fma(fma(-d, q, x), r, q);
}

/*
 * Same value with the exponent as an integer: kbias = E + 0x80000000 - 1023 (mod 2^32), so that
 * (hi >> 20) + kbias is the low word of the double 2^52 + 2^31 + (E + k) and one subtraction of
 * the constant gives E + k exactly (|E + k| < 2^31).
 */
__device__ __forceinline__ double log_scaled_i(double x, unsigned kbias, const LogTabEntry *tab) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int frac = hi & 0xFFFFF;
  const int idx = (frac + 0x800) >> 12;
  const double mant = __hiloint2double(frac | 0x3FF00000, lo);
  const double2 tb = *reinterpret_cast<const double2 *>(tab + idx);
  const double r = fma(mant, tb.x, -1.0);
  double t = fma(r, 0.2, -0.25);
  t = fma(r, t, 1.0 / 3.0);
  t = fma(r, t, -0.5);
  const double p = fma(r * r, t, r);
  const double Ek = __hiloint2double(0x43300000, (int)(((unsigned)hi >> 20) + kbias)) - LOG_EBIAS;
  return fma(Ek, 0.693147180559945309417232, tb.y + p);
}

template <typename OutT>
__device__ __forceinline__ void st_out(OutT *p, double v) {
  *p = (OutT)v;
}

/* host: the log table (computed in long double, rounded once) */
inline void logtab_host(LogTabEntry *h) {
  for (int i = 0; i < LOGTAB_N; i++) {
    double c = 1.0 + (double)i / 256.0;
    double inv = 1.0 / c;
    h[i].inv_c = inv;
    h[i].log_c = (double)(-logl((long double)inv));
  }
  h[0].inv_c = 1.0;
  h[0].log_c = 0.0;
}

}  // namespace stb
