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
// The logarithm's table holds ONLY log(c_i) (8 bytes): 1/c_i is taken from the hardware's
// single-precision reciprocal of c_i (exactly representable, so the result is a fixed function of
// i), widened to double with integer operations; the table is built on the device with the same
// instruction, T[i] = -log(rcp(c_i)), which makes log(mant) = T[i] + log1p(mant*rcp(c_i) - 1) an
// identity whatever the last bit of the reciprocal is.  In shared memory the table is kept
// LOGTAB_REP8 times, interleaved (entry i of copy j at [i*16 + j]): copy j occupies only 8-byte
// bank j and lane l reads copy l % 16, so the sixteen lanes a shared-memory load serves per pass
// never collide whatever their mantissas are.  (A single copy of a 16-byte-entry table cost ~2.7
// passes per load at random indices and was the largest user of shared-memory bandwidth of the fill.)
// The index is formed from the high word of mant = (hi & 0xFFFFF) | 0x3FF00000 as it stands:
// (mhi + 0x800) >> 12 = LOGTAB_IDX0 + i serves as table index, and as the reciprocal's argument, with the
// constants folded into the table's base address and the float's bit pattern.  (Tried: the interval's MIDPOINT
// instead of the nearest c_i -- no rounding add, one instruction per cell less -- but a power of two then no
// longer comes out as an exact multiple of ln2, and log S^2_1 = log 1 at a = 0 must print as 0 like the
// reference's; the instruction made no measurable difference.)
constexpr int LOGTAB_REP8 = 16;
constexpr unsigned LOGTAB_IDX0 = 0x3FF00u;  // ((high word of a double in [1,2)) + 0x800) >> 12 = LOGTAB_IDX0 + i

__device__ __forceinline__ double logtab_widen(float invf) {  // a positive normal float as a double: move the bits
  const unsigned fb = __float_as_uint(invf);
  return __hiloint2double((int)((fb >> 3) + 0x38000000u), (int)(fb << 29));
}
/* rcp(c_i) from idxh = LOGTAB_IDX0 + i: the float c_i has the bits 0x3F800000 + (i << 15) */
__device__ __forceinline__ double logtab_inv_h(unsigned idxh) {
  // (tried: the double-precision reciprocal seed MUFU.RCP64H on c_i's high word -- one instruction less than widening the
  // float -- 3 % SLOWER on config 2: 9.27 against 8.96 ms)
  float invf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invf) : "f"(__uint_as_float((idxh << 15) + (0x3F800000u - (LOGTAB_IDX0 << 15)))));
  return logtab_widen(invf);
}
constexpr long long FILL_WATCHDOG = 6000000000LL;  // cycles a wait may last before the fill aborts

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

constexpr double LOG_EBIAS = 4503601774854144.0;  // 2^52 + 2^31

/*
 * x / d for normal positive operands, no branch: the hardware seed (rcp.approx.ftz.f64: the upper 20
 * mantissa bits), ONE Newton step (2^-40), the quotient and one residual correction -- the correction
 * uses the exact residual x - d q, so its own error is second order (2^-80): the result is within an
 * ulp of x / d.  Five FP64 instructions.
 */
__device__ __forceinline__ double div_pos(double x, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  const double q = x * r;
  return fma(fma(-d, q, x), r, q);
}

/*
 * log(x * 2^E) for x > 0 finite normal, E given as kbias = E + 0x80000000 - 1023 (mod 2^32): then
 * (hi >> 20) + kbias is the low word of the double 2^52 + 2^31 + (E + k), k the exponent of x, and
 * one subtraction of that constant gives E + k exactly (|E + k| < 2^31).
 * Table-driven: x = 2^k * mant, mant in [1,2); c = 1 + i/256 nearest to mant;
 * r = mant*rcp(c) - 1 (|r| <= 2^-9 + 2^-23, one FMA); log(mant) = log1p(r) + T[i]; result =
 * (E+k) ln2 + T[i] + log1p(r).  mant == 1 gives exactly (E+k) ln2, so S^n_n comes out as +0.0.
 * tab_s: SHARED address of this lane's copy of the table (entries REP apart) minus LOGTAB_IDX0 entries, so
 * that the index needs no offset.
 */
template <int REP>
__device__ __forceinline__ double log_scaled_r(double x, unsigned kbias, unsigned tab_s) {
  const unsigned hi = (unsigned)__double2hiint(x);
  // (hi & 0xFFFFF) | 0x3FF00000 as ONE three-input logic instruction, the second constant from a register (the
  // compiler splits the expression into two instructions with an immediate each)
  unsigned mhi, one_hi;
  asm("mov.u32 %0, 0x3FF00000;" : "=r"(one_hi));
  asm("lop3.b32 %0, %1, 0xFFFFF, %2, 0xEA;" : "=r"(mhi) : "r"(hi), "r"(one_hi));
  const unsigned idxh = (mhi + 0x800u) >> 12;
  const double mant = __hiloint2double((int)mhi, __double2loint(x));
  double log_c;
  asm("ld.shared.f64 %0, [%1];" : "=d"(log_c) : "r"(tab_s + idxh * (unsigned)(REP * 8)));
  const double r = fma(mant, logtab_inv_h(idxh), -1.0);
  // log1p(r) = r (1 - r/2 + r^2/3 - r^3/4) + O(r^5/5), |r| <= 2^-9 + 2^-23: truncation < 6e-15
  // absolute, far inside the 1e-12 bar; seven FP64 instructions in all
  double t = fma(r, -0.25, 1.0 / 3.0);
  t = fma(r, t, -0.5);
  t = fma(r, t, 1.0);
  const double Ek = __hiloint2double(0x43300000, (int)((hi >> 20) + kbias)) - LOG_EBIAS;
  return fma(r, t, fma(Ek, 0.693147180559945309417232, log_c));
}

/* builds the 8-byte table on the device: one thread per entry */
__global__ void logtab8_build_kernel(double *tab) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i >= LOGTAB_N) return;
  const double inv = logtab_inv_h(LOGTAB_IDX0 + (unsigned)i);
  tab[i] = (inv == 1.0) ? 0.0 : -log(inv);
}

template <typename OutT>
__device__ __forceinline__ void st_out(OutT *p, double v) {
  *p = (OutT)v;
}

}  // namespace stb
