/*
 * fill_mirror.cuh -- table fill in the reference's own operation order (S_MIRROR_ORDER).
 *
 * This kernel restates, cell for cell, the arithmetic of S_remake_part (lib/stable.c:356-388 for
 * log S, :451-482 for V) with IEEE round-to-nearest intrinsics so that nvcc can not contract
 * multiplies and adds into FMAs:
 *   - V uses only + - * /  =>  bit-identical to the CPU library.
 *   - S goes through logadd (lib/stable.c:95-103): max + log(1.0 + exp(min - max)), plain log,
 *     not log1p.  CUDA's log/exp differ from glibc's by <= 1 ulp, so S agrees to a few ulps.
 * It is the parity gate for the fast kernel in fill_strip.cuh and the "exact" mode of the
 * product; it is latency-bound by design (one dependent exp->log chain per row) and is not the
 * throughput path.
 *
 * Geometry: ONE CTA, columns dealt cyclically to threads, one __syncthreads per row.  The live
 * FP64 rows (previous / current) ping-pong in a small global scratch that stays in L1/L2, so
 * float storage (S_FLOAT) still computes in FP64 exactly like SfrontN in lib/stable.c:426-448.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stb {

__device__ __forceinline__ double mirror_logadd(double V, double lp) {
  // lib/stable.c:95-103
  if (lp > V) {
    double t = lp;
    lp = V;
    V = t;
  }
  return __dadd_rn(V, log(__dadd_rn(1.0, exp(__dsub_rn(lp, V)))));
}

template <typename OutT>
__device__ __forceinline__ void store_cell(OutT *p, double v) {
  *p = (OutT)v;
}

/*
 * tabS/tabV: [N][ld] (OutT), either may be null.  s1: N doubles, s1[n-1] = log S^n_1 in the
 * reference's running-sum order (computed by the caller).  scratch: 4*(M+2) doubles.
 */
template <typename OutT>
__global__ void __launch_bounds__(1024, 1)
fill_mirror_kernel(OutT *__restrict__ tabS, OutT *__restrict__ tabV, const double *__restrict__ s1,
                   double *scratch, size_t ld, unsigned N, unsigned M, double a) {
  const unsigned tid = threadIdx.x, nt = blockDim.x;
  double *Sp = scratch, *Sc = scratch + (M + 2);
  double *Vp = scratch + 2 * (size_t)(M + 2), *Vc = scratch + 3 * (size_t)(M + 2);
  const double twoa = __dmul_rn(2.0, a);

  // row n = 1:  S^1_1 = 1 -> log 0 ; no V entry
  if (tid == 0) {
    Sp[1] = s1[0];
    if (tabS) store_cell(tabS + 0, s1[0]);
  }
  __syncthreads();

  for (unsigned n = 2; n <= N; n++) {
    const double nd = (double)n, nm1 = (double)(n - 1);
    const unsigned topS = (n - 1 < M) ? n - 1 : M;  // last off-diagonal S column of this row
    const unsigned topV = (n < M) ? n : M;          // last V column (diagonal included)
    OutT *rowS = tabS ? tabS + (size_t)(n - 1) * ld : nullptr;
    OutT *rowV = tabV ? tabV + (size_t)(n - 1) * ld : nullptr;
    for (unsigned m = tid + 1; m <= topV || m <= topS + 1; m += nt) {
      if (tabS) {
        if (m == 1) {
          double v = s1[n - 1];
          Sc[1] = v;
          store_cell(rowS + 0, v);
        } else if (m <= topS) {
          double coef;
          if (n == 3)
            coef = __dsub_rn(2.0, twoa);  // lib/stable.c:374  log(2-2*a)
          else
            coef = __dsub_rn(__dsub_rn(nd, __dmul_rn((double)m, a)), 1.0);  // N-M*a-1.0
          // the diagonal of the previous row is stored as +0.0, which is what the reference
          // substitutes literally when M==N-1 (lib/stable.c:385)
          double v = mirror_logadd(__dadd_rn(log(coef), Sp[m]), Sp[m - 1]);
          Sc[m] = v;
          store_cell(rowS + (m - 1), v);
        } else if (m == n && m <= M) {
          Sc[m] = 0.0;  // log S^n_n
          store_cell(rowS + (m - 1), 0.0);
        }
      }
      if (tabV && m >= 2 && m <= topV) {
        double v;
        if (n == 2) {
          v = __ddiv_rn(1.0, __dsub_rn(1.0, a));  // lib/stable.c:469
        } else if (m == 2) {
          // (1.0+(N-1-2*a)*V[N-3][0])/(N-1-a)   lib/stable.c:476
          v = __ddiv_rn(__dadd_rn(1.0, __dmul_rn(__dsub_rn(nm1, twoa), Vp[2])), __dsub_rn(nm1, a));
        } else {
          // lib/stable.c:478-480
          double num = (m < n) ? __dmul_rn(__dsub_rn(nm1, __dmul_rn((double)m, a)), Vp[m]) : 0.0;
          double den = __dadd_rn(__ddiv_rn(1.0, Vp[m - 1]),
                                 __dsub_rn(nm1, __dmul_rn((double)(m - 1), a)));
          v = __ddiv_rn(__dadd_rn(1.0, num), den);
        }
        Vc[m] = v;
        store_cell(rowV + (m - 1), v);
      }
    }
    __syncthreads();
    double *t = Sp;
    Sp = Sc;
    Sc = t;
    t = Vp;
    Vp = Vc;
    Vc = t;
  }
}

}  // namespace stb
