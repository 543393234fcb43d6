/*
 * sympoly.h -- elementary symmetric polynomials e_h(x_0..x_{K-1}) and a sampler of H-subsets weighted by their
 * products.  Source-compatible with the reference's lib/sympoly.h (:59 sympoly, :72 sympoly_sample,
 * SYMPOLY_MAX :32): same names, arguments and return conventions.  Host code (O(K*H) problems with K <= 32:
 * nothing for a GPU to do); the implementation is csrc/sympoly.c.
 */
#ifndef STB_SYMPOLY_H
#define STB_SYMPOLY_H

#include <stdint.h>

#include "srng.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SYMPOLY_MAX 10 /* side of the sampler's on-stack table; larger problems use the heap */

/*
 * res[0] = 1, res[h] = e_h(val[0..K-1]) exp(-*overflow) for h = 1 .. min(BK, K): values above 1 are divided
 * out of the recursion and their logarithms collected in *overflow, which is folded back (and reset to 0) when
 * it stays below 15.  res needs min(BK + 1, K) + 1 entries (at least 3): like the reference, the recursion
 * carries one entry more than asked for, and that last entry is scratch.  Returns 0.
 */
int sympoly(int K, int BK, double *val, double *res, double *overflow);

/*
 * A subset of exactly H of the K items, drawn with probability proportional to the product of its values;
 * returned as a bit vector (bit k set: item k chosen; K <= 32).  0 when H > K, K == 0 or H == 0.
 * Uniforms come from rng_unit(rng) (srng.h), one per decision, in the reference's order.
 */
uint32_t sympoly_sample(int K, int H, double *val, rngp_t rng);

#ifdef __cplusplus
}
#endif
#endif
