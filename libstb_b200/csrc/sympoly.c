/*
 * sympoly.c -- elementary symmetric polynomials and the subset sampler built on them (include/sympoly.h;
 * replaces lib/sympoly.c:62-109 and :128-239).
 *
 * Both work on the coefficients of  P_k(z) = prod_{j<=k} (1 + x_j z)  (e_h = coefficient of z^h), kept in the
 * scaled form  F_k(z) = P_k(z) / prod_{j<=k, x_j>1} x_j : a factor with x > 1 is applied as (1/x + z), one with
 * x <= 1 as (1 + x z), so no coefficient grows with the values above 1; what was divided out is tracked (as the
 * constant coefficient `unit`, and in sympoly as a logarithm).  The first item is taken as it is, like the
 * reference does (:77): results then agree with the reference's to the last bit, and the sampler -- the same
 * comparisons on the same numbers -- consumes the same uniforms and returns the same subsets.
 */
#include "sympoly.h"

#include <math.h>
#include <stdlib.h>

/* one factor applied in place to the coefficients c[1..top] (c[0] is `unit`), highest degree first so that
 * every update still sees the old lower coefficient.  grow: the degree rises by one (c[top] has no old value);
 * otherwise the polynomial is truncated at `top`, whose entry is scratch (it only receives the shift) */
static void apply_factor(double *c, int top, double x, double unit) {
  int h;
  if (x > 1) {
    c[top] = top > 1 ? c[top - 1] : unit;
    for (h = top - 1; h >= 1; h--) c[h] = c[h] / x + (h > 1 ? c[h - 1] : unit);
  } else {
    c[top] = x * (top > 1 ? c[top - 1] : unit);
    for (h = top - 1; h >= 1; h--) c[h] += x * (h > 1 ? c[h - 1] : unit);
  }
}

int sympoly(int K, int BK, double *val, double *res, double *overflow) {
  int B = BK + 1, k, h;
  double unit = 1;
  if (B > K) B = K;
  if (B <= 1) B = 2;
  for (h = K; h >= 1; --h) res[h] = 0.0; /* (the reference clears K entries whatever BK is, lib/sympoly.c:74-75) */
  res[0] = 1.0;
  *overflow = 0.0;
  if (K < 1) return 0;
  res[1] = val[0];
  for (k = 1; k != K; ++k) {
    const double x = val[k];
    apply_factor(res, k + 1 < B ? k + 1 : B, x, unit);
    if (x > 1) {
      unit /= x;
      *overflow += log(x);
    }
  }
  if (*overflow < 15) {
    const double back = exp(*overflow);
    for (h = B; h >= 1; --h) res[h] *= back;
    *overflow = 0.0;
  }
  return 0;
}

/* 1 < H < K: all prefixes' coefficients up to degree H, then the items decided from the last to the first */
static uint32_t sample_by_table(int K, int H, const double *val, rngp_t rng) {
  double stack[SYMPOLY_MAX * SYMPOLY_MAX], *tab = stack, unit = 1;
  uint32_t chosen = 0;
  int k, h;
  (void)rng;
  /* row k: coefficients 1..H of the scaled prefix polynomial over items 0..k, at tab[k*H + h-1] */
  if ((long)K * H >= SYMPOLY_MAX * SYMPOLY_MAX) {
    tab = (double *)malloc(sizeof(double) * (size_t)K * (size_t)H);
    if (!tab) return 0;
  }
  tab[0] = val[0];
  for (k = 1; k != K; ++k) {
    const double x = val[k], *prev = tab + (size_t)(k - 1) * H;
    double *row = tab + (size_t)k * H;
    const int full = k < H ? k : H; /* degrees with an old value */
    for (h = 1; h <= full; h++) {
      const double below = h > 1 ? prev[h - 2] : unit;
      row[h - 1] = x > 1 ? prev[h - 1] / x + below : prev[h - 1] + x * below;
    }
    if (full == k) row[k] = x > 1 ? (k > 0 ? prev[k - 1] : unit) : x * prev[k - 1]; /* the new top degree k+1 */
    if (x > 1) unit /= x;
  }
  /* item k stays out with probability F_{k-1,h} / F_{k,h} (same scaling on both sides once a factor x > 1 is
   * put back on the prefix that lacks it) */
  h = H;
  for (k = K - 1; h > 0 && k >= h; --k) {
    const double with_k = tab[(size_t)k * H + h - 1], without = tab[(size_t)(k - 1) * H + h - 1];
    if (with_k * rng_unit(rng) >= (val[k] <= 1 ? without : without / val[k])) {
      chosen |= (uint32_t)1 << k;
      --h;
    }
  }
  if (tab != stack) free(tab);
  if (h > 0) chosen |= ((uint32_t)1 << h) - 1u; /* as many items left as are still to be chosen: all of them */
  return chosen;
}

uint32_t sympoly_sample(int K, int H, double *val, rngp_t rng) {
  double mass = 0;
  int k = 0;
  if (K < 1 || H < 1 || K < H) return 0;
  if (K == H) return (uint32_t)(((uint64_t)1 << H) - 1);
  if (H != 1) return sample_by_table(K, H, val, rng);
  /* one item, proportional to its value */
  while (k != K) mass += val[k++];
  mass *= rng_unit(rng);
  for (k = 0; mass > 0 && k != K; ++k) mass -= val[k];
  return (uint32_t)1 << (k - 1);
}
