/*
 * rng48.h -- the random streams of the reference's samplers, re-implemented so that a chain can
 * carry its OWN state (the reference draws from glibc's global drand48/lrand48, lib/srng.h:28-34,
 * "cannot be used in multi-threaded programs") and so that the same code runs on the host and on
 * the device.
 *
 *   stb_rng48_seed/unit/lrand : glibc's 48-bit LCG  X <- (0x5DEECE66D X + 0xB) mod 2^48,
 *       srand48(s): X = (s << 16) | 0x330E;  drand48() = X / 2^48;  lrand48() = X >> 17.
 *       Bit-identical to glibc (tests/test_samplers_cpu.py compares the streams).
 *   stb_gauss_zig   : ziggurat N(0,1) in the variant of lib/gslrandist.c:194-234 (128 levels,
 *       right-most step R = 3.44428647676, exponential wedge for the tail, level chosen from
 *       lrand48()/scale exactly like :60-79).
 *   stb_gamma, stb_beta : Marsaglia-Tsang gamma with the a<1 boost and beta = g1/(g1+g2),
 *       lib/gslrandist.c:236-282.
 * The ziggurat tables are not copied from anywhere: stb_zig_tables_build() constructs them from
 * the method's defining equations (equal-area levels under exp(-x^2/2), v = f(R)(R + 1/R), R
 * solved from the condition that the top level closes at height 1) and rounds them to 12
 * significant digits, which is how the published constants are printed; the resulting Gaussian
 * stream is bit-identical to the reference's (tests/test_samplers_cpu.py).
 *
 * Plain C99 / CUDA: every function is STB_HD static inline.
 */
#ifndef STB_RNG48_H
#define STB_RNG48_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define STB_HD __host__ __device__ static inline
#else
#define STB_HD static inline
#endif

#define STB_ZIG_R 3.44428647676
#define STB_ZIG_N 128

typedef struct stb_zig_tables {
  double ytab[STB_ZIG_N];        /* height of level i: exp(-x_i^2/2), x_0 = 0, x_127 = R */
  double wtab[STB_ZIG_N];        /* 2^-24 x_{i+1} */
  unsigned long ktab[STB_ZIG_N]; /* floor(2^24 x_i / x_{i+1}) */
} stb_zig_tables;

typedef struct stb_rng48 {
  uint64_t x; /* 48-bit state */
} stb_rng48;

STB_HD void stb_rng48_seed(stb_rng48 *r, long seed) { r->x = (((uint64_t)(uint32_t)seed) << 16) | 0x330EULL; }
STB_HD uint64_t stb_rng48_next(stb_rng48 *r) {
  r->x = (0x5DEECE66DULL * r->x + 0xBULL) & 0xFFFFFFFFFFFFULL;
  return r->x;
}
/* drand48(): exact, the state has 48 bits */
STB_HD double stb_rng48_unit(stb_rng48 *r) { return (double)stb_rng48_next(r) * (1.0 / 281474976710656.0); }
/* lrand48() */
STB_HD long stb_rng48_lrand(stb_rng48 *r) { return (long)(stb_rng48_next(r) >> 17); }

STB_HD double stb_rng48_unit_pos(stb_rng48 *r) {  /* lib/gslrandist.c:53-58 */
  double u = stb_rng48_unit(r);
  while (u == 0) u = stb_rng48_unit(r);
  return u;
}

/* lib/gslrandist.c:60-79: uniform integer below n (n <= 2^30) from lrand48 */
STB_HD unsigned long stb_rng48_uniform_int(stb_rng48 *r, unsigned long n) {
  const unsigned long range = 1UL << 30;
  unsigned long scale, k;
  if (n > range || n == 0) return 0;
  scale = range / n;
  do {
    k = (unsigned long)stb_rng48_lrand(r) / scale;
  } while (k >= n);
  return k;
}

STB_HD double stb_gauss_zig(stb_rng48 *r, const stb_zig_tables *zt, double sigma) {
  unsigned long i, j;
  int sign;
  double x, y;
  for (;;) {
    i = stb_rng48_uniform_int(r, 256);      /* the step */
    j = stb_rng48_uniform_int(r, 16777216); /* 24 bits inside it */
    sign = (i & 0x80) ? +1 : -1;
    i &= 0x7f;
    x = (double)j * zt->wtab[i];
    if (j < zt->ktab[i]) break;
    if (i < 127) {
      const double y0 = zt->ytab[i], y1 = zt->ytab[i + 1];
      const double u1 = stb_rng48_unit(r);
      y = y1 + (y0 - y1) * u1;
    } else {
      const double u1 = 1.0 - stb_rng48_unit(r);
      const double u2 = stb_rng48_unit(r);
      x = STB_ZIG_R - log(u1) / STB_ZIG_R;
      y = exp(-STB_ZIG_R * (x - 0.5 * STB_ZIG_R)) * u2;
    }
    if (y < exp(-0.5 * x * x)) break;
  }
  return sign * sigma * x;
}

/* Marsaglia-Tsang, a >= 1 */
STB_HD double stb_gamma_ge1(stb_rng48 *r, const stb_zig_tables *zt, double a) {
  double x, v, u;
  const double d = a - 1.0 / 3.0;
  const double c = (1.0 / 3.0) / sqrt(d);
  for (;;) {
    do {
      x = stb_gauss_zig(r, zt, 1.0);
      v = 1.0 + c * x;
    } while (v <= 0);
    v = v * v * v;
    u = stb_rng48_unit_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return d * v;
}

STB_HD double stb_gamma(stb_rng48 *r, const stb_zig_tables *zt, double a) {
  if (a < 1) { /* boost: the uniform is drawn BEFORE the gamma(1+a), as in the reference */
    const double u = stb_rng48_unit_pos(r);
    return stb_gamma_ge1(r, zt, 1.0 + a) * pow(u, 1.0 / a);
  }
  return stb_gamma_ge1(r, zt, a);
}

STB_HD double stb_beta(stb_rng48 *r, const stb_zig_tables *zt, double a, double b) {
  const double x1 = stb_gamma(r, zt, a);
  const double x2 = stb_gamma(r, zt, b);
  return x1 / (x1 + x2);
}

#ifndef __CUDA_ARCH__
#include <stdio.h>
#include <stdlib.h>
/* host only: construct the tables (see the header comment) */
static inline double stb_zig_round12(long double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.12Lg", v);
  return strtod(buf, NULL);
}
/* the level recurrence for a given right-most step R; returns how far the top level is from closing */
static inline long double stb_zig_levels(long double R, long double *x, long double *y) {
  int i;
  const long double fR = expl(-0.5L * R * R);
  const long double v = fR * (R + 1.0L / R); /* area of every level: base strip + exponential wedge */
  x[STB_ZIG_N] = v / fR;                     /* pseudo width of the base strip */
  x[STB_ZIG_N - 1] = R;
  y[STB_ZIG_N - 1] = fR;
  for (i = STB_ZIG_N - 2; i >= 1; i--) {
    y[i] = y[i + 1] + v / x[i + 1];
    x[i] = sqrtl(-2.0L * logl(y[i]));
  }
  x[0] = 0.0L;
  y[0] = 1.0L;
  return y[1] + v / x[1] - 1.0L; /* the top level has area v too, so this is 0 for the right R */
}
static inline void stb_zig_tables_build(stb_zig_tables *zt) {
  long double x[STB_ZIG_N + 1], y[STB_ZIG_N + 1];
  /* R to full precision from the closure condition (STB_ZIG_R is its 12-digit print-out) */
  long double lo = (long double)STB_ZIG_R - 1e-6L, hi = (long double)STB_ZIG_R + 1e-6L;
  int i, it;
  for (it = 0; it < 100; it++) {
    const long double mid = 0.5L * (lo + hi);
    if (stb_zig_levels(mid, x, y) > 0)
      lo = mid;
    else
      hi = mid;
  }
  stb_zig_levels(0.5L * (lo + hi), x, y);
  for (i = 0; i < STB_ZIG_N; i++) {
    zt->ytab[i] = (i == 0) ? 1.0 : stb_zig_round12(y[i]);
    zt->wtab[i] = stb_zig_round12(x[i + 1] / 16777216.0L);
    zt->ktab[i] = (unsigned long)floorl(16777216.0L * x[i] / x[i + 1]);
  }
}
#endif

#endif
