/*
 * specfun.h -- digamma, trigamma, tetragamma, pentagamma and the inverse digamma as FP64
 * host/device functions.
 *
 * The reference's slice-sampler build takes digamma/trigamma from Mathlib's dpsifn (Amos algorithm
 * 610, lib/polygamma.c:163-563) and digammaInv from Minka's Newton iteration
 * (lib/digammainv.c:27-38); its default build uses Neal's digammaRN (lib/digamma.c:31-48).  These
 * are independent implementations of the same functions (upward recurrence to x >= 10, then the
 * Bernoulli asymptotic series to x^-14 / x^-15): |rel. error| ~ 1e-15 for x > 0, which is what
 * the parity tests compare against the reference at 1e-12.
 */
#ifndef STB_SPECFUN_H
#define STB_SPECFUN_H
#include <math.h>

#ifdef __CUDACC__
#define STB_SF_HD __host__ __device__ static inline
#else
#define STB_SF_HD static inline
#endif

/* psi(x), x > 0 */
STB_SF_HD double stb_digamma(double x) {
  double r = 0.0, f, t;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  f = 1.0 / (x * x);
  t = f * (-1.0 / 12.0 +
           f * (1.0 / 120.0 +
                f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

/* psi'(x), x > 0 */
STB_SF_HD double stb_trigamma(double x) {
  double r = 0.0, f, xi, t;
  while (x < 10.0) {
    r += 1.0 / (x * x);
    x += 1.0;
  }
  xi = 1.0 / x;
  f = xi * xi;
  t = xi * f *
      (1.0 / 6.0 +
       f * (-1.0 / 30.0 +
            f * (1.0 / 42.0 + f * (-1.0 / 30.0 + f * (5.0 / 66.0 + f * (-691.0 / 2730.0 + f * (7.0 / 6.0)))))));
  return r + xi + 0.5 * f + t;
}

/* d^2 psi / dx^2, x > 0: f(x) = f(x+1) - 2/x^3 up to x >= 15, then
 * -1/x^2 - 1/x^3 - sum (2k+1) B_2k / x^(2k+2) */
STB_SF_HD double stb_tetragamma(double x) {
  double r = 0.0, f, xi, t;
  while (x < 15.0) {
    r -= 2.0 / (x * x * x);
    x += 1.0;
  }
  xi = 1.0 / x;
  f = xi * xi;
  t = f * f *
      (1.0 / 2.0 +
       f * (-1.0 / 6.0 + f * (1.0 / 6.0 + f * (-3.0 / 10.0 + f * (5.0 / 6.0 + f * (-691.0 / 210.0 + f * (35.0 / 2.0)))))));
  return r - f - f * xi - t;
}

/* d^3 psi / dx^3, x > 0: f(x) = f(x+1) + 6/x^4, then 2/x^3 + 3/x^4 + sum (2k+1)(2k+2) B_2k / x^(2k+3) */
STB_SF_HD double stb_pentagamma(double x) {
  double r = 0.0, f, xi, t;
  while (x < 15.0) {
    const double x2 = x * x;
    r += 6.0 / (x2 * x2);
    x += 1.0;
  }
  xi = 1.0 / x;
  f = xi * xi;
  t = f * f * xi *
      (2.0 + f * (-1.0 + f * (4.0 / 3.0 + f * (-3.0 + f * (10.0 + f * (-691.0 / 15.0 + f * 280.0))))));
  return r + 2.0 * f * xi + 3.0 * f * f + t;
}

/* Neal's digamma as the reference's default build defines it (lib/digamma.c:31-48): recurrence
 * to x > 5, eight-term series */
STB_SF_HD double stb_digammaRN(double x) {
  double r = 0.0, f, t;
  while (x <= 5) {
    r -= 1 / x;
    x += 1;
  }
  f = 1 / (x * x);
  t = f * (-1 / 12.0 +
           f * (1 / 120.0 +
                f * (-1 / 252.0 +
                     f * (1 / 240.0 + f * (-1 / 132.0 + f * (691 / 32760.0 + f * (-1 / 12.0 + f * 3617 / 8160.0)))))));
  return r + log(x) - 0.5 / x + t;
}

/* x with psi(x) = y: Minka's starting point and five Newton steps (lib/digammainv.c:27-38) */
STB_SF_HD double stb_digamma_inv(double y) {
  double g;
  int i;
  if (y < -2.22)
    g = -1 / (y - stb_digamma(1.0));
  else
    g = exp(y) + 0.5;
  for (i = 0; i < 5; i++) g -= (stb_digamma(g) - y) / stb_trigamma(g);
  return g;
}

#endif
