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

/*
 * psi^(n)(x), x > 0, any order n >= 0 (MLpsigamma; lib/polygamma.c:502-523 evaluates Amos' algorithm 610 at
 * any order): psi^(n)(x) = (-1)^(n+1) n! zeta(n+1, x), the Hurwitz zeta function by direct summation up to
 * x + k >= 15 + n and the Euler-Maclaurin tail
 *   zeta(s, X) = X^(1-s)/(s-1) + X^-s/2 + sum_j B_2j/(2j)! s(s+1)...(s+2j-2) X^-(s+2j-1).
 * Orders 0..3 go to the dedicated series above.
 */
STB_SF_HD double stb_polygamma(int n, double x) {
  const double b2j[7] = {1.0 / 6.0, -1.0 / 30.0, 1.0 / 42.0, -1.0 / 30.0, 5.0 / 66.0, -691.0 / 2730.0, 7.0 / 6.0};
  double head = 0.0, s, xi, term, poch, fact2j, tail, nfact = 1.0;
  int j, k;
  if (n < 0) return NAN;
  if (n == 0) return stb_digamma(x);
  if (n == 1) return stb_trigamma(x);
  if (n == 2) return stb_tetragamma(x);
  if (n == 3) return stb_pentagamma(x);
  s = (double)n + 1.0;
  while (x < 15.0 + n) {
    head += pow(x, -s);
    x += 1.0;
  }
  xi = 1.0 / x;
  tail = pow(x, 1.0 - s) / (s - 1.0) + 0.5 * pow(x, -s);
  term = pow(x, -s) * xi; /* X^-(s+1) */
  poch = s;               /* s (s+1) ... (s+2j-2) */
  fact2j = 2.0;           /* (2j)! */
  for (j = 1; j <= 7; j++) {
    tail += b2j[j - 1] / fact2j * poch * term;
    poch *= (s + 2.0 * j - 1.0) * (s + 2.0 * j);
    fact2j *= (2.0 * j + 1.0) * (2.0 * j + 2.0);
    term *= xi * xi;
  }
  for (k = 2; k <= n; k++) nfact *= (double)k;
  return ((n & 1) ? 1.0 : -1.0) * nfact * (head + tail);
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
