/*
 * stirling_oracle.c -- CPU restatement of the reference's Stirling-table algorithm.
 *
 * TEST INFRASTRUCTURE ONLY (see header).  Plain C, glibc libm, single thread.  Each function
 * restates one piece of wbuntine/libstb and cites it; the arithmetic keeps the reference's
 * operation order so that, with the same libm, results are BIT-IDENTICAL to the compiled
 * reference (tests/test_oracle_vs_reference.py checks exactly that against oracle/_ref, and
 * tests/golden/ pins both against vectors generated from the reference).
 *
 * Parity status: pinned (against the reference itself run in this container + golden vectors).
 */
#include "stirling_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* lib/stable.c:95-103 */
static double orc_logadd(double big, double small) {
  if (small > big) {
    double t = small;
    small = big;
    big = t;
  }
  return big + log(1.0 + exp(small - big));
}

/* lib/stable.c:338-348: S1[n-1] = log S^n_1 as a running sum of log(n-1-a) */
void orc_fill_S1(unsigned N, double a, double *S1) {
  int n;
  S1[0] = 0;
  for (n = 2; n <= (int)N; n++) S1[n - 1] = S1[n - 2] + log(n - 1 - a);
}

/*
 * lib/stable.c:373-388.  Row n holds m=1 (the S1 value), m=2..min(n-1,M) from the
 * recurrence, and the diagonal m=n as 0 (the reference never stores it and substitutes a
 * literal 0, :385).
 */
void orc_fill_S(unsigned N, unsigned M, double a, double *S, size_t ld) {
  int n, m;
  double *S1 = (double *)malloc(sizeof(double) * N);
  orc_fill_S1(N, a, S1);
#define C(n, m) S[(size_t)((n)-1) * ld + ((m)-1)]
  for (n = 1; n <= (int)N; n++) {
    C(n, 1) = S1[n - 1];
    if (n >= 2 && n <= (int)M) C(n, n) = 0;
  }
  if (N >= 3 && M >= 2) C(3, 2) = orc_logadd(S1[1], log(2 - 2 * a));
  for (n = 4; n <= (int)N; n++) {
    if (M >= 2) C(n, 2) = orc_logadd(log(n - 2 * a - 1.0) + C(n - 1, 2), S1[n - 2]);
    for (m = 3; m <= (int)M && m < n; m++)
      C(n, m) = orc_logadd(log(n - m * a - 1.0) + ((m < n - 1) ? C(n - 1, m) : 0), C(n - 1, m - 1));
  }
#undef C
  free(S1);
}

/* lib/stable.c:468-482.  Row n holds m=2..min(n,M), diagonal included. */
void orc_fill_V(unsigned N, unsigned M, double a, double *V, size_t ld) {
  int n, m;
#define C(n, m) V[(size_t)((n)-1) * ld + ((m)-1)]
  if (N >= 2 && M >= 2) C(2, 2) = 1.0 / (1.0 - a);
  for (n = 3; n <= (int)N; n++) {
    if (M >= 2) C(n, 2) = (1.0 + (n - 1 - 2 * a) * C(n - 1, 2)) / (n - 1 - a);
    for (m = 3; m <= (int)M && m <= n; m++)
      C(n, m) = (1.0 + ((m < n) ? ((n - 1 - m * a) * C(n - 1, m)) : 0)) /
                (1.0 / C(n - 1, m - 1) + (n - 1 - (m - 1) * a));
  }
#undef C
}

/* lib/stable.c:1057-1084 */
double orc_asympt(double a, unsigned n, unsigned m) {
  if (a == 0) {
    double ln = log(n);
    return lgamma(n) + (m - 1) * log(ln) - lgamma(m) - lgamma(1 + (m - 1) / ln);
  } else {
    double prod = 0;
    double la1 = lgamma(1.0 - a);
    double aln = a * log((double)n);
    double np = pow(n, -a);
    prod += lgamma((double)n) - la1 - lgamma((double)m) - (m - 1.0) * log(a) - aln;
    if (np < 1e-5)
      prod -= (m - 1) * np * (1 + np * (0.5 + np / 3.0));
    else
      prod += (m - 1) * log(1.0 - np);
    return prod;
  }
}

/* lib/stable.c:905-911 */
double orc_V_asympt(double a, unsigned n, unsigned m) {
  if (a > 0) return (1.0 - pow(n, -a)) / a / (m - 1);
  {
    double ln = log(n);
    return ln / (m - 1) * exp(lgamma(1 + (m - 2) / ln) - lgamma(1 + (m - 1) / ln));
  }
}

orc_table *orc_make(unsigned N, unsigned M, unsigned maxN, unsigned maxM, double a, uint32_t flags) {
  orc_table *t = (orc_table *)calloc(1, sizeof *t);
  size_t i, cnt;
  /* lib/stable.c:118-129 */
  if (maxM < 10) maxM = 10;
  if (maxN < maxM) maxN = maxM;
  if (M < 10) M = 10;
  if (N < M) N = M;
  if (N > maxN) N = maxN;
  if (M > maxM) M = maxM;
  t->N = N;
  t->M = M;
  t->maxN = maxN;
  t->maxM = maxM;
  t->a = a;
  t->lga = lgamma(1.0 - a);
  t->flags = flags;
  t->ld = M;
  cnt = (size_t)N * t->ld;
  t->S1 = (double *)malloc(sizeof(double) * N);
  orc_fill_S1(N, a, t->S1);
  if (flags & ORC_STABLE) {
    t->S = (double *)calloc(cnt, sizeof(double));
    orc_fill_S(N, M, a, t->S, t->ld);
    if (flags & ORC_FLOAT) /* fresh float table == (float) of the FP64 value, SURVEY.md 8c */
      for (i = 0; i < cnt; i++) t->S[i] = (double)(float)t->S[i];
  }
  if (flags & ORC_UVTABLE) {
    t->V = (double *)calloc(cnt, sizeof(double));
    orc_fill_V(N, M, a, t->V, t->ld);
    if (flags & ORC_FLOAT)
      for (i = 0; i < cnt; i++) t->V[i] = (double)(float)t->V[i];
  }
  return t;
}

void orc_free(orc_table *t) {
  if (!t) return;
  free(t->S);
  free(t->V);
  free(t->S1);
  free(t);
}

/* lib/stable.c:822-873 without the cache growth: rows past N come from the closed form */
double orc_S1(const orc_table *t, unsigned n) {
  if (n == 0) return -HUGE_VAL;
  if (n <= t->N) return t->S1[n - 1];
  if (n > t->maxN && !(t->flags & ORC_ASYMPT)) return -HUGE_VAL;
  return lgamma(n - t->a) - t->lga;
}

/* lib/stable.c:941-974 (fixed extent: what lies inside N x M is read, the rest is "beyond") */
double orc_S(const orc_table *t, unsigned n, unsigned m) {
  if (!(t->flags & ORC_STABLE)) return -HUGE_VAL;
  if (n == m) return 0;
  if (m == 1) return orc_S1(t, n);
  if (n < m || m == 0) return -HUGE_VAL;
  if (m > t->M || n > t->N) {
    if (n > t->maxN && (t->flags & ORC_ASYMPT)) return orc_asympt(t->a, n, m);
    return -HUGE_VAL;
  }
  return t->S[(size_t)(n - 1) * t->ld + (m - 1)];
}

/* lib/stable.c:900-939 */
double orc_V(const orc_table *t, unsigned n, unsigned m) {
  if (!(t->flags & ORC_UVTABLE)) return 0;
  if (n > t->N || m > t->M) {
    if (n > t->maxN && (t->flags & ORC_ASYMPT)) return orc_V_asympt(t->a, n, m);
    return 0;
  }
  if (m < 2 || n < m) return 0;
  return t->V[(size_t)(n - 1) * t->ld + (m - 1)];
}

/* lib/stable.c:875-883 */
double orc_U(const orc_table *t, unsigned n, unsigned m) {
  if (m == 1) return n - t->a;
  return n - m * t->a + 1 / orc_V(t, n, m);
}

/* lib/stable.c:885-897 */
double orc_UV(const orc_table *t, unsigned n, unsigned m) {
  if (m == 1) return -HUGE_VAL;
  if (m == n + 1) return 1;
  if (m == n) return (n + 1.0) / (n - 1.0);
  return (n - m * t->a) * orc_V(t, n, m) + 1.0;
}

uint64_t orc_cells_S(uint64_t N, uint64_t M) { return (M - 1) * (M - 2) / 2 + (N - M) * (M - 1); }
uint64_t orc_cells_V(uint64_t N, uint64_t M) { return M * (M - 1) / 2 + (N - M) * (M - 1); }

/* ------------------------------------------------------------------------------------------ */
/* samplea2's seat-partition sampler, lib/samplea.c:227-341 (SAMPLEA_M builds)                  */
/* ------------------------------------------------------------------------------------------ */
/* lib/samplea.c:233-239 */
double orc_logminus(double x, double y) {
  if (y >= x) return -HUGE_VAL;
  if (y - x < -80) return x - exp(y - x);
  return x + log(1 - exp(y - x));
}

/*
 * One node of lib/samplea.c:290-321: n customers at t tables (1 < t < n) inside table tb; writes
 * the t-1 sizes m[M-1], M = t-1 .. 1.
 *   exact == 0: the reference's statements in its order -- one uniform (logu[0] = log u),
 *     rem = ptot + log u (:294), factor (l - a)(N-l+1)/(l-1) (:303), ptot and rem carried over the
 *     rounds.  Pinned against the reference itself by tests/ref_samplea2_probe.py.
 *   exact != 0: P(l | N, M+1) = C(N-1,l-1) (1-a)_{l-1} S^{N-l}_M / S^N_{M+1}; logu[M-1] = log of the
 *     uniform of round M, rem = that, ptot = S(N, M+1) of the current N, factor (l-1-a)(N-l+1)/(l-1).
 */
void orc_partition_node(const orc_table *tb, double a, unsigned n, unsigned t, const double *logu, uint16_t *m,
                        int exact) {
  int N = (int)n, M;
  double ptot = orc_S(tb, n, t);
  double rem = exact ? 0.0 : ptot + logu[0];
  for (M = (int)t - 1; M >= 1; M--) {
    int l;
    double fact = 0.0;
    if (exact) {
      ptot = orc_S(tb, N, M + 1);
      rem = logu[M - 1];
    }
    for (l = 1; l <= N - M; l++) {
      double term;
      if (l > 1) fact += log(((exact ? l - 1 : l) - a) * (N - l + 1) / (l - 1));
      term = fact + orc_S(tb, N - l, M) - ptot;
      if (term >= rem) break;
      rem = orc_logminus(rem, term);
    }
    if (l > N - M) l = N - M;
    m[M - 1] = (uint16_t)l;
    N -= l;
  }
}

/* log P(l | N, M+1) of the exact mode, for the tests' normalisation and frequency checks */
double orc_partition_logp(const orc_table *tb, double a, unsigned N, unsigned M, unsigned l) {
  double fact = 0.0;
  unsigned j;
  for (j = 2; j <= l; j++) fact += log(((double)j - 1 - a) * (N - j + 1) / (j - 1));
  return fact + orc_S(tb, N - l, M) - orc_S(tb, N, M + 1);
}

/* ------------------------------------------------------------------------------------------ */
/* table-indicator Gibbs step of the reference's demo, test/demo.c:405-434                     */
/* ------------------------------------------------------------------------------------------ */
/*
 * One or more sweeps over the tokens of R restaurants (restaurant j: tokens tok_dish[tok_off[j] .. tok_off[j+1]),
 * counts n[j*D + i], table counts t[j*D + i], T[j] = sum_i t[j][i]).  Per token of dish i with n > 1: the
 * indicator is removed with probability (t-1)/(n-1) (drawn only when t > 1, :418-422), then added with probability
 * one/(one+1), one = H[i] (b + T a) t / (n - t + 1) V^n_{t+1} as a float (:425-430).  Uniforms come from glibc's
 * 48-bit generator (erand48 = drand48 with an explicit state, lib/srng.h:28-34): shared != 0 is the demo's schedule
 * (one stream, restaurants in order), otherwise restaurant j runs on its own stream rng[j].
 */
static void orc_ti_restaurant(const orc_table *tb, double apar, double bpar, const uint32_t *dish, uint32_t ntok,
                              const float *H, const uint32_t *n, uint16_t *t, uint32_t *Tj, unsigned short xs[3]) {
  uint32_t c;
  for (c = 0; c < ntok; c++) {
    const uint32_t i = dish[c];
    float one;
    if (n[i] == 1) continue;
    if (t[i] > 1 && (n[i] - 1) * erand48(xs) < (t[i] - 1)) {
      t[i]--;
      (*Tj)--;
    }
    one = H[i] * (bpar + *Tj * apar) * (t[i]) / (n[i] - t[i] + 1) * orc_V(tb, n[i], t[i] + 1);
    if (erand48(xs) < one / (one + 1.0)) {
      t[i]++;
      (*Tj)++;
    }
  }
}

static void orc_unpack48(uint64_t x, unsigned short xs[3]) {
  xs[0] = (unsigned short)(x & 0xFFFF);
  xs[1] = (unsigned short)((x >> 16) & 0xFFFF);
  xs[2] = (unsigned short)((x >> 32) & 0xFFFF);
}
static uint64_t orc_pack48(const unsigned short xs[3]) { return (uint64_t)xs[0] | ((uint64_t)xs[1] << 16) | ((uint64_t)xs[2] << 32); }

void orc_ti_gibbs(const orc_table *tb, double apar, double bpar, size_t R, const uint32_t *tok_off, const uint32_t *tok_dish,
                  const float *H, uint32_t D, const uint32_t *n, uint16_t *t, uint32_t *T, uint64_t *rng, int shared,
                  int sweeps) {
  unsigned short xs[3];
  size_t j;
  int s;
  if (shared) {
    orc_unpack48(rng[0], xs);
    for (s = 0; s < sweeps; s++)
      for (j = 0; j < R; j++)
        orc_ti_restaurant(tb, apar, bpar, tok_dish + tok_off[j], tok_off[j + 1] - tok_off[j], H, n + j * D, t + j * D, &T[j], xs);
    rng[0] = orc_pack48(xs);
    return;
  }
  for (j = 0; j < R; j++) {
    orc_unpack48(rng[j], xs);
    for (s = 0; s < sweeps; s++)
      orc_ti_restaurant(tb, apar, bpar, tok_dish + tok_off[j], tok_off[j + 1] - tok_off[j], H, n + j * D, t + j * D, &T[j], xs);
    rng[j] = orc_pack48(xs);
  }
}
