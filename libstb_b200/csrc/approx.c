/*
 * approx.c -- the scalar closed forms beside the tables (C host code, no table, no device):
 *   S_approx / S_approx_da            include/sapprox.h   (reference: lib/sapprox.c:28-114)
 *   gammadiff / psidiff / *cache_*    include/lgamma.h    (reference: lib/lgamma.c:30-240)
 *
 * These are O(1) formulas the reference evaluates on the host next to a table look-up (SURVEY.md
 * section 8, row a10); they stay host C here for the same reason S_asympt does: one lgamma per
 * call, nothing to batch.  The arithmetic TYPES of the reference are part of its results and are
 * kept: S_approx takes the discount as a float and forms  n - j*a  and  1 - j*a  in single
 * precision (lib/sapprox.c:38-67: `n-2*a` with int n, float a), so its values carry ~1e-7
 * relative error by construction; a replacement that "fixed" that would not be a drop-in.
 */
#include <math.h>
#include <pthread.h>
#include <string.h>

#include "digamma.h"
#include "lgamma.h"
#include "sapprox.h"

/* lgamma(n - j a) - lgamma(1 - j a), arguments formed in float like the reference does */
static double lg_ratio(int n, int j, float a) {
  const float ja = (float)j * a;
  const float top = (float)n - ja, bot = 1.0f - ja;
  return lgamma(top) - lgamma(bot);
}
/* digamma(n - j a) - digamma(1 - j a), same argument formation */
static double psi_ratio(int n, int j, float a) {
  const float ja = (float)j * a;
  const float top = (float)n - ja, bot = 1.0f - ja;
  return digamma(top) - digamma(bot);
}

/*
 * log S^n_m for m <= 4 from the alternating sum
 *   S^n_m = 1/(a^(m-1) (m-1)!) sum_j (-1)^(j-1) C(m-1, j-1) Gamma(n-ja)/Gamma(1-ja)
 * (lib/sapprox.c:52-73), factored around its last term; for a < 0.001 the a -> 0 limit in
 * polygamma differences (lib/sapprox.c:42-56).  The reference asserts a > 0 (compiled out with
 * -DNDEBUG in its own Makefile); a <= 0 simply falls through the same arithmetic here.
 */
double S_approx(int n, int m, float a) {
  if (n == m) return 0.0;
  if (n < m) return -HUGE_VAL;
  if (m == 1) return lg_ratio(n, 1, a);
  if (a < 0.001) {
    const float top = (float)n - a, bot = 1.0f - a;
    const double lead = lgamma(top) - lgamma(bot);
    if (m == 2) return lead + log(digamma(top) - digamma(bot));
    if (m == 3) {
      const double d1 = digamma(top) - digamma(bot);
      return lead - log(2.0) + log(trigamma(top) - trigamma(bot) + d1 * d1);
    }
    if (m == 4) {
      const double d1 = digamma(top) - digamma(bot);
      return lead - log(6.0) +
             log((tetragamma(top) - tetragamma(bot)) + 3 * (trigamma(top) - trigamma(bot)) * d1 + d1 * d1 * d1);
    }
    return -HUGE_VAL;
  }
  if (m == 2) {
    const double g1 = lg_ratio(n, 1, a), g2 = lg_ratio(n, 2, a);
    return g2 - log(a) + log(exp(g1 - g2) - 1.0);
  }
  if (m == 3) {
    const double g1 = lg_ratio(n, 1, a), g2 = lg_ratio(n, 2, a), g3 = lg_ratio(n, 3, a);
    return g3 - 2 * log(a) - log(2.0) + log(exp(g1 - g3) - 2 * exp(g2 - g3) + 1.0);
  }
  if (m == 4) {
    const double g1 = lg_ratio(n, 1, a), g2 = lg_ratio(n, 2, a), g3 = lg_ratio(n, 3, a), g4 = lg_ratio(n, 4, a);
    return g4 - 3 * log(a) - log(6.0) + log(exp(g1 - g4) - 3 * exp(g2 - g4) + 3 * exp(g3 - g4) - 1.0);
  }
  return -HUGE_VAL;
}

/*
 * d/da log S^n_m for m <= 4 (lib/sapprox.c:82-114): each term of the sum differentiates to
 * -j (psi(n-ja) - psi(1-ja)) times itself; the terms are weighted by exp(g_j - log S).
 * `2/a` and `3/a` are single-precision divisions in the reference (int / float) and stay so.
 */
double S_approx_da(int n, int m, float a) {
  if (n == m) return 0.0;
  if (n < m) return -HUGE_VAL;
  if (m == 1) return -psi_ratio(n, 1, a);
  const double s = S_approx(n, m, a);
  if (m == 2) {
    const double w1 = exp(lg_ratio(n, 1, a) - s), w2 = exp(lg_ratio(n, 2, a) - s);
    const double d1 = -psi_ratio(n, 1, a), d2 = -2.0 * psi_ratio(n, 2, a);
    return (w1 * d1 - w2 * d2 - 1) / a;
  }
  if (m == 3) {
    const double w1 = exp(lg_ratio(n, 1, a) - s), w2 = exp(lg_ratio(n, 2, a) - s), w3 = exp(lg_ratio(n, 3, a) - s);
    const double d1 = -psi_ratio(n, 1, a), d2 = -2 * psi_ratio(n, 2, a), d3 = -3 * psi_ratio(n, 3, a);
    const float lead = -2 / a;
    return lead + (w1 * d1 - 2 * w2 * d2 + w3 * d3) / 2 / a / a;
  }
  if (m == 4) {
    const double w1 = exp(lg_ratio(n, 1, a) - s), w2 = exp(lg_ratio(n, 2, a) - s), w3 = exp(lg_ratio(n, 3, a) - s),
                 w4 = exp(lg_ratio(n, 4, a) - s);
    const double d1 = -psi_ratio(n, 1, a), d2 = -2 * psi_ratio(n, 2, a), d3 = -3 * psi_ratio(n, 3, a),
                 d4 = -4 * psi_ratio(n, 4, a);
    const float lead = -3 / a;
    return lead + (w1 * d1 - 3 * w2 * d2 + 3 * w3 * d3 - w4 * d4) / 3 / a / a / a;
  }
  return -HUGE_VAL;
}

/* ---- self-filling caches of differences (lib/lgamma.c:30-118) ---------------------------------- */
/* A cache slot holding 0 means "not computed yet" (the reference's convention, lib/lgamma.c:33,
 * 42: a difference that is exactly 0 is simply recomputed every time). */
static void cache_reset(struct gcache_s *c, double p, double base) {
  c->par = p;
  c->lgpar = base;
  memset(c->cache, 0, sizeof c->cache);
}

void gcache_init(struct gcache_s *c, double p) { cache_reset(c, p, lgamma(p)); }

double gcache_value(struct gcache_s *c, int j) {
  if (j <= 0) return 0;
  if (j >= GCACHE) return lgamma(j + c->par) - c->lgpar;
  if (c->cache[j] == 0) {
    const double p = c->par;
    /* rising factorials for the first three, lib/lgamma.c:44-49 */
    c->cache[j] = j == 1   ? log(p)
                  : j == 2 ? log(p * (p + 1))
                  : j == 3 ? log(p * (p + 1) * (p + 2))
                           : lgamma(j + p) - c->lgpar;
  }
  return c->cache[j];
}

void pcache_init(struct gcache_s *c, double p) { cache_reset(c, p, digamma(p)); }

double pcache_value(struct gcache_s *c, int j) {
  if (j <= 0) return 0;
  if (j >= GCACHE) return digamma(j + c->par) - c->lgpar;
  if (c->cache[j] == 0) {
    const double p = c->par;
    c->cache[j] = j == 1   ? 1 / p
                  : j == 2 ? 1 / p + 1 / (1 + p)
                  : j == 3 ? 1 / p + 1 / (1 + p) + 1 / (2 + p)
                           : digamma(j + p) - c->lgpar;
  }
  return c->cache[j];
}

/* S^{n+1}_{2,a} / S^n_{1,a}  (lib/lgamma.c:85-90): digamma difference below a = 0.02 */
static double ratio_s2_s1(double a, int n, double lg0) {
  if (a < 0.02) return digamma(n + 1 - a) - digamma(1 - a);
  return (1.0 - exp(lgamma(n + 1 - 2 * a) - lgamma(n + 1 - a) - lg0)) / a;
}

void qcache_init(struct gcache_s *c, double p) {
  cache_reset(c, p, p > 0 ? lgamma(1 - 2 * p) - lgamma(1 - p) : 0);
}

double qcache_value(struct gcache_s *c, int j) {
  if (j <= 0) return 0;
  if (j >= GCACHE) return ratio_s2_s1(c->par, j, c->lgpar);
  if (c->cache[j] == 0) {
    const double p = c->par;
    c->cache[j] = j == 1   ? 1 / (1 - p)
                  : j == 2 ? 3 / (2 - p)
                  : j == 3 ? (11 - 7 * p) / (3 - p) / (2 - p)
                           : ratio_s2_s1(p, j, c->lgpar);
  }
  return c->cache[j];
}

/* ---- gammadiff / psidiff (lib/lgamma.c:123-240, polygamma configuration) ------------------------ */
/*
 * For alpha <= 1/2 the reference expands lgamma(N+alpha) (digamma(N+alpha)) to third order in alpha
 * about the INTEGER N, with lgamma/digamma/trigamma/tetragamma/pentagamma tabulated at 3..1999
 * (lib/lgamma.c:124-141).  That truncation is part of its results (~alpha^4 psi'''(N)/24), so
 * it is kept; the table is filled once, under pthread_once (the reference's unguarded flag is not
 * thread-safe).
 */
#define FD_N 2000
static double fd_lg[FD_N], fd_p0[FD_N], fd_p1[FD_N], fd_p2[FD_N], fd_p3[FD_N];
static pthread_once_t fd_once = PTHREAD_ONCE_INIT;
static void fd_fill(void) {
  for (int i = 3; i < FD_N; i++) {
    fd_lg[i] = lgamma(i);
    fd_p0[i] = digamma(i);
    fd_p1[i] = trigamma(i);
    fd_p2[i] = tetragamma(i);
    fd_p3[i] = pentagamma(i);
  }
}
/* third-order expansions about the integer i */
static double lg_taylor(int i, double al) { return fd_lg[i] + al * (fd_p0[i] + al / 2 * (fd_p1[i] + al / 3 * fd_p2[i])); }
static double psi_taylor(int i, double al) { return fd_p0[i] + al * (fd_p1[i] + al / 2 * (fd_p2[i] + al / 3 * fd_p3[i])); }

double gammadiff(int N, double alpha, double lga) {
  pthread_once(&fd_once, fd_fill);
  if (N <= 0) return 0;  /* (the reference returns log(alpha) for N < 0; only N >= 0 is meaningful) */
  if (N == 1) return log(alpha);
  if (N == 2) return log(alpha * (1 + alpha));
  if (N == 3) return log(alpha * (1 + alpha) * (2 + alpha));
  if (alpha > 0.5) return lgamma(N + alpha) - lgamma(alpha);
  if (lga != 0) return (N >= FD_N ? lgamma(N + alpha) : lg_taylor(N, alpha)) - lga;
  /* lgamma(alpha) unknown: go through lgamma(3+alpha) - lgamma(alpha) = log(alpha(1+alpha)(2+alpha)) */
  const double first = log(alpha * (1 + alpha) * (2 + alpha));
  if (N >= FD_N) return first + (lgamma(N + alpha) - lg_taylor(3, alpha));
  return first + ((fd_lg[N] - fd_lg[3]) +
                  alpha * ((fd_p0[N] - fd_p0[3]) + alpha / 2 * ((fd_p1[N] - fd_p1[3]) + alpha / 3 * (fd_p2[N] - fd_p2[3]))));
}

double psidiff(int N, double alpha, double pa) {
  pthread_once(&fd_once, fd_fill);
  if (N <= 0) return 0;
  if (N == 1) return 1 / alpha;
  if (N == 2) return 1 / alpha + 1 / (1 + alpha);
  if (N == 3) return 1 / alpha + 1 / (1 + alpha) + 1 / (2 + alpha);
  if (alpha > 0.5) return digamma(N + alpha) - (pa > 0 ? pa : digamma(alpha));
  if (pa != 0) return (N >= FD_N ? digamma(N + alpha) : psi_taylor(N, alpha)) - pa;
  const double first = 1 / alpha + 1 / (1 + alpha) + 1 / (2 + alpha);
  /* DEVIATION: for N >= 2000 the reference adds lgamma(N+alpha) here (lib/lgamma.c:224-227), an
   * obvious slip for digamma(N+alpha); the correct function is used */
  if (N >= FD_N) return first + (digamma(N + alpha) - psi_taylor(3, alpha));
  return first + ((fd_p0[N] - fd_p0[3]) + alpha * (fd_p1[N] - fd_p1[3]) + alpha * alpha / 2 * (fd_p2[N] - fd_p2[3]) +
                  alpha * alpha * alpha / 6 * (fd_p3[N] - fd_p3[3]));
}
