/*
 * psample.c -- the scalar entry points of psample.h / srng.h / digamma.h: thin adapters.
 *
 * The reference's samplers (lib/samplea.c, lib/sampleb.c, lib/sslice.c) each own a sequential loop over
 * a log-posterior callback.  Here there is ONE engine for all of them, the lock-step drivers of
 * psample_batch.c (csrc/psample_core.h), which advance any number of chains, each on a random stream of
 * its own.  A scalar call is the engine with a single chain on the host back end (libm arithmetic in the
 * reference's summation order; a table evaluation is a refill by the CUDA engine), and the chain's stream
 * is glibc's own 48-bit generator: its state is taken over when the call starts and handed back when it
 * returns, so a caller that seeds with srand48() sees the reference's draws and finds the generator where
 * the reference would have left it.  (The adaptive rejection sampler draws from rand(), like the
 * reference's, lib/arms.c:913-918.)
 *
 * Error convention of the reference kept: the samplers do not return failures, they exit(1)
 * (lib/samplea.c:50-53, 212-221; lib/sampleb.c:86-89, 136-153).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arms.h"
#include "digamma.h"
#include "psample.h"
#include "psample_core.h"
#include "rng48.h"
#include "specfun.h"

/* ------------------------------------------------------------------------------------------ */
/* special functions (include/digamma.h): host builds of csrc/specfun.h                         */
/* ------------------------------------------------------------------------------------------ */
double digammaRN(double x) { return stb_digammaRN(x); }
double MLdigamma(double x) { return stb_digamma(x); }
double MLtrigamma(double x) { return stb_trigamma(x); }
double MLtetragamma(double x) { return stb_tetragamma(x); }
double MLpentagamma(double x) { return stb_pentagamma(x); }
double digammaInv(double x) { return stb_digamma_inv(x); }
/* the order is rounded to the nearest integer (lib/polygamma.c:502-523) */
double MLpsigamma(double x, double deriv) {
  if (isnan(x)) return x;
  return stb_polygamma((int)floor(deriv + 0.5), x);
}

/* ------------------------------------------------------------------------------------------ */
/* glibc's 48-bit generator as a chain's stream                                                */
/* ------------------------------------------------------------------------------------------ */
static stb_zig_tables g_zig;
static int g_zig_ready = 0;
const stb_zig_tables *stb_zig_tables_get(void) {
  if (!g_zig_ready) {
    stb_zig_tables_build(&g_zig);
    g_zig_ready = 1;
  }
  return &g_zig;
}

/* seed48() installs a state and returns the one it replaces: that is the only window glibc offers */
static uint64_t glibc48_take(void) {
  unsigned short zero[3] = {0, 0, 0};
  const unsigned short *was = seed48(zero);
  return (uint64_t)was[0] | ((uint64_t)was[1] << 16) | ((uint64_t)was[2] << 32);
}
static void glibc48_give(uint64_t x) {
  unsigned short s[3];
  int w;
  for (w = 0; w < 3; w++) s[w] = (unsigned short)(x >> (16 * w));
  seed48(s);
}

/* the distributions of srng.h on the global stream (lib/gslrandist.c:194-282) */
double gsl_rng_gaussian_ziggurat(const double sigma) {
  uint64_t s = glibc48_take();
  const double v = stb_rng48_gaussian(&s, sigma);
  glibc48_give(s);
  return v;
}
double gsl_rng_gamma(const double a) {
  uint64_t s = glibc48_take();
  const double v = stb_rng48_gamma(&s, a);
  glibc48_give(s);
  return v;
}
double gsl_rng_beta(const double a, const double b) {
  uint64_t s = glibc48_take();
  const double v = stb_rng48_beta(&s, a, b);
  glibc48_give(s);
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* which sampler                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* a compile-time switch in the reference (PSAMPLE_ARS, lib/psample.h:37), a run-time one here; unmodified
 * callers pick the reference's shipped configuration with STB_SAMPLER=ars in the environment */
static int g_mode = -1;
static int mode_now(void) {
  if (g_mode < 0) {
    const char *s = getenv("STB_SAMPLER");
    g_mode = (s && (!strcmp(s, "ars") || !strcmp(s, "ARS") || !strcmp(s, "1"))) ? STB_SAMPLER_ARS : STB_SAMPLER_SLICE;
  }
  return g_mode;
}
int stb_set_sampler(int which) {
  const int was = mode_now();
  g_mode = which == STB_SAMPLER_ARS ? STB_SAMPLER_ARS : STB_SAMPLER_SLICE;
  return was;
}
/* ARS on glibc's rand(): the engine's source with no per-chain stream */
static const stb_ars_source g_glibc_rand = {NULL};
static const stb_ars_source *ars_now(void) { return mode_now() == STB_SAMPLER_ARS ? &g_glibc_rand : NULL; }

static void give_up(const char *who, int rc) {
  fprintf(stderr, "%s: the sampler failed (%d)%s%s\n", who, rc, stb_last_error()[0] ? ": " : "", stb_last_error());
  exit(1);
}

/* ------------------------------------------------------------------------------------------ */
/* SliceSimple (lib/sslice.c:33-80): one chain of the lock-step slice sampler                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double (*post)(double, void *);
  void *pars;
} CallerDensity;

static int caller_density(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  const CallerDensity *d = (const CallerDensity *)ctx;
  size_t j;
  (void)chain;
  for (j = 0; j < cnt; j++) out[j] = d->post(x[j], d->pars);
  return 0;
}

int SliceSimple(double *xp, double (*post)(double, void *), double *bounds, rngp_t rng, int loops, void *pars) {
  CallerDensity d;
  uint64_t stream = glibc48_take();
  int rc;
  (void)rng; /* the stream is glibc's global one, as in the reference (lib/srng.h:28-30) */
  d.post = post;
  d.pars = pars;
  rc = stb_slice_lockstep(xp, 1, &bounds[0], &bounds[1], &stream, loops, caller_density, &d, NULL, 1, 0);
  glibc48_give(stream);
  return rc != 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sampleb, samplea, samplea2                                                                  */
/* ------------------------------------------------------------------------------------------ */
double sampleb(double b_in, int I, double shape, double scale, scnt_int *N, scnt_int *T, double apar, rngp_t rng,
               int loops, int verbose) {
  uint64_t stream = glibc48_take();
  int rc;
  (void)rng;
  rc = stb_concentration_step(&b_in, 1, I, shape, scale, N, T, &apar, &stream, ars_now(), loops, NULL, STB_BACKEND_HOST,
                              verbose);
  glibc48_give(stream);
  if (rc) give_up("sampleb", rc);
  return b_in;
}

/* the statistics as arrays: the engine takes n[i][k], t[i][k]; a caller with a getval callback gets them copied once */
typedef struct {
  scnt_int **n;
  stcnt_int **t;
  scnt_int *nbuf;
  stcnt_int *tbuf;
} CountRows;

static void rows_from_callback(CountRows *r, int I, const int *K, void (*getval)(scnt_int *, stcnt_int *, unsigned, unsigned)) {
  size_t total = 0, at = 0;
  int i, k;
  for (i = 0; i < I; i++) total += (size_t)K[i];
  r->n = (scnt_int **)malloc(sizeof(scnt_int *) * (size_t)(I > 0 ? I : 1));
  r->t = (stcnt_int **)malloc(sizeof(stcnt_int *) * (size_t)(I > 0 ? I : 1));
  r->nbuf = (scnt_int *)malloc(sizeof(scnt_int) * (total ? total : 1));
  r->tbuf = (stcnt_int *)malloc(sizeof(stcnt_int) * (total ? total : 1));
  if (!r->n || !r->t || !r->nbuf || !r->tbuf) {
    fprintf(stderr, "samplea: out of memory\n");
    exit(1);
  }
  for (i = 0; i < I; i++) {
    r->n[i] = r->nbuf + at;
    r->t[i] = r->tbuf + at;
    for (k = 0; k < K[i]; k++, at++) getval(&r->nbuf[at], &r->tbuf[at], (unsigned)i, (unsigned)k);
  }
}
static void rows_free(CountRows *r) {
  free(r->n);
  free(r->t);
  free(r->nbuf);
  free(r->tbuf);
}

double samplea(double apar, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
               void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng, int loops,
               int verbose) {
  CountRows own = {NULL, NULL, NULL, NULL};
  uint64_t stream;
  int rc;
  (void)rng;
  if (getval) {
    rows_from_callback(&own, I, K, getval);
    n = own.n;
    t = own.t;
  }
  stream = glibc48_take();
  rc = stb_discount_step(&apar, 1, I, K, T, n, t, bpar, 0, &stream, ars_now(), loops, NULL, STB_BACKEND_HOST, verbose);
  glibc48_give(stream);
  rows_free(&own);
  if (rc) give_up("samplea", rc);
  return apar;
}

double logminus(double x, double y) { /* log(e^x - e^y), lib/samplea.c:229-239 */
  const double d = y - x;
  if (!(d < 0)) return -HUGE_VAL;
  return d < -80 ? x - exp(d) : x + log(1 - exp(d));
}

static int g_partition_mode = STB_PARTITION_REFERENCE;
int stb_get_partition_mode(void) { return g_partition_mode; }
int stb_set_partition_mode(int mode) {
  const int was = g_partition_mode;
  if (mode == STB_PARTITION_REFERENCE || mode == STB_PARTITION_EXACT) g_partition_mode = mode;
  return was;
}

/* getval is not used by the partition step, like the reference (lib/samplea.c:290-321 reads the arrays) */
double samplea2(double mya, stable_t *S, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
                void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng, int loops,
                int verbose) {
  uint64_t stream = glibc48_take();
  int rc;
  (void)rng;
  (void)getval;
  rc = stb_discount_step_partition(&mya, S, I, K, T, n, t, bpar, &stream, ars_now(), loops,
                                   g_partition_mode == STB_PARTITION_EXACT, verbose);
  glibc48_give(stream);
  if (rc) give_up("samplea2", rc);
  return mya;
}
