/*
 * psample_batch.c -- batched samplers: C independent chains advance in lock-step, every
 * log-posterior evaluation of a round is one batched device evaluation.
 *
 * Each chain replays exactly the control flow of the scalar sampler (SliceSimple,
 * lib/sslice.c:33-80; samplea lib/samplea.c:155-225; sampleb lib/sampleb.c:79-159) on its OWN
 * 48-bit random stream, so chain c started from stb_rng48_state(seed) makes the draws the
 * reference makes after srand48(seed) -- up to the tiny differences of the evaluated densities
 * (device lgamma / tree sums vs libm / sequential sums), which only matter if they flip an
 * accept/reject comparison.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ars_machine.h"
#include "psample.h"
#include "rand31.h"
#include "rng48.h"
#include "specfun.h"
#include "stb_b200.h"
#include "stb_cuda.h"

uint64_t stb_rng48_state(long seed) {
  stb_rng48 r;
  stb_rng48_seed(&r, seed);
  return r.x;
}
double stb_rng48_drand(uint64_t *state) {
  stb_rng48 r;
  double v;
  r.x = *state;
  v = stb_rng48_unit(&r);
  *state = r.x;
  return v;
}
long stb_rng48_lrand48(uint64_t *state) {
  stb_rng48 r;
  long v;
  r.x = *state;
  v = stb_rng48_lrand(&r);
  *state = r.x;
  return v;
}
extern const stb_zig_tables *stb_zig_tables_get(void);
double stb_rng48_gaussian(uint64_t *state, double sigma) {
  stb_rng48 r;
  double v;
  r.x = *state;
  v = stb_gauss_zig(&r, stb_zig_tables_get(), sigma);
  *state = r.x;
  return v;
}
double stb_rng48_gamma(uint64_t *state, double a) {
  stb_rng48 r;
  double v;
  r.x = *state;
  v = stb_gamma(&r, stb_zig_tables_get(), a);
  *state = r.x;
  return v;
}
double stb_rng48_beta(uint64_t *state, double a, double b) {
  stb_rng48 r;
  double v;
  r.x = *state;
  v = stb_beta(&r, stb_zig_tables_get(), a, b);
  *state = r.x;
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* lock-step slice sampler                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef int (*eval_fn)(void *ctx, const double *x, const int *chain, size_t cnt, double *out);

#define TOOMANY 200
enum { PH_NEED_Y = 0, PH_TRY = 1, PH_DONE = 2 };

static void trace_add(stb_sample_stats *st, size_t c, double x, double v) {
  if (!st || !st->trace_x || !st->trace_n) return;
  if (st->trace_n[c] < st->trace_cap) {
    st->trace_x[c * st->trace_cap + st->trace_n[c]] = x;
    if (st->trace_v) st->trace_v[c * st->trace_cap + st->trace_n[c]] = v;
  }
  st->trace_n[c]++;
}

/*
 * xp[c]: start and result; lo[c], hi[c]: bounds.  Returns 0, or 1 + index of the first failing chain.
 *
 * depth > 1: speculative proposals.  A slice sampler's proposals do not depend on the density: the
 * next one is drawn uniformly from the bracket, and a rejected proposal shrinks the bracket towards
 * the current point -- the density only decides where the sequence STOPS.  So a round may evaluate
 * the next `depth` proposals of every chain at once (made from a COPY of the chain's stream) and
 * then replay the sampler over the values: the first accepted proposal ends the replay, the stream
 * advances by exactly the uniforms the sequential sampler would have drawn, and the proposals behind
 * an acceptance are dropped.  Same draws, same stream, fewer device round trips; worth it when an
 * evaluation is cheap (sampleb: a 1000-term reduction).  When it is a table fill (samplea), the device
 * works in waves of `slots` evaluations that cost the same full or not: slots > 0 hands only the
 * slots a round would leave empty in its last wave to speculative proposals (they cost nothing), so
 * the long tail of rounds with a few chains left collapses.
 */
static int slice_lockstep(double *xp, size_t C, const double *lo, const double *hi, uint64_t *rng, int loops,
                          eval_fn eval, void *ctx, stb_sample_stats *st, int depth, size_t slots) {
  const size_t W = (size_t)(depth < 1 ? 1 : depth) + 1; /* evaluations per chain and round, at most */
  int *phase = (int *)malloc(sizeof(int) * C), *left = (int *)malloc(sizeof(int) * C);
  int *tries = (int *)malloc(sizeof(int) * C), *chain = (int *)malloc(sizeof(int) * C * W);
  int *nprop = (int *)malloc(sizeof(int) * C);
  size_t *first = (size_t *)malloc(sizeof(size_t) * C);
  double *y = (double *)malloc(sizeof(double) * C), *r0 = (double *)malloc(sizeof(double) * C);
  double *r1 = (double *)malloc(sizeof(double) * C), *xq = (double *)malloc(sizeof(double) * C * W);
  double *val = (double *)malloc(sizeof(double) * C * W);
  size_t c, cnt;
  int rc = 0;
  if (depth < 1) depth = 1;
  if (!phase || !left || !tries || !chain || !nprop || !first || !y || !r0 || !r1 || !xq || !val) {
    rc = -1;
    goto done;
  }
  for (c = 0; c < C; c++) {
    if (xp[c] < lo[c] || xp[c] > hi[c]) {
      fprintf(stderr, "SliceSimple: input value %lf outside bounds [%lg,%lg] (chain %zu)\n", xp[c], lo[c], hi[c], c);
      rc = 1 + (int)c;
      goto done;
    }
    phase[c] = loops > 0 ? PH_NEED_Y : PH_DONE;
    left[c] = loops;
    tries[c] = 0;
  }
  for (;;) {
    /* the points this round evaluates */
    size_t active = 0, seen = 0, base = 0, rem = 0;
    if (slots) { /* the evaluations that fit into this round's waves beyond one per chain */
      for (c = 0; c < C; c++) active += phase[c] != PH_DONE;
      if (active) {
        const size_t extra = (active + slots - 1) / slots * slots - active;
        base = extra / active;
        rem = extra % active;
      }
    }
    cnt = 0;
    for (c = 0; c < C; c++) {
      stb_rng48 r;
      double b0, b1;
      int np, k, tr, d = depth;
      if (phase[c] == PH_DONE) continue;
      if (slots) {
        const size_t mine = 1 + base + (seen < rem ? 1 : 0);
        d = mine < (size_t)depth ? (int)mine : depth;
        seen++;
      }
      first[c] = cnt;
      r.x = rng[c]; /* a copy: the stream itself advances in the replay below */
      if (phase[c] == PH_NEED_Y) {
        xq[cnt] = xp[c];
        chain[cnt++] = (int)c;
        (void)stb_rng48_unit(&r); /* the uniform of the slice level comes first */
        b0 = lo[c];
        b1 = hi[c];
        tr = 1;
        np = d - 1;
      } else {
        b0 = r0[c];
        b1 = r1[c];
        tr = tries[c];
        np = d;
      }
      if (np > TOOMANY - tr) np = TOOMANY - tr; /* the sequential sampler gives up there */
      for (k = 0; k < np; k++) {
        const double x = b0 + stb_rng48_unit(&r) * (b1 - b0);
        xq[cnt] = x;
        chain[cnt++] = (int)c;
        if (x < xp[c])
          b0 = x;
        else
          b1 = x;
      }
      nprop[c] = np;
    }
    if (!cnt) break;
    if (eval(ctx, xq, chain, cnt, val)) {
      rc = -2;
      goto done;
    }
    if (st) {
      st->evals += cnt;
      st->rounds++;
    }
    /* replay the sequential sampler over the values */
    for (c = 0; c < C; c++) {
      stb_rng48 r;
      size_t j;
      int k;
      if (phase[c] == PH_DONE) continue;
      j = first[c];
      r.x = rng[c];
      if (phase[c] == PH_NEED_Y) {
        trace_add(st, c, xq[j], val[j]);
        y[c] = val[j] + log(stb_rng48_unit(&r));
        r0[c] = lo[c];
        r1[c] = hi[c];
        tries[c] = 1;
        phase[c] = PH_TRY;
        j++;
      }
      for (k = 0; k < nprop[c]; k++, j++) {
        const double x = r0[c] + stb_rng48_unit(&r) * (r1[c] - r0[c]); /* == xq[j], bit for bit */
        trace_add(st, c, x, val[j]);
        if (val[j] > y[c]) {
          xp[c] = x;
          left[c]--;
          phase[c] = left[c] > 0 ? PH_NEED_Y : PH_DONE;
          break;
        }
        if (x < xp[c])
          r0[c] = x;
        else
          r1[c] = x;
        if (++tries[c] >= TOOMANY) {
          fprintf(stderr, "SliceSimple: giving up after %d tries, range=[%lg,%lg] (chain %zu)\n", TOOMANY, r0[c],
                  r1[c], c);
          rc = 1 + (int)c;
          rng[c] = r.x;
          goto done;
        }
      }
      rng[c] = r.x;
    }
  }
done:
  free(phase);
  free(left);
  free(tries);
  free(chain);
  free(nprop);
  free(first);
  free(y);
  free(r0);
  free(r1);
  free(xq);
  free(val);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* lock-step adaptive rejection sampler                                                        */
/* ------------------------------------------------------------------------------------------ */
void stb_rand31_seed(stb_rand31_t *g, unsigned seed) { stb_rand31_init(g, seed); }
int stb_rand31_next(stb_rand31_t *g) { return stb_rand31_step(g); }

/*
 * One arms_simple(3, lo, hi, ...) per chain (envelope of at most 100 knots, no Metropolis step:
 * lib/arms.c:98-123 as samplea / sampleb call it), every chain a resumable machine on its own
 * rand() stream.  Each round evaluates the one point every unfinished chain is waiting for.
 * xp[c] receives the draw.  A chain whose sampler stops with an arms() error code -- 2000: the
 * log-posterior is not concave on its interval, 2001: a hundred proposals -- keeps the value it
 * came with: that is what the reference does (samplea / sampleb ignore arms_simple's return value,
 * lib/samplea.c:210-211, lib/sampleb.c:135, and the sample variable still holds its input); the
 * number of such chains is returned through *nfailed.  Returns 0, or -1/-2 (memory / evaluation
 * failure).
 */
static int ars_lockstep(double *xp, size_t C, const double *lo, const double *hi, stb_rand31_t *rnd, eval_fn eval,
                        void *ctx, stb_sample_stats *st, size_t *nfailed) {
  stb_ars_t **m = (stb_ars_t **)calloc(C, sizeof *m);
  int *state = (int *)malloc(sizeof(int) * C), *chain = (int *)malloc(sizeof(int) * C);
  double *want = (double *)malloc(sizeof(double) * C), *xq = (double *)malloc(sizeof(double) * C);
  double *val = (double *)malloc(sizeof(double) * C);
  size_t c, cnt;
  int rc = 0;
  if (!m || !state || !chain || !want || !xq || !val) {
    rc = -1;
    goto done;
  }
  for (c = 0; c < C; c++) {
    double xinit[3];
    int i;
    m[c] = stb_ars_new(100);
    if (!m[c]) {
      rc = -1;
      goto done;
    }
    for (i = 0; i < 3; i++) xinit[i] = lo[c] + (i + 1.0) * (hi[c] - lo[c]) / (3 + 1.0);
    state[c] = stb_ars_begin(m[c], xinit, 3, lo[c], hi[c], 1.0, 0, 0.0, &xp[c], 1, stb_rand31_unit, &rnd[c], &want[c]);
  }
  for (;;) {
    cnt = 0;
    for (c = 0; c < C; c++) {
      if (state[c] == STB_ARS_NEED) {
        xq[cnt] = want[c];
        chain[cnt++] = (int)c;
      }
    }
    if (!cnt) break;
    if (eval(ctx, xq, chain, cnt, val)) {
      rc = -2;
      goto done;
    }
    if (st) {
      st->evals += cnt;
      st->rounds++;
    }
    for (size_t j = 0; j < cnt; j++) {
      c = (size_t)chain[j];
      trace_add(st, c, xq[j], val[j]);
      state[c] = stb_ars_feed(m[c], val[j], &want[c]);
    }
  }
  if (nfailed) {
    *nfailed = 0;
    for (c = 0; c < C; c++)
      if (state[c] != STB_ARS_DONE) ++*nfailed;
  }
done:
  if (m)
    for (c = 0; c < C; c++) stb_ars_free(m[c]);
  free(m);
  free(state);
  free(chain);
  free(want);
  free(xq);
  free(val);
  return rc;
}

/* arms_simple(3, ...) for C independent chains with a HOST log-density: the lock-step driver with
 * the evaluations done one by one (include/stb_b200.h).  Chains whose sampler fails keep x[c]. */
typedef struct {
  double (*f)(double, void *);
  void *data;
} HostDensity;

static int host_density_eval(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  HostDensity *h = (HostDensity *)ctx;
  size_t j;
  (void)chain;
  for (j = 0; j < cnt; j++) out[j] = h->f(x[j], h->data);
  return 0;
}

int stb_arms_simple_batch(double *x, size_t C, const double *lo, const double *hi, stb_rand31_t *rnd,
                          double (*myfunc)(double x, void *mydata), void *mydata, size_t *nfailed) {
  HostDensity h;
  h.f = myfunc;
  h.data = mydata;
  if (!C) return 0;
  return ars_lockstep(x, C, lo, hi, rnd, host_density_eval, &h, NULL, nfailed);
}

/* ------------------------------------------------------------------------------------------ */
/* samplea, batched                                                                            */
/* ------------------------------------------------------------------------------------------ */
#define A_SPECULATE 8 /* at most; only into table slots a round would leave empty (see slice_lockstep) */
typedef struct {
  stb_sweep_t *sweep;
  stb_pstat_dev_t *ps;
  int bpar_per_chain;
  double *ssum, *lg;
  stb_sample_stats *st;
} ABatch;

static int aterms_batch(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  ABatch *ab = (ABatch *)ctx;
  float ms = 0.f;
  size_t j;
  for (j = 0; j < cnt; j++)
    if (x[j] <= 0) {
      fprintf(stderr, "Illegal discount value in aterms()\n");
      return 1;
    }
  if (stb_sweep_run(ab->sweep, x, cnt, NULL, ab->ssum, NULL)) return 1;
  if (ab->st) ab->st->eval_ms += stb_sweep_last_fill_ms(ab->sweep);
  if (stb_cuda_pstat_aterms_lg(ab->ps, x, chain, cnt, ab->bpar_per_chain, ab->lg, &ms)) return 1;
  if (ab->st) ab->st->eval_ms += ms;
  for (j = 0; j < cnt; j++) out[j] = ab->lg[j] + ab->ssum[j];
  return 0;
}

/*
 * The sweep handle (resident slabs for one launch's worth of tables: ~1 GB for the config-4 shape)
 * is kept between calls -- an MCMC run calls samplea once per sweep with the same table extent,
 * and allocating / freeing a gigabyte of device memory per call costs more than the evaluations.
 * stb_release_caches() frees it.  (The samplers are not thread-safe in the reference either.)
 */
static struct {
  stb_sweep_t *w;
  unsigned N, M;
} g_sweep_cache;

static void sweep_cache_drop(void) {
  if (g_sweep_cache.w) stb_sweep_free(g_sweep_cache.w);
  g_sweep_cache.w = NULL;
}

void stb_release_caches(void) {
  sweep_cache_drop();
  stb_cuda_pstat_purge();
}

static stb_sweep_t *sweep_acquire(unsigned N, unsigned M) {
  stb_sweep_t *w = g_sweep_cache.w;
  if (w && g_sweep_cache.N == N && g_sweep_cache.M == M) {
    g_sweep_cache.w = NULL;
    return w;
  }
  sweep_cache_drop();
  return stb_sweep_create(N, M, 0);
}

static void sweep_release(stb_sweep_t *w, unsigned N, unsigned M) {
  sweep_cache_drop();
  g_sweep_cache.w = w;
  g_sweep_cache.N = N;
  g_sweep_cache.M = M;
}

static int samplea_batch_core(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n,
                              stcnt_int **t, const double *bpar, int bpar_per_chain, uint64_t *rng,
                              stb_rand31_t *rnd, int loops, stb_sample_stats *st) {
  double *lo = NULL, *hi = NULL;
  uint32_t *nn = NULL, *tt = NULL;
  size_t total = 0, cnt = 0, c, slots = 0;
  int i, k, maxn = 1, maxt = 1, rc = -1;
  unsigned Mx = 0, Nx = 0;
  ABatch ab;
  memset(&ab, 0, sizeof ab);
  if (!C) return 0;
  for (i = 0; i < I; i++) total += (size_t)K[i];
  lo = (double *)malloc(sizeof(double) * C);
  hi = (double *)malloc(sizeof(double) * C);
  nn = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  tt = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  if (!lo || !hi || !nn || !tt) goto done;
  /* bounds per chain, lib/samplea.c:161-177 and :217 */
  for (c = 0; c < C; c++) {
    double mid = a[c];
    if (fabs(mid - A_MAX) / A_MAX < 0.00001) mid = A_MAX * 0.999 + A_MIN * 0.001;
    if (fabs(mid - A_MIN) / A_MIN < 0.00001) mid = A_MIN * 0.999 + A_MAX * 0.001;
    lo[c] = (mid - SQUEEZEA > A_MIN) ? mid - SQUEEZEA : A_MIN;
    /* the slice sampler may move up to A_MAX (lib/samplea.c:217); ARS stays inside the squeeze (:176-177) */
    hi[c] = (rnd && mid + SQUEEZEA < A_MAX) ? mid + SQUEEZEA : A_MAX;
  }
  /* the statistics with n > 1 in (i,k) order; table extent as lib/samplea.c:186-208 */
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++) {
      if ((int)t[i][k] >= maxt) maxt = t[i][k] + 1;
      if ((int)n[i][k] >= maxn) maxn = n[i][k] + 1;
      if (n[i][k] > 1) {
        nn[cnt] = n[i][k];
        tt[cnt] = t[i][k];
        cnt++;
      }
    }
  {
    /* the same clamps S_make applies (lib/stable.c:118-129) */
    Mx = maxt < 10 ? 10u : (unsigned)maxt;
    Nx = (unsigned)maxn < Mx ? Mx : (unsigned)maxn;
    ab.sweep = sweep_acquire(Nx, Mx);
  }
  if (!ab.sweep || stb_sweep_set_pairs(ab.sweep, nn, tt, cnt)) goto done;
  /* a round evaluates at most one point per chain plus what its last wave of tables has room for */
  slots = (size_t)stb_sweep_tables_in_flight(ab.sweep);
  ab.ssum = (double *)malloc(sizeof(double) * (C + slots));
  ab.lg = (double *)malloc(sizeof(double) * (C + slots));
  if (!ab.ssum || !ab.lg) goto done;
  ab.ps = stb_cuda_pstat_create(I, T, NULL, bpar, bpar_per_chain ? C * (size_t)I : (size_t)I, C + slots);
  if (!ab.ps) goto done;
  ab.bpar_per_chain = bpar_per_chain;
  ab.st = st;
  if (rnd) {
    rc = ars_lockstep(a, C, lo, hi, rnd, aterms_batch, &ab, st, NULL);
    for (c = 0; rc == 0 && c < C; c++)
      if (a[c] < lo[c] || a[c] > hi[c]) {
        fprintf(stderr, "Arms_simple(apar) returned value out of bounds (chain %zu)\n", c);
        rc = 1 + (int)c;
      }
  } else
    rc = slice_lockstep(a, C, lo, hi, rng, loops, aterms_batch, &ab, st, A_SPECULATE, slots);
done:
  if (ab.sweep) sweep_release(ab.sweep, Nx, Mx);
  if (ab.ps) stb_cuda_pstat_destroy(ab.ps);
  free(lo);
  free(hi);
  free(nn);
  free(tt);
  free(ab.ssum);
  free(ab.lg);
  return rc;
}

int stb_samplea_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                      const double *bpar, int bpar_per_chain, uint64_t *rng, int loops, stb_sample_stats *st) {
  return samplea_batch_core(a, C, I, K, T, n, t, bpar, bpar_per_chain, rng, NULL, loops, st);
}

int stb_samplea_batch_ars(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                          const double *bpar, int bpar_per_chain, stb_rand31_t *rnd, stb_sample_stats *st) {
  if (!rnd) return -1;
  return samplea_batch_core(a, C, I, K, T, n, t, bpar, bpar_per_chain, NULL, rnd, 1, st);
}

/* ------------------------------------------------------------------------------------------ */
/* sampleb, batched                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  stb_pstat_dev_t *ps;
  const double *Q, *apar;
  double shape;
  double *qv, *av;
  stb_sample_stats *st;
} BBatch;

static int bterms_batch(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  BBatch *bb = (BBatch *)ctx;
  float ms = 0.f;
  size_t j;
  for (j = 0; j < cnt; j++) {
    bb->qv[j] = bb->Q[chain[j]];
    bb->av[j] = bb->apar[chain[j]];
  }
  if (stb_cuda_pstat_bterms(bb->ps, x, bb->qv, bb->av, bb->shape, cnt, 0, out, &ms)) return 1;
  if (bb->st) bb->st->eval_ms += ms;
  return 0;
}

#define B_ERROR 1.0e-4
#define B_LOOPS 5
#define B_SPECULATE 4 /* slice proposals evaluated per chain and round (see slice_lockstep) */

static int sampleb_batch_core(double *b, size_t C, int I, double shape, double scale, const scnt_int *N,
                              const scnt_int *T, const double *apar, uint64_t *rng, stb_rand31_t *rnd, int loops,
                              stb_sample_stats *st) {
  double *Q = NULL, *lo = NULL, *hi = NULL, *x = NULL, *xprime = NULL, *dsum = NULL, *xs = NULL, *as = NULL;
  int *idx = NULL, *bl = NULL;
  size_t c, cnt;
  int rc = -1, i;
  float ms = 0.f;
  double Tsum0 = shape;
  BBatch bb;
  memset(&bb, 0, sizeof bb);
  if (!C) return 0;
  if (scale <= 0) {
    fprintf(stderr, "Illegal scale in sampleb()\n");
    return -1;
  }
  Q = (double *)malloc(sizeof(double) * C);
  lo = (double *)malloc(sizeof(double) * C);
  hi = (double *)malloc(sizeof(double) * C);
  x = (double *)malloc(sizeof(double) * C);
  xprime = (double *)malloc(sizeof(double) * C);
  dsum = (double *)malloc(sizeof(double) * C);
  xs = (double *)malloc(sizeof(double) * C);
  as = (double *)malloc(sizeof(double) * C);
  idx = (int *)malloc(sizeof(int) * C);
  bl = (int *)malloc(sizeof(int) * C);
  bb.qv = (double *)malloc(sizeof(double) * C * (B_SPECULATE + 1));
  bb.av = (double *)malloc(sizeof(double) * C * (B_SPECULATE + 1));
  if (!Q || !lo || !hi || !x || !xprime || !dsum || !xs || !as || !idx || !bl || !bb.qv || !bb.av) goto done;
  bb.ps = stb_cuda_pstat_create(I, T, N, NULL, 0, C * (B_SPECULATE + 1));
  if (!bb.ps) goto done;
  /* auxiliary variables: Q_c = 1/scale - sum_i log q_i, q_i ~ Beta(b_c, N_i)  (lib/sampleb.c:90-100) */
  if (stb_cuda_pstat_betaQ(bb.ps, b, rng, C, scale, Q, &ms)) goto done;
  if (st) st->eval_ms += ms;
  for (c = 0; c < C; c++)
    if (!(Q[c] == Q[c])) {
      fprintf(stderr, "Illegal q in sampleb(b=%lf)\n", b[c]);
      rc = 1 + (int)c;
      goto done;
    }
  for (i = 0; i < I; i++) Tsum0 += T[i];
  /* a == 0 chains: closed-form gamma draw (lib/sampleb.c:101-118), on the host from the chain's stream */
  for (c = 0; c < C; c++) {
    if (apar[c] == 0) {
      double myb;
      if (Tsum0 > 400) {
        do {
          myb = Tsum0 + stb_rng48_gaussian(&rng[c], 1) * sqrt(Tsum0);
        } while (myb <= 0);
      } else
        myb = stb_rng48_gamma(&rng[c], Tsum0);
      myb /= Q[c];
      if (myb < B_MIN) myb = B_MIN;
      if (myb > B_MAX) myb = B_MAX;
      b[c] = myb;
    }
  }
  /* a > 0 chains: bmax warm-up in lock-step (lib/sampleb.c:51-68), then the slice sampler; in the ARS
   * configuration there is no warm-up (lib/sampleb.c:127-140) */
  for (c = 0; c < C; c++) {
    bl[c] = rnd ? 0 : B_LOOPS;
    xprime[c] = b[c];
    if (apar[c] != 0 && !rnd) {
      if (b[c] <= 0) {
        fprintf(stderr, "Illegal concentration value in bmax()\n");
        rc = 1 + (int)c;
        goto done;
      }
      xprime[c] = b[c];
      x[c] = b[c] * 1.1;
    }
  }
  for (;;) {
    cnt = 0;
    for (c = 0; c < C; c++) {
      if (apar[c] == 0 || bl[c] <= 0) continue;
      if (fabs((x[c] - xprime[c]) / x[c]) > B_ERROR && --bl[c] > 0) {
        xs[cnt] = x[c];
        as[cnt] = apar[c];
        idx[cnt++] = (int)c;
      } else
        bl[c] = 0;
    }
    if (!cnt) break;
    if (stb_cuda_pstat_bterms(bb.ps, xs, NULL, as, shape, cnt, 1, dsum, &ms)) goto done;
    if (st) {
      st->eval_ms += ms;
      st->rounds++;
    }
    for (size_t j = 0; j < cnt; j++) {
      double val;
      c = (size_t)idx[j];
      val = (shape - 1) * apar[c] / x[c] - Q[c] * apar[c] + dsum[j];
      x[c] = xprime[c];
      xprime[c] = apar[c] * stb_digamma_inv(val / I);
    }
  }
  /* slice sampler over [B_MIN, B_MAX] from the warm start, only the a > 0 chains */
  cnt = 0;
  for (c = 0; c < C; c++)
    if (apar[c] != 0) {
      xs[cnt] = xprime[c];
      lo[cnt] = B_MIN;
      hi[cnt] = B_MAX;
      idx[cnt++] = (int)c;
    }
  if (cnt) {
    /* compact the active chains: slice_lockstep indexes chains 0..cnt-1 */
    uint64_t *r2 = (uint64_t *)malloc(sizeof(uint64_t) * cnt);
    double *Q2 = (double *)malloc(sizeof(double) * cnt), *a2 = (double *)malloc(sizeof(double) * cnt);
    stb_sample_stats sub, *sp = NULL;
    if (!r2 || !Q2 || !a2) {
      free(r2);
      free(Q2);
      free(a2);
      goto done;
    }
    for (size_t j = 0; j < cnt; j++) {
      r2[j] = rng[idx[j]];
      Q2[j] = Q[idx[j]];
      a2[j] = apar[idx[j]];
    }
    bb.Q = Q2;
    bb.apar = a2;
    bb.shape = shape;
    if (st) {
      sub = *st;
      if (cnt != C) sub.trace_x = sub.trace_v = NULL, sub.trace_n = NULL; /* traces only when every chain is active */
      sp = &sub;
    }
    bb.st = sp;
    if (rnd) {
      stb_rand31_t *g2 = (stb_rand31_t *)malloc(sizeof(stb_rand31_t) * cnt);
      if (!g2)
        rc = -1;
      else {
        for (size_t j = 0; j < cnt; j++) g2[j] = rnd[idx[j]];
        rc = ars_lockstep(xs, cnt, lo, hi, g2, bterms_batch, &bb, sp, NULL);
        for (size_t j = 0; j < cnt; j++) rnd[idx[j]] = g2[j];
        for (size_t j = 0; rc == 0 && j < cnt; j++)
          if (xs[j] < B_MIN || xs[j] > B_MAX) {
            fprintf(stderr, "Arms_simple(bpar) returned value out of bounds (chain %d)\n", idx[j]);
            rc = 1 + (int)j;
          }
        free(g2);
      }
    } else
      rc = slice_lockstep(xs, cnt, lo, hi, r2, loops, bterms_batch, &bb, sp, B_SPECULATE, 0);
    if (st) {
      st->evals = sub.evals;
      st->rounds = sub.rounds;
      st->eval_ms = sub.eval_ms;
    }
    if (rc > 0) rc = 1 + idx[rc - 1];
    for (size_t j = 0; j < cnt; j++) {
      rng[idx[j]] = r2[j];
      b[idx[j]] = xs[j];
    }
    free(r2);
    free(Q2);
    free(a2);
  } else
    rc = 0;
done:
  if (bb.ps) stb_cuda_pstat_destroy(bb.ps);
  free(Q);
  free(lo);
  free(hi);
  free(x);
  free(xprime);
  free(dsum);
  free(xs);
  free(as);
  free(idx);
  free(bl);
  free(bb.qv);
  free(bb.av);
  return rc;
}

int stb_sampleb_batch(double *b, size_t C, int I, double shape, double scale, const scnt_int *N, const scnt_int *T,
                      const double *apar, uint64_t *rng, int loops, stb_sample_stats *st) {
  return sampleb_batch_core(b, C, I, shape, scale, N, T, apar, rng, NULL, loops, st);
}

int stb_sampleb_batch_ars(double *b, size_t C, int I, double shape, double scale, const scnt_int *N,
                          const scnt_int *T, const double *apar, uint64_t *rng, stb_rand31_t *rnd,
                          stb_sample_stats *st) {
  if (!rnd) return -1;
  return sampleb_batch_core(b, C, I, shape, scale, N, T, apar, rng, rnd, 1, st);
}
