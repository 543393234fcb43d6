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
#include "digamma.h"
#include "lgamma.h"
#include "psample.h"
#include "psample_core.h"
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
typedef stb_eval_fn eval_fn;

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
int stb_slice_lockstep(double *xp, size_t C, const double *lo, const double *hi, uint64_t *rng, int loops, eval_fn eval,
                       void *ctx, stb_sample_stats *st, int depth, size_t slots) {
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
int stb_ars_lockstep(double *xp, size_t C, const double *lo, const double *hi, const stb_ars_source *src, eval_fn eval,
                     void *ctx, stb_sample_stats *st, size_t *nfailed) {
  stb_rand31_t *rnd = src ? src->streams : NULL; /* NULL: glibc's rand(), one chain */
  stb_ars_t **m = (stb_ars_t **)calloc(C, sizeof *m);
  int *state = (int *)malloc(sizeof(int) * C), *chain = (int *)malloc(sizeof(int) * C);
  double *want = (double *)malloc(sizeof(double) * C), *xq = (double *)malloc(sizeof(double) * C);
  double *val = (double *)malloc(sizeof(double) * C);
  size_t c, cnt;
  int rc = 0;
  if (!m || !state || !chain || !want || !xq || !val || (!rnd && C > 1)) {
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
    state[c] = stb_ars_begin(m[c], xinit, 3, lo[c], hi[c], 1.0, 0, 0.0, &xp[c], 1, rnd ? stb_rand31_unit : NULL,
                             rnd ? (void *)&rnd[c] : NULL, &want[c]);
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
  stb_ars_source src;
  h.f = myfunc;
  h.data = mydata;
  src.streams = rnd;
  if (!C) return 0;
  if (!rnd) return -1;
  return stb_ars_lockstep(x, C, lo, hi, &src, host_density_eval, &h, NULL, nfailed);
}

/* ------------------------------------------------------------------------------------------ */
/* the discount: log-posterior evaluators                                                      */
/* ------------------------------------------------------------------------------------------ */
#define A_SPECULATE 8 /* at most; only into table slots a round would leave empty (see stb_slice_lockstep) */

/* what both back ends share: the statistics with n > 1 flattened in (i,k) order (restaurant i owns
 * entries row[i] .. row[i+1]-1) and the table extent they need */
typedef struct {
  int I;
  const scnt_int *T;
  const double *bpar;
  int bpar_per_chain;
  uint32_t *nn, *tt;
  size_t *row, npairs;
  unsigned Nx, Mx;
  stb_sample_stats *st;
  int verbose;
  /* device: a sweep context (one table per evaluation, many per launch) and the lgamma reduction */
  stb_sweep_t *sweep;
  stb_pstat_dev_t *ps;
  double *ssum, *lg;
  /* host: ONE table handle, refilled per evaluation, and its look-ups */
  stable_t *S;
  double *cell;
} DiscountPost;

/* device back end: every point is a table of a discount sweep, reduced against the statistics on the device */
static int discount_post_device(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  DiscountPost *d = (DiscountPost *)ctx;
  float ms = 0.f;
  size_t j;
  for (j = 0; j < cnt; j++)
    if (x[j] <= 0) {
      fprintf(stderr, "samplea: the sampler proposed a discount <= 0\n");
      return 1;
    }
  if (stb_sweep_run(d->sweep, x, cnt, NULL, d->ssum, NULL)) return 1;
  if (d->st) d->st->eval_ms += stb_sweep_last_fill_ms(d->sweep);
  if (stb_cuda_pstat_aterms_lg(d->ps, x, chain, cnt, d->bpar_per_chain, d->lg, &ms)) return 1;
  if (d->st) d->st->eval_ms += ms;
  for (j = 0; j < cnt; j++) out[j] = d->lg[j] + d->ssum[j];
  return 0;
}

/*
 * host back end (the scalar samplea): per point one refill of the chain's table at the proposed discount
 * by the CUDA engine and one batched look-up (what lib/samplea.c:57-60, 74-79 do with S_remake and a call
 * of S_S per node), then the sum in the reference's order with libm -- restaurant by restaurant, its
 * lgamma terms first, then its table cells -- so that the value has the reference's rounding.
 */
static int discount_post_host(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  DiscountPost *d = (DiscountPost *)ctx;
  size_t j, p;
  int i;
  for (j = 0; j < cnt; j++) {
    const double xv = x[j], lx = log(xv);
    const double *conc = d->bpar + (d->bpar_per_chain ? (size_t)chain[j] * (size_t)d->I : 0);
    double acc = 0;
    if (!(xv > 0)) {
      fprintf(stderr, "samplea: the sampler proposed a discount <= 0\n");
      return 1;
    }
    if (d->verbose > 1) fprintf(stderr, "samplea: table %ux%u at a=%lf\n", d->Nx, d->Mx, xv);
    if (d->S ? S_remake(d->S, xv) : !(d->S = S_make(d->Nx, d->Mx, d->Nx, d->Mx, xv, S_STABLE | S_NOMIRROR))) {
      fprintf(stderr, "samplea: no table: %s\n", stb_last_error());
      return 1;
    }
    if (d->npairs && stb_S_batch(d->S, d->nn, d->tt, d->cell, d->npairs)) {
      fprintf(stderr, "samplea: table look-up failed: %s\n", stb_last_error());
      return 1;
    }
    for (i = 0; i < d->I; i++) {
      const double r = conc[i] / xv;
      acc += d->T[i] * lx + lgamma(d->T[i] + r) - lgamma(r);
      for (p = d->row[i]; p < d->row[i + 1]; p++) acc += d->cell[p];
    }
    out[j] = acc;
  }
  return 0;
}

/*
 * The sweep handle (resident slabs for one launch's worth of tables: ~1 GB for the config-4 shape)
 * is kept between calls, one per device -- an MCMC run calls samplea once per sweep with the same
 * table extent, and allocating / freeing a gigabyte of device memory per call costs more than the
 * evaluations.  stb_release_caches() frees them.  A handle in use is never in its slot, so the
 * per-device workers of stb_samplea_batch_multi (multi.c) may run concurrently; the slots are guarded
 * by a mutex.
 */
#include <pthread.h>
#define STB_MAX_DEVICES 64
static struct {
  stb_sweep_t *w;
  unsigned N, M;
} g_sweep_cache[STB_MAX_DEVICES];
static pthread_mutex_t g_sweep_cache_mutex = PTHREAD_MUTEX_INITIALIZER;

/* exchange slot `dev`: returns what was there (with its extent in *oN, *oM), leaves `put` */
static stb_sweep_t *sweep_cache_swap(int dev, stb_sweep_t *put, unsigned N, unsigned M, unsigned *oN, unsigned *oM) {
  stb_sweep_t *old;
  if (dev < 0 || dev >= STB_MAX_DEVICES) return put;
  pthread_mutex_lock(&g_sweep_cache_mutex);
  old = g_sweep_cache[dev].w;
  if (oN) *oN = g_sweep_cache[dev].N;
  if (oM) *oM = g_sweep_cache[dev].M;
  g_sweep_cache[dev].w = put;
  g_sweep_cache[dev].N = N;
  g_sweep_cache[dev].M = M;
  pthread_mutex_unlock(&g_sweep_cache_mutex);
  return old;
}

void stb_release_caches(void) {
  int dev;
  for (dev = 0; dev < STB_MAX_DEVICES; dev++) {
    stb_sweep_t *w = sweep_cache_swap(dev, NULL, 0, 0, NULL, NULL);
    if (w) stb_sweep_free(w);
  }
  stb_cuda_pstat_purge();
}

static stb_sweep_t *sweep_acquire(unsigned N, unsigned M) {
  unsigned cN = 0, cM = 0;
  stb_sweep_t *w = sweep_cache_swap(stb_cuda_current_device(), NULL, 0, 0, &cN, &cM);
  if (w && cN == N && cM == M) return w;
  if (w) stb_sweep_free(w);
  return stb_sweep_create(N, M, 0);
}

static void sweep_release(stb_sweep_t *w, unsigned N, unsigned M) {
  stb_sweep_t *old = sweep_cache_swap(stb_cuda_current_device(), w, N, M, NULL, NULL);
  if (old) stb_sweep_free(old);
}

/* [lo, hi] of a discount move from a: the prior's support [A_MIN, A_MAX], a start within 1e-5 of an end pulled
 * inside, a move of at most SQUEEZEA downwards -- and upwards too for ARS, while the slice sampler's bracket
 * reaches A_MAX (lib/samplea.c:161-177, :217) */
static void discount_bracket(double a, int ars, double *lo, double *hi) {
  double mid = a;
  if (fabs(mid - A_MAX) / A_MAX < 0.00001) mid = A_MAX * 0.999 + A_MIN * 0.001;
  if (fabs(mid - A_MIN) / A_MIN < 0.00001) mid = A_MIN * 0.999 + A_MAX * 0.001;
  *lo = (mid - SQUEEZEA > A_MIN) ? mid - SQUEEZEA : A_MIN;
  *hi = (ars && mid + SQUEEZEA < A_MAX) ? mid + SQUEEZEA : A_MAX;
}

int stb_discount_step(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                      const double *bpar, int bpar_per_chain, uint64_t *rng, const stb_ars_source *ars, int loops,
                      stb_sample_stats *st, int backend, int verbose) {
  const int host = backend == STB_BACKEND_HOST;
  double *lo = NULL, *hi = NULL;
  size_t total = 0, c, slots = 0;
  int i, k, maxn = 1, maxt = 1, rc = -1;
  DiscountPost d;
  memset(&d, 0, sizeof d);
  if (!C) return 0;
  for (i = 0; i < I; i++) total += (size_t)K[i];
  lo = (double *)malloc(sizeof(double) * C);
  hi = (double *)malloc(sizeof(double) * C);
  d.nn = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  d.tt = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  d.row = (size_t *)malloc(sizeof(size_t) * ((size_t)I + 1));
  if (!lo || !hi || !d.nn || !d.tt || !d.row) goto done;
  for (c = 0; c < C; c++) discount_bracket(a[c], ars != NULL, &lo[c], &hi[c]);
  /* nodes with n > 1 enter the sum; the table covers max n + 1 rows, max t + 1 columns (lib/samplea.c:186-208) */
  for (i = 0; i < I; i++) {
    d.row[i] = d.npairs;
    for (k = 0; k < K[i]; k++) {
      if ((int)t[i][k] >= maxt) maxt = t[i][k] + 1;
      if ((int)n[i][k] >= maxn) maxn = n[i][k] + 1;
      if (n[i][k] > 1) {
        d.nn[d.npairs] = n[i][k];
        d.tt[d.npairs] = t[i][k];
        d.npairs++;
      }
    }
  }
  d.row[I] = d.npairs;
  d.Mx = maxt < 10 ? 10u : (unsigned)maxt; /* the clamps S_make applies (lib/stable.c:118-129) */
  d.Nx = (unsigned)maxn < d.Mx ? d.Mx : (unsigned)maxn;
  d.I = I;
  d.T = T;
  d.bpar = bpar;
  d.bpar_per_chain = bpar_per_chain;
  d.st = st;
  d.verbose = verbose;
  if (host) {
    d.cell = (double *)malloc(sizeof(double) * (d.npairs ? d.npairs : 1));
    if (!d.cell) goto done;
  } else {
    d.sweep = sweep_acquire(d.Nx, d.Mx);
    if (!d.sweep || stb_sweep_set_pairs(d.sweep, d.nn, d.tt, d.npairs)) goto done;
    /* a round evaluates at most one point per chain plus what its last wave of tables has room for */
    slots = (size_t)stb_sweep_tables_in_flight(d.sweep);
    d.ssum = (double *)malloc(sizeof(double) * (C + slots));
    d.lg = (double *)malloc(sizeof(double) * (C + slots));
    if (!d.ssum || !d.lg) goto done;
    d.ps = stb_cuda_pstat_create(I, T, NULL, bpar, bpar_per_chain ? C * (size_t)I : (size_t)I, C + slots);
    if (!d.ps) goto done;
  }
  {
    const eval_fn post = host ? discount_post_host : discount_post_device;
    if (ars) {
      rc = stb_ars_lockstep(a, C, lo, hi, ars, post, &d, st, NULL);
      for (c = 0; rc == 0 && c < C; c++)
        if (a[c] < lo[c] || a[c] > hi[c]) {
          fprintf(stderr, "samplea: arms_simple left [%lg, %lg] (chain %zu)\n", lo[c], hi[c], c);
          rc = 1 + (int)c;
        }
    } else
      rc = stb_slice_lockstep(a, C, lo, hi, rng, loops, post, &d, st, host ? 1 : A_SPECULATE, slots);
  }
done:
  if (d.sweep) sweep_release(d.sweep, d.Nx, d.Mx);
  if (d.ps) stb_cuda_pstat_destroy(d.ps);
  if (d.S) S_free(d.S);
  free(lo);
  free(hi);
  free(d.nn);
  free(d.tt);
  free(d.row);
  free(d.cell);
  free(d.ssum);
  free(d.lg);
  return rc;
}

int stb_samplea_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                      const double *bpar, int bpar_per_chain, uint64_t *rng, int loops, stb_sample_stats *st) {
  return stb_discount_step(a, C, I, K, T, n, t, bpar, bpar_per_chain, rng, NULL, loops, st, STB_BACKEND_DEVICE, 0);
}

int stb_samplea_batch_ars(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                          const double *bpar, int bpar_per_chain, stb_rand31_t *rnd, stb_sample_stats *st) {
  stb_ars_source src;
  if (!rnd) return -1;
  src.streams = rnd;
  return stb_discount_step(a, C, I, K, T, n, t, bpar, bpar_per_chain, NULL, &src, 1, st, STB_BACKEND_DEVICE, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* the concentration                                                                           */
/* ------------------------------------------------------------------------------------------ */
#define B_ERROR 1.0e-4
#define B_LOOPS 5
#define B_SPECULATE 4 /* slice proposals evaluated per chain and round (see stb_slice_lockstep) */

/* the three I-term passes of a concentration update, on the device (one block per evaluation, tree sums)
 * or on the host (libm, sequential sums from the value the reference starts its sum with) */
typedef struct {
  int host, I;
  const scnt_int *T, *N;
  stb_pstat_dev_t *ps;
  /* the slice / ARS stage: per ACTIVE chain */
  const double *Q, *apar;
  double shape;
  double *qv, *av;
  stb_sample_stats *st;
} ConcPost;

/* Q_c = 1/scale - sum_i log q_i,  q_i ~ Beta(b_c, N_i) from chain c's stream (lib/sampleb.c:90-100); NaN: a q <= 0 */
static int conc_aux_Q(ConcPost *p, const double *b, uint64_t *rng, size_t C, double scale, double *Q, float *ms) {
  size_t c;
  int i;
  if (!p->host) return stb_cuda_pstat_betaQ(p->ps, b, rng, C, scale, Q, ms);
  for (c = 0; c < C; c++) {
    stb_rng48 r;
    double q = 1.0 / scale;
    int bad = 0;
    r.x = rng[c];
    for (i = 0; i < p->I; i++) {
      double v;
      if (p->N[i] == 0) continue;
      v = stb_beta(&r, stb_zig_tables_get(), b[c], (double)(int)p->N[i]);
      if (!(v > 0)) bad = 1;
      q -= log(v);
    }
    rng[c] = r.x;
    Q[c] = bad ? NAN : q;
  }
  return 0;
}

/* out[j] = init[j] + sum_i digamma(T_i + x_j / a_j): one fixed-point step of the mode search (lib/sampleb.c:59-62) */
static int conc_digamma_sums(ConcPost *p, const double *x, const double *apar, const double *init, size_t cnt,
                             double *out, float *ms) {
  size_t j;
  int i;
  if (!p->host) {
    if (stb_cuda_pstat_bterms(p->ps, x, NULL, apar, 0.0, cnt, 1, out, ms)) return 1;
    for (j = 0; j < cnt; j++) out[j] = init[j] + out[j];
    return 0;
  }
  for (j = 0; j < cnt; j++) {
    const double xa = x[j] / apar[j];
    double acc = init[j];
    for (i = 0; i < p->I; i++) acc += digamma(p->T[i] + xa);
    out[j] = acc;
  }
  return 0;
}

/* log p(b = x | ...) up to a constant (lib/sampleb.c:33-41): -Q x + (shape - 1) log x + sum_i [lgamma(T_i + x/a) - lgamma(x/a)] */
static int conc_post(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  ConcPost *p = (ConcPost *)ctx;
  float ms = 0.f;
  size_t j;
  int i;
  if (p->host) {
    for (j = 0; j < cnt; j++) {
      const double xa = x[j] / p->apar[chain[j]], base = lgamma(xa);
      double acc = -p->Q[chain[j]] * x[j] + (p->shape - 1) * log(x[j]);
      for (i = 0; i < p->I; i++) acc += lgamma(p->T[i] + xa) - base;
      out[j] = acc;
    }
    return 0;
  }
  for (j = 0; j < cnt; j++) {
    p->qv[j] = p->Q[chain[j]];
    p->av[j] = p->apar[chain[j]];
  }
  if (stb_cuda_pstat_bterms(p->ps, x, p->qv, p->av, p->shape, cnt, 0, out, &ms)) return 1;
  if (p->st) p->st->eval_ms += ms;
  return 0;
}

int stb_concentration_step(double *b, size_t C, int I, double shape, double scale, const scnt_int *N, const scnt_int *T,
                           const double *apar, uint64_t *rng, const stb_ars_source *ars, int loops, stb_sample_stats *st,
                           int backend, int verbose) {
  const int host = backend == STB_BACKEND_HOST, depth = host ? 1 : B_SPECULATE;
  double *Q = NULL, *lo = NULL, *hi = NULL, *x = NULL, *xprime = NULL, *dsum = NULL, *xs = NULL, *as = NULL, *init = NULL;
  int *idx = NULL, *bl = NULL;
  size_t c, cnt;
  int rc = -1, i;
  float ms = 0.f;
  double Tsum0 = shape;
  ConcPost bb;
  memset(&bb, 0, sizeof bb);
  if (!C) return 0;
  if (scale <= 0) {
    fprintf(stderr, "sampleb: the prior's scale must be positive\n");
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
  init = (double *)malloc(sizeof(double) * C);
  idx = (int *)malloc(sizeof(int) * C);
  bl = (int *)malloc(sizeof(int) * C);
  bb.qv = (double *)malloc(sizeof(double) * C * (size_t)(depth + 1));
  bb.av = (double *)malloc(sizeof(double) * C * (size_t)(depth + 1));
  if (!Q || !lo || !hi || !x || !xprime || !dsum || !xs || !as || !init || !idx || !bl || !bb.qv || !bb.av) goto done;
  bb.host = host;
  bb.I = I;
  bb.T = T;
  bb.N = N;
  if (!host) {
    bb.ps = stb_cuda_pstat_create(I, T, N, NULL, 0, C * (size_t)(depth + 1));
    if (!bb.ps) goto done;
  }
  if (conc_aux_Q(&bb, b, rng, C, scale, Q, &ms)) goto done;
  if (st) st->eval_ms += ms;
  for (c = 0; c < C; c++)
    if (!(Q[c] == Q[c])) {
      fprintf(stderr, "sampleb: an auxiliary Beta(%lf, N) draw came out as 0 (chain %zu)\n", b[c], c);
      rc = 1 + (int)c;
      goto done;
    }
  for (i = 0; i < I; i++) Tsum0 += T[i];
  /* a == 0 chains: the posterior is Gamma(shape + sum T, Q), drawn directly from the chain's stream; a Gaussian
   * stands in for the Gamma when its shape passes 400 (lib/sampleb.c:101-118) */
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
      if (verbose > 1) fprintf(stderr, "sampleb: b ~ Gamma(%lg, %lg) -> %lf\n", Tsum0, Q[c], myb);
      b[c] = myb;
    }
  }
  /* a > 0 chains: the mode search warms the slice sampler up, in lock-step (lib/sampleb.c:51-68); the ARS
   * configuration starts from its own abscissae instead (lib/sampleb.c:127-140) */
  for (c = 0; c < C; c++) {
    bl[c] = ars ? 0 : B_LOOPS;
    xprime[c] = b[c];
    if (apar[c] != 0 && !ars) {
      if (b[c] <= 0) {
        fprintf(stderr, "sampleb: the concentration must be positive (chain %zu)\n", c);
        rc = 1 + (int)c;
        goto done;
      }
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
        init[cnt] = (shape - 1) * apar[c] / x[c] - Q[c] * apar[c];
        idx[cnt++] = (int)c;
      } else
        bl[c] = 0;
    }
    if (!cnt) break;
    if (conc_digamma_sums(&bb, xs, as, init, cnt, dsum, &ms)) goto done;
    if (st) {
      st->eval_ms += ms;
      st->rounds++;
    }
    for (size_t j = 0; j < cnt; j++) {
      c = (size_t)idx[j];
      x[c] = xprime[c];
      xprime[c] = apar[c] * stb_digamma_inv(dsum[j] / I);
    }
  }
  /* the sampler over [B_MIN, B_MAX], only the a > 0 chains, compacted (the drivers index chains 0..cnt-1) */
  cnt = 0;
  for (c = 0; c < C; c++)
    if (apar[c] != 0) {
      if (verbose > 1 && !ars) fprintf(stderr, "sampleb: mode search %lg -> %lg (Q = %lg)\n", b[c], xprime[c], Q[c]);
      xs[cnt] = xprime[c];
      lo[cnt] = B_MIN;
      hi[cnt] = B_MAX;
      idx[cnt++] = (int)c;
    }
  if (cnt) {
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
    if (ars) {
      stb_ars_source sub_src;
      stb_rand31_t *g2 = ars->streams ? (stb_rand31_t *)malloc(sizeof(stb_rand31_t) * cnt) : NULL;
      sub_src.streams = g2;
      if (ars->streams && !g2)
        rc = -1;
      else {
        for (size_t j = 0; g2 && j < cnt; j++) g2[j] = ars->streams[idx[j]];
        rc = stb_ars_lockstep(xs, cnt, lo, hi, &sub_src, conc_post, &bb, sp, NULL);
        for (size_t j = 0; g2 && j < cnt; j++) ars->streams[idx[j]] = g2[j];
        for (size_t j = 0; rc == 0 && j < cnt; j++)
          if (xs[j] < B_MIN || xs[j] > B_MAX) {
            fprintf(stderr, "sampleb: arms_simple left [%lg, %d] (chain %d)\n", B_MIN, B_MAX, idx[j]);
            rc = 1 + (int)j;
          }
        free(g2);
      }
    } else
      rc = stb_slice_lockstep(xs, cnt, lo, hi, r2, loops, conc_post, &bb, sp, depth, 0);
    if (st) {
      st->evals = sub.evals;
      st->rounds = sub.rounds;
      st->eval_ms = sub.eval_ms;
    }
    if (rc > 0) rc = 1 + idx[rc - 1];
    for (size_t j = 0; j < cnt; j++) {
      rng[idx[j]] = r2[j];
      b[idx[j]] = xs[j];
      if (verbose > 1) fprintf(stderr, "sampleb: b | Q = %lg -> %lf\n", Q2[j], xs[j]);
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
  free(init);
  free(idx);
  free(bl);
  free(bb.qv);
  free(bb.av);
  return rc;
}

int stb_sampleb_batch(double *b, size_t C, int I, double shape, double scale, const scnt_int *N, const scnt_int *T,
                      const double *apar, uint64_t *rng, int loops, stb_sample_stats *st) {
  return stb_concentration_step(b, C, I, shape, scale, N, T, apar, rng, NULL, loops, st, STB_BACKEND_DEVICE, 0);
}

int stb_sampleb_batch_ars(double *b, size_t C, int I, double shape, double scale, const scnt_int *N,
                          const scnt_int *T, const double *apar, uint64_t *rng, stb_rand31_t *rnd,
                          stb_sample_stats *st) {
  stb_ars_source src;
  if (!rnd) return -1;
  src.streams = rnd;
  return stb_concentration_step(b, C, I, shape, scale, N, T, apar, rng, &src, 1, st, STB_BACKEND_DEVICE, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* the discount without a table refill per evaluation (samplea2)                               */
/* ------------------------------------------------------------------------------------------ */
/*
 * lib/samplea.c:227-341.  Step 1 samples, for every node with 1 < t < n, how its n customers split over
 * its t tables (ratios of S_S values of the CALLER's table: stb_partition_sample, one kernel for all
 * nodes).  Given the sizes the likelihood of the discount is a product of rising factorials (1-a)_{s-1},
 * one per table of size s > 1, so step 2 -- one slice / ARS step -- needs no Stirling numbers.
 *
 * The sizes do not change during step 2, so the log-posterior's table terms are a fixed LIST of
 * arguments j = s - 1 of  lgamma(j + 1 - a) - lgamma(1 - a), restaurant by restaurant in the order the
 * reference visits them (:115-143): a node with t = 1 gives n - 1; a node with sampled sizes gives each
 * size > 1 from the last sampled to the first, then whatever is left of n.  The evaluator walks the list.
 */
typedef struct {
  int I;
  const scnt_int *T;
  const double *bpar;
  size_t *row;  /* restaurant i owns arg[row[i] .. row[i+1]-1] */
  int *arg;
} PartitionPost;

static int partition_post_host(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  PartitionPost *d = (PartitionPost *)ctx;
  size_t j, p;
  int i;
  (void)chain;
  for (j = 0; j < cnt; j++) {
    const double xv = x[j], lx = log(xv);
    struct gcache_s rising; /* lgamma(j + 1 - a) - lgamma(1 - a), cached for small j (include/lgamma.h) */
    double acc = 0;
    if (!(xv > 0)) {
      fprintf(stderr, "samplea2: the sampler proposed a discount <= 0\n");
      return 1;
    }
    gcache_init(&rising, 1 - xv);
    for (i = 0; i < d->I; i++) {
      const double r = d->bpar[i] / xv;
      acc += d->T[i] * lx + lgamma(d->T[i] + r) - lgamma(r);
      for (p = d->row[i]; p < d->row[i + 1]; p++) acc += gcache_value(&rising, d->arg[p]);
    }
    out[j] = acc;
  }
  return 0;
}

int stb_discount_step_partition(double *a, stable_t *S, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                                const double *bpar, uint64_t *rng, const stb_ars_source *ars, int loops, int exact,
                                int verbose) {
  size_t nodes = 0, sizes = 0, nargs = 0, j = 0, o = 0;
  uint32_t *pn = NULL, *poff = NULL;
  uint16_t *pt = NULL, *m = NULL;
  double *logu = NULL, lo, hi;
  stb_rng48 r;
  PartitionPost d;
  int i, k, rc = -1;
  (void)verbose;
  memset(&d, 0, sizeof d);
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++) {
      const scnt_int nk = n[i][k];
      const stcnt_int tk = t[i][k];
      if (tk > 1 && tk < nk) {
        nodes++;
        sizes += (size_t)tk - 1;
        nargs += (size_t)tk; /* at most t - 1 sampled sizes and the remainder */
      } else if (nk != 0 && tk != nk)
        nargs++;
    }
  pn = (uint32_t *)malloc(sizeof(uint32_t) * (nodes ? nodes : 1));
  poff = (uint32_t *)malloc(sizeof(uint32_t) * (nodes ? nodes : 1));
  pt = (uint16_t *)calloc(nodes ? nodes : 1, sizeof(uint16_t));
  m = (uint16_t *)malloc(sizeof(uint16_t) * (sizes ? sizes : 1));
  logu = (double *)calloc(exact ? (sizes ? sizes : 1) : (nodes ? nodes : 1), sizeof(double));
  d.row = (size_t *)malloc(sizeof(size_t) * ((size_t)I + 1));
  d.arg = (int *)malloc(sizeof(int) * (nargs ? nargs : 1));
  if (!pn || !poff || !pt || !m || !logu || !d.row || !d.arg) goto done;
  /* the uniforms come off the chain's stream in (i,k) order: one per node (:293), or one per round in the exact mode */
  r.x = *rng;
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++)
      if (t[i][k] > 1 && t[i][k] < n[i][k]) {
        pn[j] = n[i][k];
        pt[j] = t[i][k];
        poff[j] = (uint32_t)o;
        if (exact) {
          int M;
          for (M = (int)t[i][k] - 1; M >= 1; M--) logu[o + (size_t)M - 1] = log(stb_rng48_unit(&r));
        } else
          logu[j] = log(stb_rng48_unit(&r));
        o += (size_t)t[i][k] - 1;
        j++;
      }
  *rng = r.x;
  if (nodes && stb_partition_sample(S, *a, pn, pt, logu, poff, nodes, m, sizes, exact)) {
    fprintf(stderr, "samplea2: partition sampling failed: %s\n", stb_last_error());
    goto done;
  }
  /* the argument list of the likelihood (see above) */
  nargs = 0;
  o = 0;
  for (i = 0; i < I; i++) {
    d.row[i] = nargs;
    for (k = 0; k < K[i]; k++) {
      scnt_int left = n[i][k];
      const stcnt_int tk = t[i][k];
      if (left == 0 || tk == left) continue;
      if (tk == 1) {
        d.arg[nargs++] = (int)left - 1;
        continue;
      }
      for (int l = (int)tk - 2; l >= 0; l--) {
        if (m[o + (size_t)l] > 1) d.arg[nargs++] = (int)m[o + (size_t)l] - 1;
        left -= m[o + (size_t)l];
      }
      if (left > 0) d.arg[nargs++] = (int)left - 1;
      o += (size_t)tk - 1;
    }
  }
  d.row[I] = nargs;
  d.I = I;
  d.T = T;
  d.bpar = bpar;
  discount_bracket(*a, ars != NULL, &lo, &hi);
  if (ars) {
    rc = stb_ars_lockstep(a, 1, &lo, &hi, ars, partition_post_host, &d, NULL, NULL);
    if (rc == 0 && (*a < lo || *a > hi)) {
      fprintf(stderr, "samplea2: arms_simple left [%lg, %lg]\n", lo, hi);
      rc = 1;
    }
  } else
    rc = stb_slice_lockstep(a, 1, &lo, &hi, rng, loops, partition_post_host, &d, NULL, 1, 0);
done:
  free(pn);
  free(poff);
  free(pt);
  free(m);
  free(logu);
  free(d.row);
  free(d.arg);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* samplea2 for C chains (stb_samplea2_batch)                                                  */
/* ------------------------------------------------------------------------------------------ */
/*
 * Chain c updates its discount a[c] the table-free way (lib/samplea.c:227-341): its own Stirling table at a[c]
 * (one table per chain out of a discount sweep -- the scalar samplea2 is handed the caller's table), the
 * partition step of every node with 1 < t < n against it on the device, the chain's uniforms taken from its
 * stream rng[c] in the order the scalar call takes them (stb_cuda_sweep_partition), then ONE slice / ARS step
 * for all chains in lock-step.  The sampled sizes of a chain are kept as a histogram of the likelihood's
 * arguments j = size - 1 (that is all the likelihood needs), so an evaluation is a device reduction over at most
 * max n bins in a fixed order, next to the same lgamma reduction over the restaurants samplea's batch uses.
 */
typedef struct {
  stb_sweep_t *sweep;
  stb_pstat_dev_t *ps;
  int bpar_per_chain;
  double *lg, *hs;
  stb_sample_stats *st;
} PartitionBatchPost;

static int partition_post_device(void *ctx, const double *x, const int *chain, size_t cnt, double *out) {
  PartitionBatchPost *d = (PartitionBatchPost *)ctx;
  float ms = 0.f;
  size_t j;
  for (j = 0; j < cnt; j++)
    if (!(x[j] > 0)) {
      fprintf(stderr, "samplea2: the sampler proposed a discount <= 0\n");
      return 1;
    }
  if (stb_cuda_pstat_aterms_lg(d->ps, x, chain, cnt, d->bpar_per_chain, d->lg, &ms)) return 1;
  if (d->st) d->st->eval_ms += ms;
  if (stb_sweep_hist_eval(d->sweep, x, chain, cnt, d->hs, &ms)) return 1;
  if (d->st) d->st->eval_ms += ms;
  for (j = 0; j < cnt; j++) out[j] = d->lg[j] + d->hs[j];
  return 0;
}

int stb_discount_step_partition_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n,
                                      stcnt_int **t, const double *bpar, int bpar_per_chain, uint64_t *rng,
                                      const stb_ars_source *ars, int loops, int exact, stb_sample_stats *st,
                                      double *partition_ms) {
  size_t nodes = 0, draws = 0, j = 0, c;
  uint32_t *pn = NULL, *pdraw = NULL, *hbase = NULL;
  uint16_t *pt = NULL;
  double *lo = NULL, *hi = NULL;
  unsigned maxn = 1, maxt = 1, Nx, Mx, hbins;
  PartitionBatchPost d;
  float ms = 0.f;
  int i, k, rc = -1;
  memset(&d, 0, sizeof d);
  if (partition_ms) *partition_ms = 0;
  if (!C) return 0;
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++) {
      if (n[i][k] > maxn) maxn = n[i][k];
      if (t[i][k] > maxt) maxt = t[i][k];
      if (t[i][k] > 1 && t[i][k] < n[i][k]) nodes++;
    }
  Mx = maxt < 10 ? 10u : maxt; /* the clamps S_make applies (lib/stable.c:118-129) */
  Nx = maxn < Mx ? Mx : maxn;
  hbins = maxn + 1;
  pn = (uint32_t *)malloc(sizeof(uint32_t) * (nodes ? nodes : 1));
  pdraw = (uint32_t *)malloc(sizeof(uint32_t) * (nodes ? nodes : 1));
  pt = (uint16_t *)malloc(sizeof(uint16_t) * (nodes ? nodes : 1));
  hbase = (uint32_t *)calloc(hbins, sizeof(uint32_t));
  lo = (double *)malloc(sizeof(double) * C);
  hi = (double *)malloc(sizeof(double) * C);
  if (!pn || !pdraw || !pt || !hbase || !lo || !hi) goto done;
  /* the uniforms come off a chain's stream in (i,k) order: one per node (:293), or one per round in the exact mode;
   * a node with one table puts n - 1 into every chain's argument list (:115-143) */
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++) {
      const scnt_int nk = n[i][k];
      const stcnt_int tk = t[i][k];
      if (tk > 1 && tk < nk) {
        pn[j] = nk;
        pt[j] = tk;
        pdraw[j] = (uint32_t)draws;
        draws += exact ? (size_t)tk - 1 : 1;
        j++;
      } else if (nk != 0 && tk == 1 && nk > 1)
        hbase[nk - 1]++;
    }
  for (c = 0; c < C; c++) discount_bracket(a[c], ars != NULL, &lo[c], &hi[c]);
  d.sweep = sweep_acquire(Nx, Mx);
  if (!d.sweep || stb_sweep_set_nodes(d.sweep, pn, pt, pdraw, nodes, hbase, hbins)) goto done;
  if (stb_sweep_partition(d.sweep, a, C, rng, exact, NULL, &ms)) {
    fprintf(stderr, "samplea2: partition sampling failed: %s\n", stb_last_error());
    goto done;
  }
  if (partition_ms) *partition_ms = ms;
  if (st) st->eval_ms += ms;
  for (c = 0; c < C; c++) rng[c] = stb_cuda_lcg48_jump(rng[c], draws);
  {
    const size_t slots = C * 5 + 64; /* a round evaluates at most depth + 1 = 5 points per chain */
    d.lg = (double *)malloc(sizeof(double) * slots);
    d.hs = (double *)malloc(sizeof(double) * slots);
    if (!d.lg || !d.hs) goto done;
    d.ps = stb_cuda_pstat_create(I, T, NULL, bpar, bpar_per_chain ? C * (size_t)I : (size_t)I, slots);
    if (!d.ps) goto done;
  }
  d.bpar_per_chain = bpar_per_chain;
  d.st = st;
  if (ars) {
    rc = stb_ars_lockstep(a, C, lo, hi, ars, partition_post_device, &d, st, NULL);
    for (c = 0; rc == 0 && c < C; c++)
      if (a[c] < lo[c] || a[c] > hi[c]) {
        fprintf(stderr, "samplea2: arms_simple left [%lg, %lg] (chain %zu)\n", lo[c], hi[c], c);
        rc = 1 + (int)c;
      }
  } else
    rc = stb_slice_lockstep(a, C, lo, hi, rng, loops, partition_post_device, &d, st, 4, 0);
done:
  if (d.sweep) sweep_release(d.sweep, Nx, Mx);
  if (d.ps) stb_cuda_pstat_destroy(d.ps);
  free(pn);
  free(pdraw);
  free(pt);
  free(hbase);
  free(lo);
  free(hi);
  free(d.lg);
  free(d.hs);
  return rc;
}

int stb_samplea2_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                       const double *bpar, int bpar_per_chain, uint64_t *rng, int loops, stb_sample_stats *st) {
  return stb_discount_step_partition_batch(a, C, I, K, T, n, t, bpar, bpar_per_chain, rng, NULL, loops,
                                           stb_get_partition_mode() == STB_PARTITION_EXACT, st, NULL);
}
