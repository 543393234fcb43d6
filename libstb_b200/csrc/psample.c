/*
 * psample.c -- the scalar samplers of psample.h / srng.h / digamma.h (C host code).
 *
 * Mirrors the control flow of the reference's slice-sampler configuration:
 *   SliceSimple  lib/sslice.c:33-80        samplea  lib/samplea.c:46-83, 155-225
 *   sampleb      lib/sampleb.c:33-68, 79-159
 * The random draws consume glibc's global drand48 stream in the reference's order, so a caller
 * that seeds with srand48() sees the same draws.  samplea's table refill per evaluation is the
 * CUDA table engine (S_make / S_remake + one batched gather); everything else here is O(I) host
 * arithmetic, exactly as in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arms.h"
#include "digamma.h"
#include "lgamma.h"
#include "psample.h"
#include "rng48.h"
#include "specfun.h"
#include "stb_b200.h"

/* ------------------------------------------------------------------------------------------ */
/* special functions                                                                           */
/* ------------------------------------------------------------------------------------------ */
double digammaRN(double x) { return stb_digammaRN(x); }
double MLdigamma(double x) { return stb_digamma(x); }
double MLtrigamma(double x) { return stb_trigamma(x); }
double MLtetragamma(double x) { return stb_tetragamma(x); }
double MLpentagamma(double x) { return stb_pentagamma(x); }
double MLpsigamma(double x, double deriv) {
  /* lib/polygamma.c:502-523 rounds the order to the nearest integer; orders 0..3 are what libstb uses */
  if (isnan(x)) return x;
  switch ((int)floor(deriv + 0.5)) {
    case 0: return stb_digamma(x);
    case 1: return stb_trigamma(x);
    case 2: return stb_tetragamma(x);
    case 3: return stb_pentagamma(x);
    default: return NAN;
  }
}
double digammaInv(double x) { return stb_digamma_inv(x); }

/* ------------------------------------------------------------------------------------------ */
/* distributions over glibc's global 48-bit stream                                             */
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

/* take over / hand back glibc's drand48 state: seed48() returns the previous state */
static void rng_import(stb_rng48 *r) {
  unsigned short tmp[3] = {0, 0, 0};
  unsigned short *old = seed48(tmp);
  r->x = (uint64_t)old[0] | ((uint64_t)old[1] << 16) | ((uint64_t)old[2] << 32);
}
static void rng_export(const stb_rng48 *r) {
  unsigned short s[3];
  s[0] = (unsigned short)(r->x & 0xFFFF);
  s[1] = (unsigned short)((r->x >> 16) & 0xFFFF);
  s[2] = (unsigned short)((r->x >> 32) & 0xFFFF);
  seed48(s);
}

double gsl_rng_gaussian_ziggurat(const double sigma) {
  stb_rng48 r;
  double v;
  rng_import(&r);
  v = stb_gauss_zig(&r, stb_zig_tables_get(), sigma);
  rng_export(&r);
  return v;
}
double gsl_rng_gamma(const double a) {
  stb_rng48 r;
  double v;
  rng_import(&r);
  v = stb_gamma(&r, stb_zig_tables_get(), a);
  rng_export(&r);
  return v;
}
double gsl_rng_beta(const double a, const double b) {
  stb_rng48 r;
  double v;
  rng_import(&r);
  v = stb_beta(&r, stb_zig_tables_get(), a, b);
  rng_export(&r);
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* SliceSimple, lib/sslice.c:33-80                                                             */
/* ------------------------------------------------------------------------------------------ */
#define TOOMANY 200

int SliceSimple(double *xp, double (*post)(double, void *), double *bounds, rngp_t rng, int loops, void *pars) {
  double x = *xp, y, range[2];
  int tries;
  (void)rng;
  if (x < bounds[0] || x > bounds[1]) {
    fprintf(stderr, "SliceSimple: input value %lf outside bounds [%lg,%lg]\n", x, bounds[0], bounds[1]);
    return 1;
  }
  while (loops-- > 0) {
    y = post(x, pars);
    range[0] = bounds[0];
    range[1] = bounds[1];
    y += log(rng_unit(rng));
    for (tries = 1; tries < TOOMANY; tries++) {
      x = range[0] + rng_unit(rng) * (range[1] - range[0]);
      if (post(x, pars) > y) {
        *xp = x;
        break;
      }
      /* shrink towards the last accepted point: the posterior is assumed unimodal */
      if (x < *xp)
        range[0] = x;
      else
        range[1] = x;
    }
    if (tries >= TOOMANY) {
      fprintf(stderr, "SliceSimple: giving up after %d tries, range=[%lg,%lg]\n", TOOMANY, range[0], range[1]);
      return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sampleb, lib/sampleb.c                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double shape, Q, apar;
  int I;
  scnt_int *T;
} BLData;

static double bterms(double x, void *mydata) { /* lib/sampleb.c:33-41 */
  BLData *mp = (BLData *)mydata;
  int i;
  double lg = lgamma(x / mp->apar);
  double val = -mp->Q * x + (mp->shape - 1) * log(x);
  for (i = 0; i < mp->I; i++) val += lgamma(mp->T[i] + x / mp->apar) - lg;
  return val;
}

#define B_ERROR 1.0e-4
#define B_LOOPS 5
static double bmax(double x, BLData *mp) { /* lib/sampleb.c:51-68: a few fixed-point steps towards the mode */
  double x_prime = x;
  int loops = B_LOOPS, i;
  if (x <= 0) {
    fprintf(stderr, "Illegal concentration value in bmax()\n");
    exit(1);
  }
  x *= 1.1;
  while (fabs((x - x_prime) / x) > B_ERROR && --loops > 0) {
    double val = (mp->shape - 1) * mp->apar / x - mp->Q * mp->apar;
    for (i = 0; i < mp->I; i++) val += digamma(mp->T[i] + x / mp->apar);
    x = x_prime;
    x_prime = mp->apar * digammaInv(val / mp->I);
  }
  return x_prime;
}

/* which sampler samplea / sampleb / samplea2 run: a compile-time switch in the reference
 * (PSAMPLE_ARS, lib/psample.h:37), a run-time one here.  Unmodified callers pick the reference's
 * default build with STB_SAMPLER=ars in the environment (read once, at the first use). */
static int g_sampler_state = -1;
static int sampler_mode(void) {
  if (g_sampler_state < 0) {
    const char *s = getenv("STB_SAMPLER");
    g_sampler_state = (s && (!strcmp(s, "ars") || !strcmp(s, "ARS") || !strcmp(s, "1"))) ? STB_SAMPLER_ARS : STB_SAMPLER_SLICE;
  }
  return g_sampler_state;
}
#define g_sampler (sampler_mode())
int stb_set_sampler(int which) {
  const int old = sampler_mode();
  g_sampler_state = which == STB_SAMPLER_ARS ? STB_SAMPLER_ARS : STB_SAMPLER_SLICE;
  return old;
}

double sampleb(double b_in, int I, double shape, double scale, scnt_int *N, scnt_int *T, double apar, rngp_t rng,
               int loops, int verbose) {
  double Q, q, myb;
  int i;
  if (scale <= 0) {
    fprintf(stderr, "Illegal scale in sampleb()\n");
    exit(1);
  }
  Q = 1.0 / scale;
  for (i = 0; i < I; i++) {
    if (N[i] <= 0) continue;
    q = rng_beta(rng, b_in, (int)N[i]);
    if (q <= 0) {
      fprintf(stderr, "Illegal q in sampleb(b=%lf)\n", b_in);
      exit(1);
    }
    Q -= log(q);
  }
  if (apar == 0) {
    double Tsum = shape;
    for (i = 0; i < I; i++) Tsum += T[i];
    if (Tsum > 400) { /* the gamma is a narrow Gaussian by now */
      do {
        myb = Tsum + rng_gaussian(rng, 1) * sqrt(Tsum);
      } while (myb <= 0);
    } else
      myb = rng_gamma(rng, Tsum);
    myb /= Q;
    if (myb < B_MIN) myb = B_MIN;
    if (myb > B_MAX) myb = B_MAX;
    if (verbose > 1) fprintf(stderr, "Sample b ~ gamma(%lg,%lg) = %lf\n", Tsum, Q, myb);
  } else {
    double initb[3] = {B_MIN, 1, B_MAX};
    BLData bld;
    bld.Q = Q;
    bld.I = I;
    bld.T = T;
    bld.apar = apar;
    bld.shape = shape;
    if (g_sampler == STB_SAMPLER_ARS) { /* lib/sampleb.c:127-140 */
      initb[1] = b_in;
      if (fabs(initb[1] - B_MAX) / B_MAX < 0.00001) initb[1] = B_MAX * 0.999 + B_MIN * 0.001;
      if (fabs(initb[1] - B_MIN) / B_MIN < 0.00001) initb[1] = B_MIN * 0.999 + B_MAX * 0.001;
      arms_simple(3, initb, initb + 2, bterms, &bld, 0, initb + 1, &myb);
      if (myb < B_MIN || myb > B_MAX) {
        fprintf(stderr, "Arms_simple(bpar) returned value out of bounds\n");
        exit(1);
      }
    } else {
      myb = bmax(b_in, &bld);
      if (verbose > 1) fprintf(stderr, "Max b (%lg,%lg) -> %lg\n", b_in, Q, myb);
      initb[1] = B_MAX;
      if (SliceSimple(&myb, bterms, initb, rng, loops, &bld)) {
        fprintf(stderr, "SliceSimple error\n");
        exit(1);
      }
    }
    if (verbose > 1) fprintf(stderr, "Sample b ~ G(%lg) = %lf\n", Q, myb);
  }
  return myb;
}

/* ------------------------------------------------------------------------------------------ */
/* samplea, lib/samplea.c                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int I;
  int *K;
  scnt_int *T;
  double *bpar;
  int maxt, maxn;
  stable_t *S;
  int verbose;
  /* the statistics with n > 1, flattened once in (i,k) order; first[i] .. first[i+1] belong to i */
  size_t cnt, *first;
  uint32_t *nn, *tt;
  double *val;
} ALData;

static double aterms(double x, void *mydata) { /* lib/samplea.c:46-83 */
  ALData *mp = (ALData *)mydata;
  double val = 0;
  int i;
  size_t j;
  if (x <= 0) {
    fprintf(stderr, "Illegal discount value in aterms()\n");
    exit(1);
  }
  if (mp->verbose > 1) fprintf(stderr, "Extending S for M=%d a=%lf\n", mp->maxt, x);
  if (mp->S)
    S_remake(mp->S, x);
  else
    mp->S = S_make(mp->maxn, mp->maxt, mp->maxn, mp->maxt, x, S_STABLE | S_NOMIRROR);
  if (!mp->S) {
    fprintf(stderr, "Out of memory for S table\n");
    exit(1);
  }
  /* every S_S(n_ik, t_ik) of this evaluation in one gather, then summed in the reference's order */
  if (mp->cnt && stb_S_batch(mp->S, mp->nn, mp->tt, mp->val, mp->cnt)) {
    fprintf(stderr, "samplea: table look-up failed: %s\n", stb_last_error());
    exit(1);
  }
  for (i = 0; i < mp->I; i++) {
    val += mp->T[i] * log(x) + lgamma(mp->T[i] + mp->bpar[i] / x) - lgamma(mp->bpar[i] / x);
    for (j = mp->first[i]; j < mp->first[i + 1]; j++) val += mp->val[j];
  }
  return val;
}

double samplea(double mya, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
               void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng,
               int loops, int verbose) {
  double inita[3] = {A_MIN, 1, A_MAX};
  int i, k;
  size_t total = 0, j = 0;
  ALData ald;
  inita[1] = mya;
  if (fabs(inita[1] - A_MAX) / A_MAX < 0.00001) inita[1] = A_MAX * 0.999 + A_MIN * 0.001;
  if (fabs(inita[1] - A_MIN) / A_MIN < 0.00001) inita[1] = A_MIN * 0.999 + A_MAX * 0.001;
  /* one MCMC step moves by less than SQUEEZEA */
  if (inita[1] - SQUEEZEA > A_MIN) inita[0] = inita[1] - SQUEEZEA;
  if (inita[1] + SQUEEZEA < A_MAX) inita[2] = inita[1] + SQUEEZEA;
  memset(&ald, 0, sizeof ald);
  ald.T = T;
  ald.I = I;
  ald.K = K;
  ald.bpar = bpar;
  ald.verbose = verbose;
  ald.maxt = 1;
  ald.maxn = 1;
  for (i = 0; i < I; i++) total += (size_t)K[i];
  ald.first = (size_t *)malloc(sizeof(size_t) * ((size_t)I + 1));
  ald.nn = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  ald.tt = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
  ald.val = (double *)malloc(sizeof(double) * (total ? total : 1));
  if (!ald.first || !ald.nn || !ald.tt || !ald.val) {
    fprintf(stderr, "Out of memory in samplea()\n");
    exit(1);
  }
  for (i = 0; i < I; i++) {
    ald.first[i] = j;
    for (k = 0; k < K[i]; k++) {
      scnt_int myn;
      stcnt_int myt;
      if (getval)
        getval(&myn, &myt, i, k);
      else {
        myn = n[i][k];
        myt = t[i][k];
      }
      if ((int)myt >= ald.maxt) ald.maxt = myt + 1;
      if ((int)myn >= ald.maxn) ald.maxn = myn + 1;
      if (myn > 1) {
        ald.nn[j] = myn;
        ald.tt[j] = myt;
        j++;
      }
    }
  }
  ald.first[I] = j;
  ald.cnt = j;
  if (g_sampler == STB_SAMPLER_ARS) { /* lib/samplea.c:209-215 */
    arms_simple(3, inita, inita + 2, aterms, &ald, 0, inita + 1, &mya);
    if (mya < inita[0] || mya > inita[2]) {
      fprintf(stderr, "Arms_simple(apar) returned value out of bounds\n");
      exit(1);
    }
  } else {
    /* the slice sampler may move anywhere in [inita[0], A_MAX] (lib/samplea.c:217-218) */
    inita[1] = A_MAX;
    if (SliceSimple(&mya, aterms, inita, rng, loops, &ald)) {
      fprintf(stderr, "SliceSimple error\n");
      exit(1);
    }
  }
  if (ald.S) S_free(ald.S);
  free(ald.first);
  free(ald.nn);
  free(ald.tt);
  free(ald.val);
  return mya;
}

/* ------------------------------------------------------------------------------------------ */
/* samplea2, lib/samplea.c:227-341 (SAMPLEA_M)                                                  */
/* ------------------------------------------------------------------------------------------ */
double logminus(double x, double y) { /* lib/samplea.c:229-239 */
  if (y >= x) return -HUGE_VAL;
  if (y - x < -80) return x - exp(y - x);
  return x + log(1 - exp(y - x));
}

static int g_partition_mode = STB_PARTITION_REFERENCE;
int stb_set_partition_mode(int mode) {
  int old = g_partition_mode;
  if (mode == STB_PARTITION_REFERENCE || mode == STB_PARTITION_EXACT) g_partition_mode = mode;
  return old;
}

typedef struct {
  int I;
  int *K;
  scnt_int *T;
  scnt_int **n;
  stcnt_int **t;
  void (*val)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k);
  double *bpar;
  stcnt_int *m; /* sampled table sizes, t-1 per node with 1 < t < n, in (i,k) order */
} AL2Data;

/* log p(a = x | table sizes) up to a constant, summed in the reference's order (lib/samplea.c:87-150) */
static double aterms2(double x, void *mydata) {
  AL2Data *mp = (AL2Data *)mydata;
  double val = 0;
  struct gcache_s lgp;
  stcnt_int *mm = mp->m;
  int i, k;
  if (x <= 0) {
    fprintf(stderr, "Illegal discount value in aterms2()\n");
    exit(1);
  }
  gcache_init(&lgp, 1 - x);
  for (i = 0; i < mp->I; i++) {
    val += mp->T[i] * log(x) + lgamma(mp->T[i] + mp->bpar[i] / x) - lgamma(mp->bpar[i] / x);
    for (k = 0; k < mp->K[i]; k++) {
      scnt_int n;
      stcnt_int t;
      if (mp->val)
        mp->val(&n, &t, i, k);
      else {
        n = mp->n[i][k];
        t = mp->t[i][k];
      }
      if (n == 0 || t == n) continue;
      if (t == 1)
        val += gcache_value(&lgp, n - 1);
      else {
        int l;
        for (l = t - 2; l >= 0; l--) {
          if (mm[l] > 1) val += gcache_value(&lgp, mm[l] - 1);
          n -= mm[l];
        }
        if (n > 0) val += gcache_value(&lgp, n - 1);
        mm += t - 1;
      }
    }
  }
  return val;
}

double samplea2(double mya, stable_t *S, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
                void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng,
                int loops, int verbose) {
  double inita[3] = {A_MIN, 1, A_MAX};
  int i, k;
  size_t n_m = 0, cnt = 0, j = 0, o = 0;
  const int exact = g_partition_mode == STB_PARTITION_EXACT;
  AL2Data ald;
  uint32_t *pn, *poff;
  uint16_t *pt;
  double *plogu;
  (void)verbose;
  inita[1] = mya;
  if (fabs(inita[1] - A_MAX) / A_MAX < 0.00001) inita[1] = A_MAX * 0.999 + A_MIN * 0.001;
  if (fabs(inita[1] - A_MIN) / A_MIN < 0.00001) inita[1] = A_MIN * 0.999 + A_MAX * 0.001;
  if (inita[1] - SQUEEZEA > A_MIN) inita[0] = inita[1] - SQUEEZEA;
  if (inita[1] + SQUEEZEA < A_MAX) inita[2] = inita[1] + SQUEEZEA;
  ald.T = T;
  ald.n = n;
  ald.t = t;
  ald.I = I;
  ald.K = K;
  ald.val = getval;
  ald.bpar = bpar;
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++)
      if (t[i][k] > 1 && t[i][k] < n[i][k]) {
        n_m += (size_t)t[i][k] - 1;
        cnt++;
      }
  ald.m = (stcnt_int *)malloc(sizeof(*ald.m) * (n_m ? n_m : 1));
  pn = (uint32_t *)malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
  poff = (uint32_t *)malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
  pt = (uint16_t *)malloc(sizeof(uint16_t) * (cnt ? cnt : 1));
  plogu = (double *)malloc(sizeof(double) * (exact ? (n_m ? n_m : 1) : (cnt ? cnt : 1)));
  if (!ald.m || !pn || !poff || !pt || !plogu) {
    fprintf(stderr, "Out of memory for samplea()\n");
    exit(1);
  }
  /* one uniform per node in (i,k) order (lib/samplea.c:293), then every node's walk in one kernel */
  for (i = 0; i < I; i++)
    for (k = 0; k < K[i]; k++)
      if (t[i][k] > 1 && t[i][k] < n[i][k]) {
        pn[j] = n[i][k];
        pt[j] = t[i][k];
        if (exact) { /* one uniform per round, M = t-1 .. 1 */
          int M;
          for (M = (int)t[i][k] - 1; M >= 1; M--) plogu[o + (size_t)M - 1] = log(rng_unit(rng));
        } else
          plogu[j] = log(rng_unit(rng));
        poff[j] = (uint32_t)o;
        o += (size_t)t[i][k] - 1;
        j++;
      }
  if (cnt && stb_partition_sample(S, mya, pn, pt, plogu, poff, cnt, ald.m, n_m, exact)) {
    fprintf(stderr, "samplea2: partition sampling failed: %s\n", stb_last_error());
    exit(1);
  }
  free(pn);
  free(poff);
  free(pt);
  free(plogu);
  if (g_sampler == STB_SAMPLER_ARS) { /* lib/samplea.c:322-328, with the data pointer the reference forgets */
    arms_simple(3, inita, inita + 2, aterms2, &ald, 0, inita + 1, &mya);
    if (mya < inita[0] || mya > inita[2]) {
      fprintf(stderr, "Arms_simple(apar) returned value out of bounds\n");
      exit(1);
    }
  } else {
    inita[1] = A_MAX;
    if (SliceSimple(&mya, aterms2, inita, rng, loops, &ald)) {
      fprintf(stderr, "SliceSimple error\n");
      exit(1);
    }
  }
  free(ald.m);
  return mya;
}
