/*
 * multi.c -- several devices from ONE process (stb_b200.h, "several devices" section).
 *
 * The two sharded workloads of the engine are sets of independent units: the tables of a discount
 * sweep (one full refill per discount, the unit of work of samplea's log-posterior,
 * lib/samplea.c:57-60) and the chains of the batched samplers (one samplea / sampleb call per chain,
 * lib/samplea.c:155-225, lib/sampleb.c:79-159).  A single table does not shard: its rows are sequential
 * and every cell needs its left neighbour.
 *
 * Unit j goes to devices[j % ndev].  Every device runs the single-device engine on its share -- one host
 * thread per device for the duration of a call, each with its own CUDA stream and resident slabs; there
 * is no traffic between devices while they work -- and writes its results straight into the caller's
 * host arrays at the units' own positions (the "final gather" is the engine's own device-to-host copy,
 * a few KB per unit).  A unit's result does not depend on which device made it or on what shared its
 * launch, so any device list gives the same bits as one device.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psample.h"
#include "stb_b200.h"
#include "stb_cuda.h"

#define MULTI_MAX_DEV 64

/* the device list of a call: NULL / ndev <= 0 means every visible device */
static int device_list(const int *devices, int ndev, int *out) {
  const int have = stb_cuda_device_count();
  int g;
  if (have <= 0) {
    stb_cuda_set_error("no CUDA device", 0);
    return -1;
  }
  if (!devices || ndev <= 0) {
    ndev = have > MULTI_MAX_DEV ? MULTI_MAX_DEV : have;
    for (g = 0; g < ndev; g++) out[g] = g;
    return ndev;
  }
  if (ndev > MULTI_MAX_DEV) {
    stb_cuda_set_error("too many devices in the list", 0);
    return -1;
  }
  for (g = 0; g < ndev; g++) {
    if (devices[g] < 0 || devices[g] >= have) {
      stb_cuda_set_error("device list names a device that does not exist", 0);
      return -1;
    }
    out[g] = devices[g];
  }
  return ndev;
}

/* run fn(job g) for g < n, one thread each (job 0 on the calling thread); every job records rc / err */
typedef struct {
  int device, rc;
  char err[256];
  void *arg;
  void (*fn)(void *arg, int *rc);
} Job;

static void *job_main(void *p) {
  Job *j = (Job *)p;
  const int prev = stb_cuda_current_device();
  j->rc = stb_cuda_use_device(j->device);
  if (!j->rc) j->fn(j->arg, &j->rc);
  if (j->rc) snprintf(j->err, sizeof j->err, "device %d: %s", j->device, stb_cuda_last_error());
  if (prev >= 0) stb_cuda_use_device(prev);
  return NULL;
}

static int run_jobs(Job *jobs, int n) {
  pthread_t th[MULTI_MAX_DEV];
  int started[MULTI_MAX_DEV];
  int g, rc = 0;
  for (g = 1; g < n; g++) {
    started[g] = pthread_create(&th[g], NULL, job_main, &jobs[g]) == 0;
    if (!started[g]) job_main(&jobs[g]); /* no thread to be had: do it here, after the others were started */
  }
  job_main(&jobs[0]);
  for (g = 1; g < n; g++)
    if (started[g]) pthread_join(th[g], NULL);
  for (g = 0; g < n; g++)
    if (jobs[g].rc && !rc) {
      rc = jobs[g].rc;
      stb_cuda_set_error(jobs[g].err, 0); /* the failing worker's message, in the caller's thread */
    }
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* discount sweep                                                                              */
/* ------------------------------------------------------------------------------------------ */
struct stb_sweep_multi {
  int ndev, dev[MULTI_MAX_DEV];
  stb_sweep_dev_t *w[MULTI_MAX_DEV];
  unsigned N, M;
  float ms[MULTI_MAX_DEV];
  double last_ms;
};

typedef struct {
  stb_sweep_multi_t *mw;
  int g;
  /* create */
  int is_float;
  /* set_pairs */
  const uint32_t *n, *m;
  size_t npairs;
  /* run */
  const double *a;
  size_t na;
  double *gather_out, *sum_out, *lastrow_out;
} SweepArg;

static void sweep_create_job(void *p, int *rc) {
  SweepArg *s = (SweepArg *)p;
  s->mw->w[s->g] = stb_cuda_sweep_create(s->mw->N, s->mw->M, s->is_float);
  *rc = s->mw->w[s->g] ? 0 : -1;
}
static void sweep_pairs_job(void *p, int *rc) {
  SweepArg *s = (SweepArg *)p;
  *rc = stb_cuda_sweep_set_pairs(s->mw->w[s->g], s->n, s->m, s->npairs);
}
static void sweep_run_job(void *p, int *rc) {
  SweepArg *s = (SweepArg *)p;
  const size_t G = (size_t)s->mw->ndev, g = (size_t)s->g;
  const size_t mine = s->na > g ? (s->na - g + G - 1) / G : 0; /* units g, g+G, g+2G, ... */
  s->mw->ms[g] = 0.f;
  *rc = stb_cuda_sweep_run_dealt(s->mw->w[g], s->a, mine, g, G, s->gather_out, s->sum_out, s->lastrow_out, &s->mw->ms[g]);
}

static int sweep_jobs(stb_sweep_multi_t *mw, SweepArg *proto, void (*fn)(void *, int *)) {
  Job jobs[MULTI_MAX_DEV];
  SweepArg args[MULTI_MAX_DEV];
  int g;
  for (g = 0; g < mw->ndev; g++) {
    args[g] = *proto;
    args[g].mw = mw;
    args[g].g = g;
    memset(&jobs[g], 0, sizeof jobs[g]);
    jobs[g].device = mw->dev[g];
    jobs[g].arg = &args[g];
    jobs[g].fn = fn;
  }
  return run_jobs(jobs, mw->ndev);
}

stb_sweep_multi_t *stb_sweep_multi_create(const int *devices, int ndev, unsigned N, unsigned M, uint32_t flags) {
  stb_sweep_multi_t *mw = (stb_sweep_multi_t *)calloc(1, sizeof *mw);
  SweepArg proto;
  if (!mw) return NULL;
  mw->ndev = device_list(devices, ndev, mw->dev);
  if (mw->ndev <= 0) {
    free(mw);
    return NULL;
  }
  mw->N = N;
  mw->M = M;
  memset(&proto, 0, sizeof proto);
  proto.is_float = (flags & S_FLOAT) != 0;
  if (sweep_jobs(mw, &proto, sweep_create_job)) {
    stb_sweep_multi_free(mw);
    return NULL;
  }
  return mw;
}

int stb_sweep_multi_set_pairs(stb_sweep_multi_t *mw, const uint32_t *n, const uint32_t *m, size_t npairs) {
  SweepArg proto;
  if (!mw) return 1;
  memset(&proto, 0, sizeof proto);
  proto.n = n;
  proto.m = m;
  proto.npairs = npairs;
  return sweep_jobs(mw, &proto, sweep_pairs_job);
}

int stb_sweep_multi_run(stb_sweep_multi_t *mw, const double *a, size_t na, double *gather_out, double *sum_out,
                        double *lastrow_out) {
  SweepArg proto;
  int g, rc;
  if (!mw) return 1;
  memset(&proto, 0, sizeof proto);
  proto.a = a;
  proto.na = na;
  proto.gather_out = gather_out;
  proto.sum_out = sum_out;
  proto.lastrow_out = lastrow_out;
  rc = sweep_jobs(mw, &proto, sweep_run_job);
  mw->last_ms = 0;
  for (g = 0; g < mw->ndev; g++)
    if (mw->ms[g] > mw->last_ms) mw->last_ms = mw->ms[g];
  return rc;
}

double stb_sweep_multi_last_fill_ms(const stb_sweep_multi_t *mw) { return mw ? mw->last_ms : 0; }
int stb_sweep_multi_devices(const stb_sweep_multi_t *mw) { return mw ? mw->ndev : 0; }
double stb_sweep_multi_device_ms(const stb_sweep_multi_t *mw, int g) { return (mw && g >= 0 && g < mw->ndev) ? mw->ms[g] : 0; }

void stb_sweep_multi_free(stb_sweep_multi_t *mw) {
  int g;
  if (!mw) return;
  for (g = 0; g < mw->ndev; g++) stb_cuda_sweep_destroy(mw->w[g]); /* runs on the handle's device */
  free(mw);
}

/* ------------------------------------------------------------------------------------------ */
/* batched chains                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int g, G, which; /* which: 0 samplea, 1 sampleb */
  double *x;       /* a[] / b[], all chains */
  size_t C;
  int I;
  const int *K;
  const scnt_int *T, *N;
  scnt_int **n;
  stcnt_int **t;
  const double *bpar;
  int bpar_per_chain;
  const double *apar;
  double shape, scale;
  uint64_t *rng;
  int loops;
  stb_sample_stats st;
  int want_stats;
} ChainArg;

static void chain_job(void *p, int *rc) {
  ChainArg *c = (ChainArg *)p;
  const size_t G = (size_t)c->G, g = (size_t)c->g;
  const size_t mine = c->C > g ? (c->C - g + G - 1) / G : 0;
  double *x = NULL, *aux = NULL;
  uint64_t *rng = NULL;
  size_t k;
  *rc = 0;
  if (!mine) return;
  x = (double *)malloc(sizeof(double) * mine);
  rng = (uint64_t *)malloc(sizeof(uint64_t) * mine);
  /* per-chain inputs that travel with the chain: bpar rows (samplea), the discount (sampleb) */
  if (c->which == 0 && c->bpar_per_chain) aux = (double *)malloc(sizeof(double) * mine * (size_t)c->I);
  if (c->which == 1) aux = (double *)malloc(sizeof(double) * mine);
  if (!x || !rng || ((c->which == 1 || c->bpar_per_chain) && !aux)) {
    stb_cuda_set_error("out of host memory", 0);
    *rc = -1;
    goto done;
  }
  for (k = 0; k < mine; k++) {
    const size_t ch = g + k * G;
    x[k] = c->x[ch];
    rng[k] = c->rng[ch];
    if (c->which == 0 && c->bpar_per_chain) memcpy(aux + k * (size_t)c->I, c->bpar + ch * (size_t)c->I, sizeof(double) * (size_t)c->I);
    if (c->which == 1) aux[k] = c->apar[ch];
  }
  memset(&c->st, 0, sizeof c->st);
  if (c->which == 0)
    *rc = stb_samplea_batch(x, mine, c->I, c->K, c->T, c->n, c->t, c->bpar_per_chain ? aux : c->bpar, c->bpar_per_chain, rng,
                            c->loops, c->want_stats ? &c->st : NULL);
  else
    *rc = stb_sampleb_batch(x, mine, c->I, c->shape, c->scale, c->N, c->T, aux, rng, c->loops, c->want_stats ? &c->st : NULL);
  if (*rc > 0) *rc = 1 + (int)(g + (size_t)(*rc - 1) * G); /* the failing chain, in the caller's numbering */
  for (k = 0; k < mine; k++) { /* a failed call leaves its chains as the single-device call would */
    const size_t ch = g + k * G;
    c->x[ch] = x[k];
    c->rng[ch] = rng[k];
  }
done:
  free(x);
  free(rng);
  free(aux);
}

static int chain_jobs(const int *devices, int ndev, ChainArg *proto, stb_sample_stats *st) {
  Job jobs[MULTI_MAX_DEV];
  ChainArg args[MULTI_MAX_DEV];
  int dev[MULTI_MAX_DEV];
  int g, rc, G = device_list(devices, ndev, dev);
  if (G <= 0) return -1;
  if ((size_t)G > proto->C && proto->C) G = (int)proto->C;
  for (g = 0; g < G; g++) {
    args[g] = *proto;
    args[g].g = g;
    args[g].G = G;
    args[g].want_stats = st != NULL;
    memset(&jobs[g], 0, sizeof jobs[g]);
    jobs[g].device = dev[g];
    jobs[g].arg = &args[g];
    jobs[g].fn = chain_job;
  }
  rc = run_jobs(jobs, G);
  if (st) { /* evaluations add up; the devices work side by side: rounds and device time are the slowest one's */
    for (g = 0; g < G; g++) {
      st->evals += args[g].st.evals;
      if (args[g].st.rounds > st->rounds) st->rounds = args[g].st.rounds;
      if (args[g].st.eval_ms > st->eval_ms) st->eval_ms = args[g].st.eval_ms;
    }
  }
  return rc;
}

int stb_samplea_batch_multi(const int *devices, int ndev, double *a, size_t C, int I, const int *K, const scnt_int *T,
                            scnt_int **n, stcnt_int **t, const double *bpar, int bpar_per_chain, uint64_t *rng, int loops,
                            stb_sample_stats *st) {
  ChainArg proto;
  if (!C) return 0;
  memset(&proto, 0, sizeof proto);
  proto.which = 0;
  proto.x = a;
  proto.C = C;
  proto.I = I;
  proto.K = K;
  proto.T = T;
  proto.n = n;
  proto.t = t;
  proto.bpar = bpar;
  proto.bpar_per_chain = bpar_per_chain;
  proto.rng = rng;
  proto.loops = loops;
  return chain_jobs(devices, ndev, &proto, st);
}

int stb_sampleb_batch_multi(const int *devices, int ndev, double *b, size_t C, int I, double shape, double scale,
                            const scnt_int *N, const scnt_int *T, const double *apar, uint64_t *rng, int loops,
                            stb_sample_stats *st) {
  ChainArg proto;
  if (!C) return 0;
  memset(&proto, 0, sizeof proto);
  proto.which = 1;
  proto.x = b;
  proto.C = C;
  proto.I = I;
  proto.N = N;
  proto.T = T;
  proto.apar = apar;
  proto.shape = shape;
  proto.scale = scale;
  proto.rng = rng;
  proto.loops = loops;
  return chain_jobs(devices, ndev, &proto, st);
}
