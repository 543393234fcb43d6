/*
 * psample.h -- posterior samplers for the Pitman-Yor discount a and concentration b.
 *
 * Source-compatible with the reference's lib/psample.h (B_MIN/B_MAX :58-59, scnt_int :64,
 * stcnt_int :68, sampleb :79-84, A_MIN/A_MAX/SQUEEZEA :89-94, samplea :104-110, SliceSimple
 * :45-51) in its SLICE-SAMPLER configuration (PSAMPLE_ARS undefined): samplea/sampleb run
 * SliceSimple, not ARS.  samplea's log-posterior refills a Stirling table per evaluation
 * (lib/samplea.c:57-60); here that refill is the CUDA table engine of stable.h.
 * The batched forms (thousands of independent chains in lock-step) are in stb_b200.h.
 */
#ifndef STB_B200_PSAMPLE_H
#define STB_B200_PSAMPLE_H

#include <stdint.h>

#include "srng.h"
#include "stable.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Shrinking-interval slice sampler for a unimodal log-posterior post(x, pars).  *xp: start
 * (inside bounds) and result.  Returns non-zero on error (start outside the bounds, or 200
 * proposals without acceptance).  (lib/sslice.c:33-80)
 */
int SliceSimple(double *xp, double (*post)(double, void *), double *bounds, rngp_t rng, int loops, void *pars);

#define B_MIN 0.01
#define B_MAX 2000

typedef uint32_t scnt_int;  /* counts */
typedef uint16_t stcnt_int; /* table counts */

/*
 * One MCMC update of the concentration b given the discount apar, gamma(shape, scale) prior,
 * I restaurants with customer totals N[] and table totals T[].  (lib/sampleb.c:79-159)
 */
double sampleb(double b_in, int I, double shape, double scale, scnt_int *N, scnt_int *T, double apar, rngp_t rng,
               int loops, int verbose);

#define A_MIN 0.01
#define A_MAX 0.98
#define SQUEEZEA 0.2

/*
 * Which sampler samplea / sampleb run.  The reference fixes this at compile time
 * (lib/psample.h:37 PSAMPLE_ARS: ARS is its shipped default, the slice sampler the alternative);
 * here it is a run-time switch, default STB_SAMPLER_SLICE.  With STB_SAMPLER_ARS the two behave
 * like the reference's default build: arms_simple(3, ...) on [a-SQUEEZEA, a+SQUEEZEA] resp.
 * [B_MIN, B_MAX], uniforms from rand() (lib/samplea.c:209-215, lib/sampleb.c:127-140).
 * Returns the previous setting.  Process-wide, like the reference's generator state.
 */
#define STB_SAMPLER_SLICE 0
#define STB_SAMPLER_ARS 1
int stb_set_sampler(int which);

/*
 * One MCMC update of the discount a (uniform prior on [A_MIN, A_MAX], moves squeezed to
 * +-SQUEEZEA).  n[i][k], t[i][k]: customers / tables of dish k in restaurant i (k < K[i]), or
 * the callback getval(&n,&t,i,k) when non-NULL; T[i] = sum_k t[i][k]; bpar[i]: concentration of
 * restaurant i.  Builds and frees its own Stirling table.  (lib/samplea.c:155-225)
 */
double samplea(double apar, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
               void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng,
               int loops, int verbose);

/*
 * The table-free discount update (lib/samplea.c:227-341; compiled in the reference only with
 * SAMPLEA_M, lib/psample.h:30).  S: the caller's table at the current discount mya.  For every node
 * with 1 < t < n the sizes of its t tables are sampled given (n, t) -- one uniform per node, in
 * (i,k) order, from rng -- through ratios of S_S values (here: one kernel over all nodes,
 * stb_partition_sample); given the sizes the posterior of a needs no Stirling numbers, and one
 * slice-sampling (or ARS) step over it returns the new discount.  n and t must be arrays (the
 * partition step does not use getval, like the reference).
 * Deviation: in the reference's ARS build this function hands arms_simple a NULL data pointer
 * (lib/samplea.c:323-324) and would crash; here the ARS mode passes the statistics.
 */
double samplea2(double mya, stable_t *S, int I, int *K, scnt_int *T, scnt_int **n, stcnt_int **t,
                void (*getval)(scnt_int *n, stcnt_int *t, unsigned i, unsigned k), double *bpar, rngp_t rng,
                int loops, int verbose);
/* log(exp(x) - exp(y)), -inf when y >= x (lib/samplea.c:229-239) */
double logminus(double x, double y);

/* ------------------------------------------------------------------------------------------ */
/* batched samplers: C independent chains in lock-step, one batched evaluation per round        */
/* ------------------------------------------------------------------------------------------ */

/*
 * Per-chain random streams: glibc's 48-bit generator with one state per chain.
 * stb_rng48_state(seed) is the state srand48(seed) sets; the draw functions advance *state exactly
 * like drand48 / lrand48 / the distributions of srng.h advance the global state.
 */
uint64_t stb_rng48_state(long seed);
double stb_rng48_drand(uint64_t *state);
long stb_rng48_lrand48(uint64_t *state);
double stb_rng48_gaussian(uint64_t *state, double sigma);
double stb_rng48_gamma(uint64_t *state, double a);
double stb_rng48_beta(uint64_t *state, double a, double b);

typedef struct stb_sample_stats {
  uint64_t evals;  /* log-posterior evaluations over all chains (speculative slice proposals included) */
  uint64_t rounds; /* lock-step rounds (batched evaluations) */
  double eval_ms;  /* device milliseconds spent in the evaluations */
  /* optional trace: chain c's evaluated points / values in order at [c*trace_cap + k], count in trace_n[c] */
  double *trace_x, *trace_v;
  uint32_t *trace_n;
  uint32_t trace_cap;
} stb_sample_stats;

/*
 * samplea for C chains over the SAME statistics: a[c] in/out, rng[c] in/out.  bpar: [I], or
 * [C][I] when bpar_per_chain.  Every round fills one Stirling table per active chain at that
 * chain's proposed discount (a discount sweep) and reduces it against the statistics.
 * Returns 0; 1+c when chain c failed like the scalar sampler would exit (start outside
 * [max(A_MIN, a-SQUEEZEA), A_MAX], 200 rejected proposals); negative on device / memory errors.
 */
int stb_samplea_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                      const double *bpar, int bpar_per_chain, uint64_t *rng, int loops, stb_sample_stats *st);

/*
 * samplea2 (lib/samplea.c:227-341, the reference's -DSAMPLEA_M build) for C chains over the same statistics:
 * chain c samples the seat partition of every node with 1 < t < n against its OWN Stirling table at a[c] (one
 * table per chain out of a discount sweep; the scalar samplea2 is handed the caller's table), drawing its
 * uniforms from rng[c] in the scalar call's order, and then makes one slice step on the partition's
 * likelihood -- a device reduction over the chain's histogram of sampled table sizes.  The partition mode is
 * stb_set_partition_mode's.  Chain c's new discount and stream equal the scalar samplea2's on a table at a[c]
 * with the stream rng[c].  Return values as stb_samplea_batch.
 */
int stb_samplea2_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                       const double *bpar, int bpar_per_chain, uint64_t *rng, int loops, stb_sample_stats *st);

/* sampleb for C chains: b[c] in/out, apar[c] the chain's discount, rng[c] in/out */
int stb_sampleb_batch(double *b, size_t C, int I, double shape, double scale, const scnt_int *N, const scnt_int *T,
                      const double *apar, uint64_t *rng, int loops, stb_sample_stats *st);

/*
 * The same over several devices of this process (csrc/multi.c): chain c runs on devices[c % ndev], every
 * device advances its chains with the single-device call at the same time, no traffic between devices;
 * a[] / b[] / rng[] are updated in place.  devices == NULL or ndev <= 0: every visible device.  Chain c's
 * draw and stream do not depend on the device list.  In *st the evaluations add up; rounds and eval_ms
 * are the slowest device's; traces are not kept.
 */
int stb_samplea_batch_multi(const int *devices, int ndev, double *a, size_t C, int I, const int *K, const scnt_int *T,
                            scnt_int **n, stcnt_int **t, const double *bpar, int bpar_per_chain, uint64_t *rng, int loops,
                            stb_sample_stats *st);
int stb_sampleb_batch_multi(const int *devices, int ndev, double *b, size_t C, int I, double shape, double scale,
                            const scnt_int *N, const scnt_int *T, const double *apar, uint64_t *rng, int loops,
                            stb_sample_stats *st);

/*
 * The same in the reference's DEFAULT (ARS) configuration: every chain runs arms_simple(3, ...)
 * (lib/samplea.c:209-215 on [a - SQUEEZEA, a + SQUEEZEA] clipped to [A_MIN, A_MAX];
 * lib/sampleb.c:127-140 on [B_MIN, B_MAX]) as a resumable machine, the chains advance in lock-step
 * and each round's log-posterior evaluations are one device batch.  rnd[c]: the chain's rand()
 * stream (stb_rand31_t, include/stb_b200.h), in/out; rng[c] (sampleb): its 48-bit stream for the
 * auxiliary Beta draws.  Returns 0, or 1 + the index of the first failing chain.
 */
struct stb_rand31;
int stb_samplea_batch_ars(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                          const double *bpar, int bpar_per_chain, struct stb_rand31 *rnd, stb_sample_stats *st);
int stb_sampleb_batch_ars(double *b, size_t C, int I, double shape, double scale, const scnt_int *N,
                          const scnt_int *T, const double *apar, uint64_t *rng, struct stb_rand31 *rnd,
                          stb_sample_stats *st);

#ifdef __cplusplus
}
#endif
#endif
