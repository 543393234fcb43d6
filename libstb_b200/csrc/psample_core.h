/*
 * psample_core.h -- the one sampler engine behind psample.h (internal).
 *
 * Every entry point of psample.h -- the scalar samplea / sampleb / samplea2 / SliceSimple of the
 * reference's API and the batched forms -- runs the same lock-step drivers (psample_batch.c): C chains,
 * each with its own random stream, advance round by round, and a round's log-posterior evaluations are
 * ONE call of an evaluator.  What differs is the evaluator's back end:
 *   STB_BACKEND_DEVICE  a discount sweep plus reduction kernels (thousands of chains per round);
 *   STB_BACKEND_HOST    the caller's thread with libm, summed in the reference's order, the table of an
 *                       evaluation refilled by the CUDA engine -- the scalar API: one chain whose stream is
 *                       glibc's own, taken over for the call and handed back (psample.c).
 */
#ifndef STB_PSAMPLE_CORE_H
#define STB_PSAMPLE_CORE_H

#include <stddef.h>
#include <stdint.h>

#include "psample.h"
#include "stb_b200.h"

/* values of a log-posterior at cnt points; chain[j] says which chain asks for x[j]; non-zero: failed */
typedef int (*stb_eval_fn)(void *ctx, const double *x, const int *chain, size_t cnt, double *out);

#define STB_BACKEND_DEVICE 0
#define STB_BACKEND_HOST 1

/* the uniforms of the adaptive rejection sampler: one stb_rand31_t per chain, or (streams == NULL) glibc's
 * rand(), which the reference's ARMS draws from (lib/arms.c:913-918) -- a single chain only */
typedef struct stb_ars_source {
  stb_rand31_t *streams;
} stb_ars_source;

/* SliceSimple (lib/sslice.c:33-80) for C chains: xp[c] start / result inside [lo[c], hi[c]], rng[c] the chain's
 * 48-bit stream.  depth: proposals evaluated ahead per chain and round; slots: see psample_batch.c.
 * Returns 0, 1 + the first failing chain, or a negative value (memory / evaluator). */
int stb_slice_lockstep(double *xp, size_t C, const double *lo, const double *hi, uint64_t *rng, int loops, stb_eval_fn eval,
                       void *ctx, stb_sample_stats *st, int depth, size_t slots);
/* arms_simple(3, lo, hi, ., ., 0, ., &xp[c]) (lib/arms.c:98-123) for C chains */
int stb_ars_lockstep(double *xp, size_t C, const double *lo, const double *hi, const stb_ars_source *src, stb_eval_fn eval,
                     void *ctx, stb_sample_stats *st, size_t *nfailed);

/* one update of the discount (lib/samplea.c:155-225) for C chains; ars == NULL: slice sampler on rng[] */
int stb_discount_step(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                      const double *bpar, int bpar_per_chain, uint64_t *rng, const stb_ars_source *ars, int loops,
                      stb_sample_stats *st, int backend, int verbose);
/* one update of the concentration (lib/sampleb.c:79-159) for C chains; rng[] also feeds the auxiliary Beta draws */
int stb_concentration_step(double *b, size_t C, int I, double shape, double scale, const scnt_int *N, const scnt_int *T,
                           const double *apar, uint64_t *rng, const stb_ars_source *ars, int loops, stb_sample_stats *st,
                           int backend, int verbose);
/* the table-free update of the discount (lib/samplea.c:227-341): seat partitions from the caller's table S (one
 * kernel over all nodes), then one sampler step over the partition's likelihood; one chain, host back end */
int stb_discount_step_partition(double *a, stable_t *S, int I, const int *K, const scnt_int *T, scnt_int **n, stcnt_int **t,
                                const double *bpar, uint64_t *rng, const stb_ars_source *ars, int loops, int exact,
                                int verbose);
/* the same for C chains, every chain against its own table out of a discount sweep (device back end);
 * *partition_ms (may be NULL): device time of the fills and the partition kernels */
int stb_discount_step_partition_batch(double *a, size_t C, int I, const int *K, const scnt_int *T, scnt_int **n,
                                      stcnt_int **t, const double *bpar, int bpar_per_chain, uint64_t *rng,
                                      const stb_ars_source *ars, int loops, int exact, stb_sample_stats *st,
                                      double *partition_ms);
int stb_get_partition_mode(void);

/* device steps of the batched partition update on a sweep handle (stable.c -> stb_cuda.cu) */
int stb_sweep_set_nodes(stb_sweep_t *w, const uint32_t *n, const uint16_t *t, const uint32_t *draw, size_t count,
                        const uint32_t *hbase, unsigned hbins);
int stb_sweep_partition(stb_sweep_t *w, const double *a, size_t na, const uint64_t *x0, int exact, uint32_t *hist_out,
                        float *ms);
int stb_sweep_hist_eval(stb_sweep_t *w, const double *x, const int *chain, size_t cnt, double *out, float *ms);

#endif
