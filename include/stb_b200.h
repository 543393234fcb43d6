/*
 * stb_b200.h -- batched / device-side extensions of the stable.h API.
 *
 * The reference's API is scalar host calls (lib/stable.h:128-190); these entry points are the
 * batched forms of the same look-ups, for callers that want thousands of cells at once
 * without walking the host mirror, plus a few introspection helpers used by the tests and the
 * benchmark.  Everything here is plain C ABI: pointers, sizes, no CUDA or torch types.
 */
#ifndef STB_B200_H
#define STB_B200_H

#include <stddef.h>
#include <stdint.h>

#include "stable.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * out[i] = S_S(sp, n[i], m[i])  (resp. S_V) for i < count, evaluated by one gather kernel.
 * HOST pointers; the table is first grown to cover the batch exactly as the scalar calls
 * would (lib/stable.c:950-965), pairs beyond maxN/maxM answer -inf (S) / 0 (V); the
 * S_ASYMPT and S_QUITONBOUND behaviours of the scalar calls are not applied.
 * Returns non-zero on error (no such table, CUDA failure).
 */
int stb_S_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
int stb_V_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
/* S_U / S_UV (lib/stable.c:875-897) from the V table in the same kernel: U = n - m a + 1/V (m == 1: n - a;
 * m == 0: NaN, the scalar call exits), UV = (n - m a) V + 1 (m == 1: -inf, m == n+1: 1, m == n: (n+1)/(n-1)) */
int stb_U_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
int stb_UV_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);

/*
 * The seat-partition sampler inside samplea2 (lib/samplea.c:290-321) for `count` nodes in one kernel, one
 * thread per node: node j has n[j] customers at t[j] tables (1 < t[j] < n[j]); its t[j]-1 sampled table
 * sizes are written at m_out[off[j] .. off[j]+t[j]-2], entry M-1 by round M = t-1 .. 1 (the last table's
 * size is what remains).  a: the discount the table was filled at.
 *   exact == 0 (STB_PARTITION_REFERENCE): the reference's arithmetic, operation by operation; logu[j] =
 *     log of node j's ONE uniform.  As written there the remainder starts at S(n,t) + log u while the
 *     terms are normalised by S(n,t), so beyond the smallest counts no walk stops early: the result is one
 *     table of n-t+1 customers and t-1 singletons, whatever u.  Mirrored for parity with -DSAMPLEA_M builds.
 *   exact != 0 (STB_PARTITION_EXACT): draws from P(l | N, M+1) = C(N-1,l-1) (1-a)_{l-1} S^{N-l}_M / S^N_{M+1},
 *     the size of the table holding a given customer when N customers sit at M+1 tables, round by round;
 *     logu has n_m entries, logu[off[j] + M-1] = log of the uniform of round M.
 * The table is grown to cover the nodes first; counts beyond its maximum size are an error.  HOST
 * pointers.  Returns non-zero on error.
 */
#define STB_PARTITION_REFERENCE 0
#define STB_PARTITION_EXACT 1
int stb_partition_sample(stable_t *sp, double a, const uint32_t *n, const uint16_t *t, const double *logu,
                         const uint32_t *off, size_t count, uint16_t *m_out, size_t n_m, int exact);
/* which of the two samplea2 uses (default STB_PARTITION_REFERENCE); returns the previous setting */
int stb_set_partition_mode(int mode);

/* same, n/m/out are DEVICE pointers on the table's device; no growth is attempted */
int stb_S_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
int stb_V_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
int stb_U_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);
int stb_UV_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count);

/*
 * Grow the filled extent to at least n<=N, m<=M (within maxN/maxM) -- the exported form of
 * the reference's static S_extend (lib/stable.c:564), same growth policy.  0 on success.
 */
int stb_extend(stable_t *sp, unsigned N, unsigned M);

/*
 * Bulk export: copy rows n0 .. n0+nrows-1 (1-based) of the S (which_V==0) or V table into
 * dst as doubles, row pitch = the slab's ld (see stb_device_table); dst holds nrows*ld
 * doubles.  Only cells with 1<=m<=min(n,usedM) are meaningful.  0 on success.
 */
int stb_read_rows(stable_t *sp, int which_V, unsigned n0, unsigned nrows, double *dst);

/*
 * Discount sweep -- the batched form of "S_remake(sp, a) then look up my statistics" that
 * samplea's log-posterior does once per evaluation (lib/samplea.c:57-82).  For every discount
 * a[j] a full N x M log S table is filled on the device (several tables side by side per
 * launch, streamed through a few resident slabs) and reduced to what the caller asked for:
 *   gather_out [na][npairs]  S_S(n[i], m[i]) at a[j]      (NULL: not wanted)
 *   sum_out    [na]          sum_i of those values         (NULL: not wanted)
 *   lastrow_out[na][M]       log S^N_m, m = 1..M           (NULL: not wanted)
 * S_S conventions apply to the pairs (n==m -> 0, m==1 -> log S^n_1, m==0 or n<m or beyond N/M
 * -> -HUGE_VAL).  All pointers are HOST pointers.  flags: 0 or S_FLOAT.  Returns non-zero on error.
 */
typedef struct stb_sweep stb_sweep_t;
stb_sweep_t *stb_sweep_create(unsigned N, unsigned M, uint32_t flags);
int stb_sweep_set_pairs(stb_sweep_t *w, const uint32_t *n, const uint32_t *m, size_t npairs);
int stb_sweep_run(stb_sweep_t *w, const double *a, size_t na, double *gather_out, double *sum_out,
                  double *lastrow_out);
/* device milliseconds of the most recent stb_sweep_run (CUDA events around its queue of fills and
 * reductions: the fills are > 99 % of it); tables a launch fills side by side (one CTA group each), and tables
 * per launch in all -- a launch fills several rounds of them back to back into as many resident slabs */
double stb_sweep_last_fill_ms(const stb_sweep_t *w);
int stb_sweep_tables_in_flight(const stb_sweep_t *w);
int stb_sweep_tables_per_launch(const stb_sweep_t *w);
void stb_sweep_free(stb_sweep_t *w);

/*
 * Several devices from ONE process (csrc/multi.c).  The tables of a sweep are independent units -- the
 * reference's unit of work is one S_remake per evaluation of samplea's log-posterior (lib/samplea.c:57-60) --
 * so table j goes to devices[j % ndev]; every device runs the single-device engine on its share at the
 * same time (one host thread per device for the duration of a call, no traffic between devices) and
 * writes its results into row j of the caller's HOST arrays.  devices == NULL or ndev <= 0: every visible
 * device.  A device may be named more than once (its shares then take turns on it).  The results equal
 * the single-device results bit for bit, whatever the device list.  Arguments and return values as the
 * single-device calls above; on failure stb_last_error() names the failing device.
 */
typedef struct stb_sweep_multi stb_sweep_multi_t;
stb_sweep_multi_t *stb_sweep_multi_create(const int *devices, int ndev, unsigned N, unsigned M, uint32_t flags);
int stb_sweep_multi_set_pairs(stb_sweep_multi_t *w, const uint32_t *n, const uint32_t *m, size_t npairs);
int stb_sweep_multi_run(stb_sweep_multi_t *w, const double *a, size_t na, double *gather_out, double *sum_out,
                        double *lastrow_out);
/* device milliseconds of the most recent run: the slowest device's (they work side by side), one device's */
double stb_sweep_multi_last_fill_ms(const stb_sweep_multi_t *w);
double stb_sweep_multi_device_ms(const stb_sweep_multi_t *w, int g);
int stb_sweep_multi_devices(const stb_sweep_multi_t *w);
void stb_sweep_multi_free(stb_sweep_multi_t *w);

/*
 * A device-side consumer of the V table: table-indicator Gibbs sweeps of a Pitman-Yor /
 * Dirichlet-process mixture over R restaurants, reading V^n_m straight from the table's slab in device memory
 * -- the per-token update of the reference's demo (test/demo.c:405-434), whose inner operation is one scalar
 * S_V look-up per token on the host.  Restaurant j serves D dishes: n[j*D + i] customers eat dish i at
 * t[j*D + i] tables (1 <= t <= n wherever n > 0), T[j] = sum_i t[j*D + i]; its tokens, in the order they are
 * visited, are the dishes tok_dish[tok_off[j] .. tok_off[j+1]).  For a token of dish i with n > 1 a sweep
 *   - removes its table indicator with probability (t - 1)/(n - 1)            (drawn only when t > 1),
 *   - adds one with probability one/(one + 1), one = (float)(H[i] (b + T a) t/(n - t + 1) V^n_{t+1}),
 * a = the table's discount, b = bpar, the arithmetic operation by operation as in the demo.  Restaurants are
 * independent given (a, b, H): with shared_stream == 0 each runs on its own 48-bit stream rng[j] (the state of a
 * stb_rng48_t, i.e. what seed48 would be handed) -- one thread per restaurant, `sweeps` sweeps per launch.
 * shared_stream != 0 is the demo's own schedule: ONE stream rng[0], restaurants visited in order, which
 * reproduces the reference's draws uniform for uniform (the parity mode; serial by construction).
 * t, T and rng are updated in place.  The table must have been made with S_UVTABLE; it is grown (up to its
 * maximum size) to cover the counts before the launch.  Returns 0, or non-zero with stb_last_error() set.
 */
int stb_ti_gibbs(stable_t *sp, double bpar, size_t R, const uint32_t *tok_off, const uint32_t *tok_dish, const float *H,
                 uint32_t D, const uint32_t *n, uint16_t *t, uint32_t *T, uint64_t *rng, int shared_stream, int sweeps);
/* device milliseconds of the most recent stb_ti_gibbs kernel on this table */
double stb_last_gibbs_ms(const stable_t *sp);

/*
 * Per-chain replacement for the C library's rand() / srand() (glibc's additive feedback generator,
 * csrc/rand31.h): stb_rand31_seed(g, s) puts g in the state srand(s) puts the global generator in,
 * stb_rand31_next(g) returns what rand() would return next.  The batched ARS samplers draw each
 * chain's uniforms from its own stb_rand31_t (the reference's ARMS uses rand(), lib/arms.c:913-918).
 */
typedef struct stb_rand31 {
  int32_t r[31];
  int32_t f, b;
} stb_rand31_t;
void stb_rand31_seed(stb_rand31_t *g, unsigned seed);
int stb_rand31_next(stb_rand31_t *g);

/*
 * arms_simple(3, &lo[c], &hi[c], myfunc, mydata, 0, ., &x[c]) (include/arms.h) for C independent
 * chains advanced in lock-step, chain c drawing its uniforms from rnd[c]: the driver of the batched
 * ARS samplers with a host log-density.  A chain whose sampler stops with an arms() error code keeps
 * x[c]; *nfailed (may be NULL) counts them.  Returns 0, or a negative value when memory runs out.
 */
int stb_arms_simple_batch(double *x, size_t C, const double *lo, const double *hi, stb_rand31_t *rnd,
                          double (*myfunc)(double x, void *mydata), void *mydata, size_t *nfailed);

/* frees device memory the batched samplers keep between calls (the sweep handle of the last
 * stb_samplea_batch: one launch's worth of table slabs) */
void stb_release_caches(void);

/* device milliseconds of the most recent fill (CUDA events around the kernel) */
double stb_last_fill_ms(const stable_t *sp);
/* device milliseconds of the most recent stb_partition_sample kernel on this table */
double stb_last_partition_ms(const stable_t *sp);
/* device address and row pitch (elements) of the S (which_V==0) or V slab; cell (n,m) at [(n-1)*ld+m-1] */
const void *stb_device_table(const stable_t *sp, int which_V, size_t *ld);
/* usable CUDA devices; 0 means every S_make will fail (there is no CPU path) */
int stb_device_count(void);
const char *stb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
