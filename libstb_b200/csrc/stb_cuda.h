/*
 * stb_cuda.h -- the thin C-ABI between the C host layer (stable.c, samplers.c) and the
 * CUDA translation units.  Plain pointers and sizes only.  Every function returns 0 on
 * success and a non-zero cudaError_t-like code on failure unless stated otherwise; the C
 * layer maps failures onto the reference's conventions (NULL from S_make, yaps_quit on
 * growth failure -- lib/stable.c:115-116, 924-925, 963-964).
 *
 * There is deliberately no CPU implementation behind any of these: without a CUDA device
 * the calls fail and the public API reports the failure.
 */
#ifndef STB_CUDA_H
#define STB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stb_dev stb_dev_t;

/* which table */
#define STB_TAB_S 0
#define STB_TAB_V 1

/* fill algorithms */
#define STB_FILL_LINEAR 0 /* scaled linear-domain recurrence, strip-pipelined (default) */
#define STB_FILL_MIRROR 1 /* reference operation order, log domain (S) / ratio recursion (V) */

/* number of usable CUDA devices (0 when none / driver missing) */
int stb_cuda_device_count(void);
/* last error text of the CALLING THREAD for diagnostics (thread-local storage) */
const char *stb_cuda_last_error(void);
/* the calling thread's current device (-1: none), and cudaSetDevice for it (0 on success): handles are
 * created on the current device; every call on a handle runs on the handle's device and restores the
 * caller's.  Used by the per-device worker threads of the multi-device entry points (multi.c). */
int stb_cuda_current_device(void);
int stb_cuda_use_device(int dev);

/*
 * Create the device side of one table object on the current device.
 *   want_S / want_V : which tables to hold;  is_float : store cells as float.
 */
stb_dev_t *stb_cuda_table_create(int want_S, int want_V, int is_float);
void stb_cuda_table_destroy(stb_dev_t *d);

/*
 * Make room for rows n=1..N and columns m=1..M.  Cells are laid out dense row-major: cell
 * (n,m) at [(n-1)*ld + (m-1)], ld = M rounded up to 32 elements so that every row starts on a
 * 128-byte boundary.  With keep!=0 the cells already filled survive a re-allocation.
 */
int stb_cuda_table_reserve(stb_dev_t *d, unsigned N, unsigned M, int keep);
/* room for a table that grows to N x M and is then refilled as a whole: nothing is kept, the old slabs are
 * freed first, the capacity at least doubles (capped at maxN x maxM) */
int stb_cuda_table_grow(stb_dev_t *d, unsigned N, unsigned M, unsigned maxN, unsigned maxM);
size_t stb_cuda_table_ld(const stb_dev_t *d);
size_t stb_cuda_table_bytes(const stb_dev_t *d); /* device bytes currently held */

/*
 * Fill rows 1..N, columns 1..M at discount a with the chosen algorithm, then wait for
 * completion.  Cells with n<=startN && m<=startM are declared already valid for this a
 * (startN==0: nothing is) -- the counterpart of S_remake_part's start arguments,
 * lib/stable.c:321-323.  Column 1 of the S table is log S^n_1; `s1_host` (N doubles, may be
 * NULL) receives it.  With STB_FILL_MIRROR the caller instead passes the host-computed S1
 * running sum in `s1_host` (uploaded and used as the m=1 column, lib/stable.c:338-348).
 */
int stb_cuda_fill(stb_dev_t *d, double a, unsigned startN, unsigned startM, unsigned N, unsigned M,
                  int algo, double *s1_host);
/* the m = 1 column (log S^n_1, n = 1..N) the most recent strip-kernel fill left on the device */
int stb_cuda_read_s1(stb_dev_t *d, unsigned N, double *dst);
/* milliseconds the device spent in the most recent stb_cuda_fill (CUDA events) */
float stb_cuda_last_fill_ms(const stb_dev_t *d);
/* milliseconds the device spent in the most recent stb_cuda_partition kernel */
float stb_cuda_last_partition_ms(const stb_dev_t *d);

/*
 * Copy rows [row0, row0+nrows) (0-based row index = n-1) of one table to host memory as
 * doubles (float tables are widened on the host side of the copy).  dst holds nrows*ld doubles.
 */
int stb_cuda_read_rows(stb_dev_t *d, int which, unsigned row0, unsigned nrows, double *dst);

/*
 * out[i] = look-up of (n[i], m[i]) in table `which`, evaluated on the device with the scalar
 * API's in-range conventions (S: n==m -> 0, m==0 or n<m -> -inf; V: m<2 or n<m -> 0); pairs
 * beyond usedN/usedM answer -inf / 0 (the C layer grows the table first when it may).
 * n, m, out are HOST pointers when on_device==0 (staged) or DEVICE pointers otherwise.
 */
int stb_cuda_gather(stb_dev_t *d, int which, double a, unsigned usedN, unsigned usedM, const uint32_t *n,
                    const uint32_t *m, double *out, size_t count, int on_device);
/*
 * samplea2's seat-partition sampler (lib/samplea.c:290-321) for `count` nodes at once: node j has
 * n[j] customers at t[j] tables (1 < t < n, inside the filled table), logu[j] = log of its uniform
 * draw (exact != 0: logu[off[j] + M-1] for round M, n_m entries), and writes its t[j]-1 table sizes at
 * m_out[off[j] ..].  Host pointers.
 */
int stb_cuda_partition(stb_dev_t *d, double a, const uint32_t *n, const uint16_t *t, const double *logu,
                       const uint32_t *off, size_t count, uint16_t *m_out, size_t n_m, int exact);
/* `which` values beyond the two tables: ratios computed from V in the same kernel (a: the discount) */
#define STB_GATHER_U 2
#define STB_GATHER_UV 3

/*
 * Discount sweep: fill MANY tables of one extent (N x M, log S only), one per discount, and keep
 * from each only what a caller such as samplea's log-posterior needs (lib/samplea.c:57-82: a
 * full refill at the new discount, then a sum of S_S(n_ik, t_ik) over the statistics) -- the
 * values at a fixed set of (n,m) pairs, their sum, and the last row.  The tables themselves are
 * streamed through a few resident slabs (as many as one launch fills side by side) and
 * overwritten by the next wave.
 */
typedef struct stb_sweep_dev stb_sweep_dev_t;
stb_sweep_dev_t *stb_cuda_sweep_create(unsigned N, unsigned M, int is_float);
void stb_cuda_sweep_destroy(stb_sweep_dev_t *w);
/* the look-up set, HOST arrays; S_S conventions (n==m -> 0, m==0 or n<m -> -inf, beyond N/M -> -inf) */
int stb_cuda_sweep_set_pairs(stb_sweep_dev_t *w, const uint32_t *n, const uint32_t *m, size_t npairs);
/*
 * Run the sweep over a[0..na).  Any of the outputs may be NULL:
 *   gather_out [na][npairs]  value at each pair;   sum_out [na]  their sum (fixed-order tree);
 *   lastrow_out [na][M]  row N of each table.
 * HOST pointers.  *fill_ms receives the device time spent in the fill kernels.
 */
int stb_cuda_sweep_run(stb_sweep_dev_t *w, const double *a, size_t na, double *gather_out, double *sum_out,
                       double *lastrow_out, float *fill_ms);
/* the share of a sweep dealt to this device: unit k < na of the run is the caller's unit first + k*stride
 * (its discount a[first + k*stride], its results in that row of the caller's arrays) */
int stb_cuda_sweep_run_dealt(stb_sweep_dev_t *w, const double *a, size_t na, size_t first, size_t stride,
                             double *gather_out, double *sum_out, double *lastrow_out, float *fill_ms);
/* tables one launch fills side by side for this extent on this device */
int stb_cuda_sweep_tables_in_flight(const stb_sweep_dev_t *w);
int stb_cuda_sweep_tables_per_launch(const stb_sweep_dev_t *w);
/* samplea2 over many chains (stb_cuda.cu): nodes of the partition step, the step itself against one table per
 * chain, and the evaluation of the likelihood's table terms from the per-chain histograms of sampled sizes */
int stb_cuda_sweep_set_nodes(stb_sweep_dev_t *w, const uint32_t *n, const uint16_t *t, const uint32_t *draw, size_t count,
                             const uint32_t *hbase, unsigned hbins);
int stb_cuda_sweep_partition(stb_sweep_dev_t *w, const double *a, size_t na, const uint64_t *x0, int exact,
                             uint32_t *hist_out, float *ms);
int stb_cuda_sweep_hist_eval(stb_sweep_dev_t *w, const double *x, const int *chain, size_t cnt, double *out, float *ms);
uint64_t stb_cuda_lcg48_jump(uint64_t x, uint64_t k);

/*
 * Per-restaurant statistics on the device and the I-term reductions of the batched samplers
 * (psample_cuda.cu).  T, N: uint32[I] (either may be NULL when unused); bpar: bpar_elems doubles,
 * [I] shared by all chains or [C][I]; max_evals: most evaluations one call will carry.
 */
typedef struct stb_pstat_dev stb_pstat_dev_t;
stb_pstat_dev_t *stb_cuda_pstat_create(int I, const uint32_t *T, const uint32_t *N, const double *bpar,
                                       size_t bpar_elems, size_t max_evals);
void stb_cuda_pstat_destroy(stb_pstat_dev_t *p); /* parks the context for the next create of the same shape */
void stb_cuda_pstat_purge(void);                  /* frees the parked context */
/* out[j] = sum_i [T_i log x_j + lgamma(T_i + b_i/x_j) - lgamma(b_i/x_j)], b row = chain[j] if per chain */
int stb_cuda_pstat_aterms_lg(stb_pstat_dev_t *p, const double *x, const int *chain, size_t cnt, int bpar_per_chain,
                             double *out, float *ms);
/* digamma_sum==0: out[j] = -Q_j x_j + (shape-1) log x_j + sum_i [lgamma(T_i + x_j/a_j) - lgamma(x_j/a_j)];
 * digamma_sum!=0: out[j] = sum_i digamma(T_i + x_j/a_j) */
int stb_cuda_pstat_bterms(stb_pstat_dev_t *p, const double *x, const double *Q, const double *apar, double shape,
                          size_t cnt, int digamma_sum, double *out, float *ms);
/* Q[c] = 1/scale - sum_i log Beta(b_in[c], N_i) on chain c's own 48-bit stream rng[c] (in/out) */
int stb_cuda_pstat_betaQ(stb_pstat_dev_t *p, const double *b_in, uint64_t *rng, size_t C, double scale, double *Q,
                         float *ms);
void stb_cuda_set_error(const char *what, int code);

/*
 * Table-indicator Gibbs sweeps over R restaurants reading V^n_m from the device table (gibbs_cuda.cu; the
 * per-token update of test/demo.c:405-434).  Host arrays in, t / T / rng updated in place; *ms = device time
 * of the kernel.  Returns 0 or an error code (text in stb_cuda_last_error()).
 */
int stb_cuda_ti_gibbs(stb_dev_t *d, unsigned usedN, unsigned usedM, double apar, double bpar, size_t R,
                      const uint32_t *tok_off, const uint32_t *tok_dish, const float *H, uint32_t D, const uint32_t *n,
                      uint16_t *t, uint32_t *T, uint64_t *rng, int shared_stream, int sweeps, float *ms);
int stb_cuda_table_device(const stb_dev_t *d);
int stb_cuda_table_is_float(const stb_dev_t *d);

/* raw device pointer of a table (for the batched samplers and for tests) */
void *stb_cuda_table_ptr(stb_dev_t *d, int which);

/* pinned host memory for mirrors */
void *stb_cuda_host_alloc(size_t bytes);
void stb_cuda_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
