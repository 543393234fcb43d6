/*
 * stable.h -- generalised second-order Stirling number tables, B200 (sm_100a) engine.
 *
 * Source-compatible replacement for the table API of wbuntine/libstb
 * (reference interface: lib/stable.h:38-44 flags, :128-190 prototypes, :192-196 ISFINITE).
 * Same function names, argument meaning, return conventions and flag values; the table
 * cells themselves are produced by CUDA kernels and live in HBM, with a host-side mirror
 * that keeps the scalar look-ups S_S/S_V/S_U/S_UV O(1).
 *
 * The struct below is NOT layout-compatible with the reference's (its fields were declared
 * private there, lib/stable.h:57-61, and no caller touches them); callers only ever hold a
 * `stable_t *` obtained from S_make().
 */
#ifndef STB_B200_STABLE_H
#define STB_B200_STABLE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Flag bits for S_make(); values identical to lib/stable.h:38-44.
 *   S_STABLE      maintain the log S table
 *   S_UVTABLE     maintain the V ratio table (U and UV are derived from it)
 *   S_FLOAT       store cells as float (all arithmetic still FP64)
 *   S_VERBOSE     one-line S_report() to stderr after each S_remake()
 *   S_QUITONBOUND exit(1) instead of returning log(0)/0 when maxN/maxM are exceeded
 *   S_THREADS     serialise table growth with a mutex so look-ups may run concurrently
 *   S_ASYMPT      past maxN answer from the closed-form asymptote
 */
#define S_STABLE 1
#define S_UVTABLE 2
#define S_FLOAT 4
#define S_VERBOSE 8
#define S_QUITONBOUND 16
#define S_THREADS 32
#define S_ASYMPT 64
/*
 * Extension bits (not in the reference; they sit above its flag range).
 *   S_MIRROR_ORDER  fill with the reference's own operation order (log-domain logadd for S,
 *                   ratio recursion for V; V is then bit-identical to the CPU library).
 *                   Default is the scaled linear-domain recurrence (same results to ~1e-14).
 *   S_NOMIRROR      never build the eager host mirror; scalar look-ups fetch row blocks on
 *                   demand (default: eager when the table is small, lazy when it is large).
 */
#define S_MIRROR_ORDER (1u << 16)
#define S_NOMIRROR (1u << 17)

/* kept so that code written against the reference's threaded build still compiles */
#ifdef H_THREADS
#define S_USE_THREADS
#endif

struct stb_table_impl; /* private: device buffers, host mirror, mutex */

typedef struct stable_s {
  /* inclusive bounds, fixed for the table's life */
  unsigned maxM, maxN;
  /* inclusive bounds currently filled; grow on demand up to the maxima */
  unsigned usedM, usedN;
  /* host cache of S1[n-1] = log S^n_1 and the number of slots it has */
  unsigned usedN1;
  double *S1;
  double lga; /* lgamma(1-a) */
  double a;   /* discount */
  uint32_t flags;
  uint32_t memalloced; /* bytes (host+device) as the reference counts them; wraps at 4 GiB */
  char *tag;
  struct stb_table_impl *impl;
} stable_t;

/*
 * Build tables for n<=initN, m<=initM (both raised to 10, capped by the maxima) at discount a.
 * Returns NULL on allocation/CUDA failure or when neither S_STABLE nor S_UVTABLE is set.
 * (replaces lib/stable.c:110-312)
 */
stable_t *S_make(unsigned initN, unsigned initM, unsigned maxN, unsigned maxM, double a,
                 uint32_t flags);
void S_tag(stable_t *S, char *tag);

/* Refill the current extent for a new discount; non-zero on error. (lib/stable.c:549-554) */
int S_remake(stable_t *sp, double a);

/* Release everything; NULL is accepted. (lib/stable.c:980-1023) */
void S_free(stable_t *sp);

/*
 * log S^n_{m,a}.  n==m -> 0;  m==1 -> S_S1;  n<m or m==0 -> -HUGE_VAL;  grows the table
 * when (n,m) is inside the maxima, else asymptote (S_ASYMPT, n>maxN) / exit (S_QUITONBOUND)
 * / -HUGE_VAL.  (lib/stable.c:941-974)
 */
double S_S(stable_t *sp, unsigned n, unsigned m);

/* log S^n_{1,a} = lgamma(n-a) - lgamma(1-a), cached.  (lib/stable.c:822-873) */
double S_S1(stable_t *sp, unsigned n);

/* closed-form large-n approximation of log S^n_{m,a}.  (lib/stable.c:1057-1084) */
double S_asympt(stable_t *sp, unsigned n, unsigned m);

/* U^n_m = S^{n+1}_m / S^n_m = n - m a + 1/V^n_m.  (lib/stable.c:875-883) */
double S_U(stable_t *sp, unsigned n, unsigned m);
/* U^n_m * V^n_m with one look-up.  (lib/stable.c:885-897) */
double S_UV(stable_t *sp, unsigned n, unsigned m);
/* V^n_m = S^n_m / S^n_{m-1}, m>=2; 0 when out of bounds.  (lib/stable.c:900-939) */
double S_V(stable_t *sp, unsigned n, unsigned m);

/* one line of statistics.  (lib/stable.c:1025-1055) */
void S_report(stable_t *sp, FILE *fp);

#ifdef isfinite
#define ISFINITE(x) isfinite(x)
#else
#define ISFINITE(x) finite(x)
#endif

#ifdef __cplusplus
}
#endif
#endif
