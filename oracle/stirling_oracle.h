/*
 * stirling_oracle.h -- CPU restatement of the reference's table algorithm.
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product (libstb_b200) never links or calls this.
 */
#ifndef STIRLING_ORACLE_H
#define STIRLING_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* flag bits, same values as lib/stable.h:38-44 */
#define ORC_STABLE 1
#define ORC_UVTABLE 2
#define ORC_FLOAT 4
#define ORC_QUITONBOUND 16
#define ORC_ASYMPT 64

/* Dense tables: cell (n,m) at [(n-1)*ld + (m-1)], valid for 1<=m<=min(n,M). */
void orc_fill_S1(unsigned N, double a, double *S1);
void orc_fill_S(unsigned N, unsigned M, double a, double *S, size_t ld);
void orc_fill_V(unsigned N, unsigned M, double a, double *V, size_t ld);

/* a fixed-extent table object with the reference's look-up conventions (no growth) */
typedef struct orc_table {
  unsigned N, M, maxN, maxM;
  double a, lga;
  uint32_t flags;
  size_t ld;
  double *S, *V, *S1;
} orc_table;

orc_table *orc_make(unsigned N, unsigned M, unsigned maxN, unsigned maxM, double a, uint32_t flags);
void orc_free(orc_table *t);
double orc_S(const orc_table *t, unsigned n, unsigned m);
double orc_S1(const orc_table *t, unsigned n);
double orc_V(const orc_table *t, unsigned n, unsigned m);
double orc_U(const orc_table *t, unsigned n, unsigned m);
double orc_UV(const orc_table *t, unsigned n, unsigned m);
double orc_asympt(double a, unsigned n, unsigned m);
double orc_V_asympt(double a, unsigned n, unsigned m);

/* samplea2's partition sampler (lib/samplea.c:227-341): see stirling_oracle.c */
double orc_logminus(double x, double y);
void orc_partition_node(const orc_table *tb, double a, unsigned n, unsigned t, const double *logu, uint16_t *m,
                        int exact);
double orc_partition_logp(const orc_table *tb, double a, unsigned N, unsigned M, unsigned l);

/* the table-indicator Gibbs step of test/demo.c:405-434 over R restaurants: see stirling_oracle.c */
void orc_ti_gibbs(const orc_table *tb, double apar, double bpar, size_t R, const uint32_t *tok_off, const uint32_t *tok_dish,
                  const float *H, uint32_t D, const uint32_t *n, uint16_t *t, uint32_t *T, uint64_t *rng, int shared,
                  int sweeps);

/* number of stored cells with m>=2, SURVEY.md section 8 */
uint64_t orc_cells_S(uint64_t N, uint64_t M);
uint64_t orc_cells_V(uint64_t N, uint64_t M);

#ifdef __cplusplus
}
#endif
#endif
