/*
 * shim_gcache.c -- TEST INFRASTRUCTURE.  Records the arguments of the reference's gcache_value calls.
 *
 * The reference's samplea2 (lib/samplea.c:244-341) keeps the table sizes it samples in a private
 * array; the only trace they leave is the sequence of gcache_value(&lgp, size-1) calls its aterms2
 * makes (lib/samplea.c:118-143).  Loaded with RTLD_GLOBAL before oracle/_ref/libstb_ref_slice_m.so,
 * this definition is the one the reference's calls bind to: it logs j and forwards to the
 * reference's own function (address handed over by the test).  Used only by
 * tests/ref_samplea2_probe.py, in a process of its own.
 */
#include <stddef.h>

static double (*real_fn)(void *, int);
static int *rec;
static size_t rec_cap, rec_cnt;

void shim_set(void *real, int *buf, size_t cap) {
  real_fn = (double (*)(void *, int))real;
  rec = buf;
  rec_cap = cap;
  rec_cnt = 0;
}
size_t shim_count(void) { return rec_cnt; }

double gcache_value(void *lpg, int j) {
  if (rec_cnt < rec_cap) rec[rec_cnt] = j;
  rec_cnt++;
  return real_fn(lpg, j);
}
