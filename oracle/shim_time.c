/*
 * shim_time.c -- TEST INFRASTRUCTURE.  A fixed clock for the reference's test programs.
 *
 * test/demo.c and test/check.c reseed drand48 from time(NULL) before their Gibbs loops
 * (rng_time, lib/srng.h:29; test/demo.c:345, test/check.c:634), so two runs never agree.
 * LD_PRELOADed, this time() answers a constant: the programs become deterministic and the same
 * source linked with the reference library and with libstb_b200 can be compared line by line
 * (tests/test_dropin_gpu.py, tests/golden/make_golden_programs.py).
 */
#include <stdlib.h>
#include <time.h>

time_t time(time_t *t) {
  const char *s = getenv("STB_FAKE_TIME");
  time_t v = s ? (time_t)atol(s) : (time_t)1700000000;
  if (t) *t = v;
  return v;
}
