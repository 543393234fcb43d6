/*
 * bench_dropin.c -- the DEFAULT drop-in path of stable.h, timed: the pattern of an
 * unmodified libstb caller (test/demo.c:333-334, 427-428, 467-488): S_make with default flags, a run of
 * S_remake calls (a new discount after every samplea), and scalar S_V / S_S look-ups per token.
 * Compiled twice by oracle/build_ref.sh from this one source: against the reference's lib/stable.h and
 * library (oracle/_ref/dropin_bench_ref) and against this repo's include/ and libstb_b200.so
 * (oracle/_ref/dropin_bench_b200).  Measurement tool (bench.py extras); not part of the product.
 *
 *   usage: dropin_bench N M remakes lookups [touched_rows]
 * prints one JSON object: seconds of S_make, of the S_remake calls, of the look-ups, and a checksum.
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "stable.h"

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
  const unsigned N = argc > 1 ? (unsigned)atoi(argv[1]) : 10000, M = argc > 2 ? (unsigned)atoi(argv[2]) : 1000;
  const int remakes = argc > 3 ? atoi(argv[3]) : 20;
  const long lookups = argc > 4 ? atol(argv[4]) : 10000000L;
  const unsigned rows = argc > 5 ? (unsigned)atoi(argv[5]) : N; /* look-ups fall into the last `rows` rows */
  double t0, t_make, t_remake, t_first, t_look, sum = 0;
  unsigned long long s = 88172645463325252ULL;
  stable_t *sp;
  long i;
  int r;
  t0 = now();
  sp = S_make(N, M, N, M, 0.5, S_STABLE | S_UVTABLE);
  t_make = now() - t0;
  if (!sp) {
    fprintf(stderr, "S_make failed\n");
    return 1;
  }
  t0 = now();
  for (r = 0; r < remakes; r++) {
    S_remake(sp, 0.3 + 0.02 * r);
    sum += S_V(sp, N - 1, M / 2); /* a caller looks at the table it has just made */
  }
  t_remake = now() - t0;
  t0 = now();
  sum += S_S(sp, N - 3, M - 3);
  t_first = now() - t0;
  t0 = now();
  for (i = 0; i < lookups; i++) {
    unsigned n, m;
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17; /* xorshift64 */
    n = N - 2 - (unsigned)(s % (rows > 4 ? rows - 4 : 1));
    m = 2 + (unsigned)((s >> 32) % (M - 4));
    if (m > n) m = n;
    sum += S_V(sp, n, m);
  }
  t_look = now() - t0;
  printf("{\"N\": %u, \"M\": %u, \"make_s\": %.6f, \"remakes\": %d, \"remake_s_each\": %.6f, \"first_lookup_s\": %.6f, "
         "\"lookups\": %ld, \"lookup_s\": %.6f, \"lookup_ns_each\": %.2f, \"checksum\": %.10g}\n",
         N, M, t_make, remakes, t_remake / (remakes > 0 ? remakes : 1), t_first, lookups, t_look,
         1e9 * t_look / (lookups > 0 ? lookups : 1), sum);
  S_free(sp);
  return 0;
}
