/*
 * srng.h -- random number interface of the samplers (reference: lib/srng.h:15-34).
 *
 * Like the reference, the scalar samplers draw from glibc's GLOBAL drand48/lrand48 stream, so a
 * caller that seeds with srand48() gets the same draws from either library.  The distribution
 * functions below are this repo's own implementations (libstb_b200/csrc/rng48.h); they consume
 * the global stream exactly like their counterparts in lib/gslrandist.c:194-282.
 * The batched samplers of stb_b200.h carry one 48-bit state per chain instead.
 */
#ifndef STB_B200_SRNG_H
#define STB_B200_SRNG_H
#include <stdlib.h>
#include <time.h>
#ifdef __cplusplus
extern "C" {
#endif

double gsl_rng_gaussian_ziggurat(const double sigma);
double gsl_rng_beta(const double a, const double b);
double gsl_rng_gamma(const double a);

typedef void *rngp_t;

#define rng_seed(rng, seed) srand48(seed);
#define rng_time(rng, seed) \
  {                         \
    *(seed) = time(NULL);   \
    srand48(*(seed));       \
  }
#define rng_unit(rng) drand48()
#define rng_beta(rng, a, b) gsl_rng_beta(a, b)
#define rng_gamma(rng, a) gsl_rng_gamma(a)
#define rng_gaussian(rng, a) gsl_rng_gaussian_ziggurat(a)
#define rng_free(rng)

#ifdef __cplusplus
}
#endif
#endif
