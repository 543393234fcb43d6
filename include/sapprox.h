/*
 * sapprox.h -- closed forms of log S^n_m for m <= 4 and their derivative in the discount
 * (reference interface: lib/sapprox.h:24-29, implementation lib/sapprox.c:28-114).
 *
 * Same names, argument types (note: the discount is a FLOAT, and the reference forms n - j*a in
 * single precision before calling lgamma, which is what limits these to ~1e-7 relative) and
 * conventions: 0 for n == m, -HUGE_VAL for n < m or m > 4.  Built like the reference's polygamma
 * configuration (lib/digamma.h:25 LS_NOPOLYGAMMA undefined): for a < 0.001 the m = 2..4 forms
 * switch to the digamma/trigamma/tetragamma expansions.
 */
#ifndef STB_B200_SAPPROX_H
#define STB_B200_SAPPROX_H
#ifdef __cplusplus
extern "C" {
#endif

double S_approx(int n, int m, float a);    /* lib/sapprox.c:33-77 */
double S_approx_da(int n, int m, float a); /* lib/sapprox.c:82-114 */

#ifdef __cplusplus
}
#endif
#endif
