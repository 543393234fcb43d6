/*
 * digamma.h -- digamma / trigamma / inverse digamma (reference: lib/digamma.h:36-71,
 * lib/digammainv.c:27).  FP64; the same code runs on the device for the batched samplers
 * (libstb_b200/csrc/specfun.h).
 */
#ifndef STB_B200_DIGAMMA_H
#define STB_B200_DIGAMMA_H
#ifdef __cplusplus
extern "C" {
#endif

double digammaRN(double x);   /* Neal's series, the reference's default digamma (lib/digamma.c:31-48) */
double MLdigamma(double x);   /* x > 0 */
double MLtrigamma(double x);  /* x > 0 */
double MLtetragamma(double x);  /* second derivative of psi, x > 0 */
double MLpentagamma(double x);  /* third derivative of psi, x > 0 */
double MLpsigamma(double x, double deriv); /* psi^(n)(x), n = round(deriv) >= 0 (lib/polygamma.c:502-523) */
double digammaInv(double x);  /* Minka start + 5 Newton steps (lib/digammainv.c:27-38) */

#define digamma(x) MLdigamma(x)
#define trigamma(x) MLtrigamma(x)
#define tetragamma(x) MLtetragamma(x)
#define pentagamma(x) MLpentagamma(x)

#ifdef __cplusplus
}
#endif
#endif
