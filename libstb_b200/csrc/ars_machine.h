/*
 * ars_machine.h -- the adaptive-rejection sampler of ars.c as a resumable machine (internal): it
 * runs until it needs the log-density at a point, hands the point out and continues when the
 * value is fed back.  arms()/arms_simple() drive one machine with a callback; the batched
 * samplers drive one machine per chain in lock-step and evaluate a round's points in one batch.
 */
#ifndef STB_ARS_MACHINE_H
#define STB_ARS_MACHINE_H

typedef struct stb_ars stb_ars_t;

#define STB_ARS_NEED 1 /* a point to evaluate is in *x_out */
#define STB_ARS_DONE 0 /* all requested points sampled */

stb_ars_t *stb_ars_new(int npoint);
void stb_ars_free(stb_ars_t *e);
/* unif: uniform(0,1) source (NULL: rand(), as the reference); returns STB_ARS_NEED or an arms() error code */
int stb_ars_begin(stb_ars_t *e, const double *xinit, int ninit, double xl, double xr, double convex, int dometrop,
                  double xprev, double *xsamp, int nsamp, double (*unif)(void *), void *ustate, double *x_out);
/* feed the value at the point last handed out: STB_ARS_NEED, STB_ARS_DONE or an arms() error code */
int stb_ars_feed(stb_ars_t *e, double y, double *x_out);
int stb_ars_neval(const stb_ars_t *e);
double stb_ars_centile(stb_ars_t *e, double q);

#endif
