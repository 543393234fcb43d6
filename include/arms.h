/*
 * arms.h -- adaptive rejection (Metropolis) sampling on an interval (reference interface:
 * lib/arms.h:3-16; Gilks, Best & Tan 1995).  Same prototypes and return codes as the reference:
 * 0 success; 1001 fewer than 3 starting points; 1002 more starting points than envelope room;
 * 1003 starting points not inside the bounds; 1004 starting points not ascending; 1005 a centile
 * outside [0,100]; 1006 out of memory; 1007 previous iterate outside the bounds; 1008 negative
 * convexity adjustment; 2000 log-density not concave and no Metropolis step; 2001 more than 100
 * proposals.  Uniforms come from rand(): seed with srand().  (libstb_b200/csrc/ars.c)
 */
#ifndef STB_B200_ARMS_H
#define STB_B200_ARMS_H
#ifdef __cplusplus
extern "C" {
#endif

/* one draw; ninit equally spaced starting points inside (*xl, *xr), envelope of at most 100 knots */
int arms_simple(int ninit, double *xl, double *xr, double (*myfunc)(double x, void *mydata), void *mydata,
                int dometrop, double *xprev, double *xsamp);

int arms(double *xinit, int ninit, double *xl, double *xr, double (*myfunc)(double x, void *mydata), void *mydata,
         double *convex, int npoint, int dometrop, double *xprev, double *xsamp, int nsamp, double *qcent,
         double *xcent, int ncent, int *neval);

double expshift(double y, double y0);

#define YCEIL 50. /* largest shifted log-density exponentiated */

#ifdef __cplusplus
}
#endif
#endif
