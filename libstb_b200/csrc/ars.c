/*
 * ars.c -- adaptive rejection (Metropolis) sampling from a log-density on an interval:
 * arms_simple / arms of include/arms.h (reference interface: lib/arms.h:3-12; the reference's
 * implementation, lib/arms.c:98-919, is Gilks' ARMS -- Gilks, Best & Tan, Appl. Statist. 44 (1995)).
 * The reference's DEFAULT sampler configuration (lib/psample.h:37 PSAMPLE_ARS) drives samplea /
 * sampleb with it (lib/samplea.c:209-215, lib/sampleb.c:127-140); stb_set_sampler() selects it here.
 *
 * Restated from the method, not translated: the envelope is ONE sorted array instead of a linked
 * pool.  The structure the method maintains makes that natural -- envelope knots alternate
 *     bound, evaluated point, chord intersection, evaluated point, ..., evaluated point, bound
 * so knot j is an evaluated point of the log-density iff j is odd, "the evaluated point beyond my
 * neighbour" is j +- 2 / j +- 3, and incorporating a new point is inserting two entries.  The
 * arithmetic of each formula (chord intersections with their YEPS guards, the piecewise
 * exponential integral and its inversion, the shifted exp/log) is kept operation for operation,
 * and the uniforms are drawn in the same order from the same generator (rand(), lib/arms.c:913-918),
 * so under the same srand() seed the draws are bit-identical to the reference's
 * (tests/test_ars_cpu.py).  Host C: control flow only; the densities it calls (aterms: one table
 * refill per evaluation) are where the device work is.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arms.h"
#include "ars_machine.h"
#include "yaps.h"

#define X_EPS 0.00001 /* a new point is kept this (relative) distance away from its neighbours */
#define Y_EPS 0.1     /* below this rise a piece is integrated as a straight line */
#define EY_EPS 0.001  /* relative exp(y) difference below which a piece counts as flat */

typedef struct {
  double x, y; /* knot */
  double ey;   /* exp(y - ymax + YCEIL) */
  double cum;  /* integral of the exponentiated envelope up to x */
} knot_t;

/* the sampled (not yet incorporated) point: lies on the piece (right-1, right) */
typedef struct {
  double x, y, ey;
  int right;
  int evaluated;
} trial_t;

/*
 * The sampler is a RESUMABLE machine: it runs until it needs the log-density at a point, hands
 * the point out (ARS_NEED) and continues when the value is fed back.  The scalar entry points
 * below drive it with the caller's callback; the batched samplers (psample_batch.c) drive
 * thousands of machines in lock-step and evaluate each round's points in one device batch.
 * One code path for the arithmetic either way.
 */
enum { W_INIT = 0, W_YPREV, W_TRIAL, W_ADJUST, W_IDLE };

struct stb_ars {
  knot_t *k; /* sorted left to right; odd indices are evaluated points */
  int n, cap;
  double ymax;
  double convex;
  int neval;
  int metro;
  double xprev, yprev;
  double (*unif)(void *);
  void *ustate;
  /* progress */
  int waiting;  /* what the value being waited for is */
  int init_i, ninit;
  trial_t p;
  double u_y;   /* the rejection threshold log(u * envelope) of the current trial */
  int iq;       /* index of the point being incorporated */
  int tries, got, nsamp;
  double *xsamp;
  double xl, xr;
};

double expshift(double y, double y0) { /* exponentiate y shifted by y0 without overflow */
  return (y - y0 > -2.0 * YCEIL) ? exp(y - y0 + YCEIL) : 0.0;
}
static double log_unshift(double ey, double y0) { return log(ey) + y0 - YCEIL; }

static double uniform_rand(void *unused) {
  (void)unused;
  return ((double)rand() + 0.5) / 2147483648.0;
}

/*
 * Place knot j (even: a bound or a chord intersection) where the chords through the evaluated
 * points on its two sides meet.  Returns 1 when the log-density is found non-concave and no
 * Metropolis step is available to pay for it.
 */
static int place_intersection(stb_ars_t *e, int j) {
  knot_t *k = e->k;
  const int has_l = j >= 3, has_r = j + 3 <= e->n - 1, has_across = j >= 1 && j + 1 <= e->n - 1;
  double gl = 0, gr = 0, gacross = 0, dl = 0, dr = 0;
  if (j & 1) yaps_quit("ars: knot %d is not an intersection\n", j);
  if (has_l) gl = (k[j - 1].y - k[j - 3].y) / (k[j - 1].x - k[j - 3].x);
  if (has_r) gr = (k[j + 1].y - k[j + 3].y) / (k[j + 1].x - k[j + 3].x);
  if (has_across) gacross = (k[j + 1].y - k[j - 1].y) / (k[j + 1].x - k[j - 1].x);
  if (has_across && has_l && gl < gacross) { /* convex on the left */
    if (!e->metro) return 1;
    gl = gl + (1.0 + e->convex) * (gacross - gl);
  }
  if (has_across && has_r && gr > gacross) { /* convex on the right */
    if (!e->metro) return 1;
    gr = gr + (1.0 + e->convex) * (gacross - gr);
  }
  if (has_l && has_across) {
    dr = (gl - gacross) * (k[j + 1].x - k[j - 1].x);
    if (dr < Y_EPS) dr = Y_EPS;
  }
  if (has_r && has_across) {
    dl = (gacross - gr) * (k[j + 1].x - k[j - 1].x);
    if (dl < Y_EPS) dl = Y_EPS;
  }
  if (has_l && has_r && has_across) {
    k[j].x = (dl * k[j + 1].x + dr * k[j - 1].x) / (dl + dr);
    k[j].y = (dl * k[j + 1].y + dr * k[j - 1].y + dl * dr) / (dl + dr);
  } else if (has_l && has_across) { /* second knot from the right */
    k[j].x = k[j + 1].x;
    k[j].y = k[j + 1].y + dr;
  } else if (has_r && has_across) { /* second knot from the left */
    k[j].x = k[j - 1].x;
    k[j].y = k[j - 1].y + dl;
  } else if (has_l) { /* right bound: extend the last chord */
    k[j].y = k[j - 1].y + gl * (k[j].x - k[j - 1].x);
  } else if (has_r) { /* left bound */
    k[j].y = k[j + 1].y - gr * (k[j + 1].x - k[j].x);
  } else
    yaps_quit("ars: no chord on either side of knot %d\n", j);
  if ((j >= 1 && k[j].x < k[j - 1].x) || (j + 1 <= e->n - 1 && k[j].x > k[j + 1].x))
    yaps_quit("ars: intersection outside its interval (imprecision)\n");
  return 0;
}

/* exponentiate the envelope (shifted by its maximum) and integrate it piece by piece */
static void integrate(stb_ars_t *e) {
  knot_t *k = e->k;
  int j;
  e->ymax = k[0].y;
  for (j = 1; j < e->n; j++)
    if (k[j].y > e->ymax) e->ymax = k[j].y;
  for (j = 0; j < e->n; j++) k[j].ey = expshift(k[j].y, e->ymax);
  k[0].cum = 0.;
  for (j = 1; j < e->n; j++) {
    double a;
    if (k[j - 1].x == k[j].x)
      a = 0.;
    else if (fabs(k[j].y - k[j - 1].y) < Y_EPS)
      a = 0.5 * (k[j].ey + k[j - 1].ey) * (k[j].x - k[j - 1].x);
    else
      a = ((k[j].ey - k[j - 1].ey) / (k[j].y - k[j - 1].y)) * (k[j].x - k[j - 1].x);
    k[j].cum = k[j - 1].cum + a;
  }
}

/* the point of the envelope at cumulative probability prob */
static void invert_cdf(stb_ars_t *e, double prob, trial_t *p) {
  const knot_t *k = e->k;
  double xl = 0, xr = 0;
  int r = e->n - 1;
  const double u = prob * k[r].cum;
  while (k[r - 1].cum > u) r--;
  p->right = r;
  p->evaluated = 0;
  {
    const double frac = (u - k[r - 1].cum) / (k[r].cum - k[r - 1].cum);
    if (k[r - 1].x == k[r].x) {
      p->x = k[r].x;
      p->y = k[r].y;
      p->ey = k[r].ey;
    } else {
      const double yl = k[r - 1].y, yr = k[r].y, eyl = k[r - 1].ey, eyr = k[r].ey;
      xl = k[r - 1].x;
      xr = k[r].x;
      if (fabs(yr - yl) < Y_EPS) { /* the piece was integrated as a straight line */
        if (fabs(eyr - eyl) > EY_EPS * fabs(eyr + eyl))
          p->x = xl + ((xr - xl) / (eyr - eyl)) * (-eyl + sqrt((1. - frac) * eyl * eyl + frac * eyr * eyr));
        else
          p->x = xl + (xr - xl) * frac;
        p->ey = ((p->x - xl) / (xr - xl)) * (eyr - eyl) + eyl;
        p->y = log_unshift(p->ey, e->ymax);
      } else { /* exponential piece */
        p->x = xl + ((xr - xl) / (yr - yl)) * (-yl + log_unshift(((1. - frac) * eyl + frac * eyr), e->ymax));
        p->y = ((p->x - xl) / (xr - xl)) * (yr - yl) + yl;
        p->ey = expshift(p->y, e->ymax);
      }
    }
  }
  /* (for a zero-length piece xl = xr = 0 here, as in the reference: the guard then tests against 0) */
  if (p->x < xl || p->x > xr) yaps_quit("ars: sampled point outside its piece (imprecision)\n");
}

/* ---- the machine ------------------------------------------------------------------------------ */
#define ARS_NEED STB_ARS_NEED
#define ARS_DONE STB_ARS_DONE
#define ARS_GO_ON (-100) /* internal: the proposal is settled, carry on */
enum { IQ_FULL = -1, IQ_METROPOLIS = -2 }; /* e->iq when no point is being incorporated */

/*
 * Second half of incorporating the evaluated trial point (after its position was possibly moved
 * off its neighbours and re-evaluated): re-place the four intersections its chords touch,
 * re-integrate, then the rejection test.  Returns 1 accept, 0 reject, -1 envelope violation.
 */
static int incorporate_finish(stb_ars_t *e) {
  const int iq = e->iq;
  if (place_intersection(e, iq - 1)) return -1;
  if (place_intersection(e, iq + 1)) return -1;
  if (iq - 2 >= 0 && place_intersection(e, iq - 3)) return -1;
  if (iq + 2 <= e->n - 1 && place_intersection(e, iq + 3)) return -1;
  integrate(e);
  return e->u_y >= e->p.y ? 0 : 1;
}

/* one proposal settled: count it; returns ARS_DONE, an error code, or ARS_GO_ON */
static int proposal_outcome(stb_ars_t *e, int verdict) {
  if (verdict == 1)
    e->xsamp[e->got++] = e->p.x;
  else if (verdict != 0)
    return 2000;
  if (++e->tries > 100) return 2001; /* the reference gives up after 100 proposals (lib/arms.c:227-233) */
  return e->got < e->nsamp ? ARS_GO_ON : ARS_DONE;
}

/* run proposals until a log-density value is needed or sampling ends */
static int run_proposals(stb_ars_t *e, double *x_out) {
  for (;;) {
    const knot_t *k;
    trial_t *p = &e->p;
    int r, rc;
    invert_cdf(e, e->unif(e->ustate), p);
    k = e->k;
    r = p->right;
    e->u_y = log_unshift(e->unif(e->ustate) * p->ey, e->ymax);
    if (!e->metro && r - 2 >= 0 && r + 1 <= e->n - 1) { /* both ends of the piece have a neighbour beyond */
      const int sl = ((r - 1) & 1) ? r - 1 : r - 2, sr = (r & 1) ? r : r + 1; /* evaluated points around */
      const double ysq = (k[sr].y * (p->x - k[sl].x) + k[sl].y * (k[sr].x - p->x)) / (k[sr].x - k[sl].x);
      if (e->u_y <= ysq) { /* accepted by the squeeze: no evaluation */
        rc = proposal_outcome(e, 1);
        if (rc != ARS_GO_ON) return rc;
        continue;
      }
    }
    e->waiting = W_TRIAL;
    *x_out = p->x;
    return ARS_NEED;
  }
}

/* the value of the trial point has arrived */
static int trial_value(stb_ars_t *e, double ynew, double *x_out) {
  trial_t *p = &e->p;
  knot_t *k = e->k;
  if (!e->metro || e->u_y >= ynew) {
    int at = p->right, iq, lo, hi;
    p->y = ynew;
    p->ey = expshift(p->y, e->ymax);
    p->evaluated = 1;
    if (e->n > e->cap - 2) { /* no room in the envelope: the point is not incorporated */
      e->iq = IQ_FULL;
      return ARS_GO_ON;
    }
    /* two new knots: the point and one more intersection */
    memmove(k + at + 2, k + at, (size_t)(e->n - at) * sizeof *k);
    e->n += 2;
    /* the piece's left end is an evaluated point (odd index): new intersection first, then the point */
    iq = ((at - 1) & 1) ? at + 1 : at;
    e->iq = iq;
    k[iq].x = p->x;
    k[iq].y = p->y;
    /* keep the new point off its neighbours (the evaluated points two knots away, or the bounds) */
    lo = iq - 2 >= 0 ? iq - 2 : iq - 1;
    hi = iq + 2 <= e->n - 1 ? iq + 2 : iq + 1;
    if (k[iq].x < (1. - X_EPS) * k[lo].x + X_EPS * k[hi].x) {
      k[iq].x = (1. - X_EPS) * k[lo].x + X_EPS * k[hi].x;
      e->waiting = W_ADJUST;
      *x_out = k[iq].x;
      return ARS_NEED;
    } else if (k[iq].x > X_EPS * k[lo].x + (1. - X_EPS) * k[hi].x) {
      k[iq].x = X_EPS * k[lo].x + (1. - X_EPS) * k[hi].x;
      e->waiting = W_ADJUST;
      *x_out = k[iq].x;
      return ARS_NEED;
    }
    return ARS_GO_ON;
  }
  /* Metropolis step against the previous iterate */
  {
    int l = 0;
    double w, zold, znew, yold = e->yprev, u;
    while (k[l + 1].x < e->xprev) l++;
    w = (e->xprev - k[l].x) / (k[l + 1].x - k[l].x);
    zold = k[l].y + w * (k[l + 1].y - k[l].y);
    znew = p->y;
    if (yold < zold) zold = yold;
    if (ynew < znew) znew = ynew;
    w = ynew - znew - yold + zold;
    if (w > 0.0) w = 0.0;
    w = (w > -YCEIL) ? exp(w) : 0.0;
    u = e->unif(e->ustate);
    if (u > w) { /* stay */
      p->x = e->xprev;
      p->y = e->yprev;
      p->ey = expshift(p->y, e->ymax);
      p->evaluated = 1;
      p->right = l + 1;
    } else {
      e->xprev = p->x;
      e->yprev = ynew;
    }
  }
  e->iq = IQ_METROPOLIS; /* settled by the Metropolis step: accepted */
  return ARS_GO_ON;
}

stb_ars_t *stb_ars_new(int npoint) {
  stb_ars_t *e = (stb_ars_t *)calloc(1, sizeof *e);
  if (!e) return NULL;
  e->k = (knot_t *)malloc((size_t)(npoint > 0 ? npoint : 1) * sizeof(knot_t));
  if (!e->k) {
    free(e);
    return NULL;
  }
  e->cap = npoint;
  return e;
}

void stb_ars_free(stb_ars_t *e) {
  if (!e) return;
  free(e->k);
  free(e);
}

int stb_ars_neval(const stb_ars_t *e) { return e->neval; }

/*
 * Start sampling nsamp points.  Argument checks as the reference makes them, in its order
 * (lib/arms.c:286-318).  Returns ARS_NEED with the first point to evaluate in *x_out, or a code.
 */
int stb_ars_begin(stb_ars_t *e, const double *xinit, int ninit, double xl, double xr, double convex, int dometrop,
                  double xprev, double *xsamp, int nsamp, double (*unif)(void *), void *ustate, double *x_out) {
  int i;
  const int n0 = 2 * ninit + 1;
  if (ninit < 3) return 1001;
  if (e->cap < n0) return 1002;
  if (xinit[0] <= xl || xinit[ninit - 1] >= xr) return 1003;
  for (i = 1; i < ninit; i++)
    if (xinit[i] <= xinit[i - 1]) return 1004;
  if (convex < 0.0) return 1008;
  e->n = n0;
  e->convex = convex;
  e->neval = 0;
  e->metro = dometrop;
  e->xprev = xprev;
  e->unif = unif ? unif : uniform_rand;
  e->ustate = ustate;
  e->xl = xl;
  e->xr = xr;
  e->xsamp = xsamp;
  e->nsamp = nsamp;
  e->got = e->tries = 0;
  memset(e->k, 0, (size_t)e->cap * sizeof(knot_t));
  e->k[0].x = xl;
  e->k[n0 - 1].x = xr;
  for (i = 0; i < ninit; i++) e->k[2 * i + 1].x = xinit[i];
  e->ninit = ninit;
  e->init_i = 0;
  e->waiting = W_INIT;
  *x_out = xinit[0];
  return ARS_NEED;
}

/* feed the log-density at the point last handed out; returns ARS_NEED (next point in *x_out), ARS_DONE or a code */
int stb_ars_feed(stb_ars_t *e, double y, double *x_out) {
  int rc, verdict;
  e->neval++;
  switch (e->waiting) {
    case W_INIT:
      e->k[2 * e->init_i + 1].y = y;
      if (++e->init_i < e->ninit) {
        *x_out = e->k[2 * e->init_i + 1].x;
        return ARS_NEED;
      }
      for (int j = 0; j < e->n; j += 2)
        if (place_intersection(e, j)) return 2000;
      integrate(e);
      if (e->metro) {
        if (e->xprev < e->xl || e->xprev > e->xr) {
          e->xsamp[0] = e->xprev < e->xl ? e->xl : e->xr;
          return 1007;
        }
        e->waiting = W_YPREV;
        *x_out = e->xprev;
        return ARS_NEED;
      }
      return run_proposals(e, x_out);
    case W_YPREV:
      e->yprev = y;
      return run_proposals(e, x_out);
    case W_TRIAL:
      rc = trial_value(e, y, x_out);
      if (rc == ARS_NEED) return rc;
      if (e->iq == IQ_METROPOLIS)
        verdict = 1;
      else if (e->iq == IQ_FULL)
        verdict = e->u_y >= e->p.y ? 0 : 1; /* envelope full: plain rejection test */
      else
        verdict = incorporate_finish(e);
      break;
    case W_ADJUST:
      e->k[e->iq].y = y;
      verdict = incorporate_finish(e);
      break;
    default:
      return 2002;
  }
  e->waiting = W_IDLE;
  rc = proposal_outcome(e, verdict);
  if (rc != ARS_GO_ON) return rc;
  return run_proposals(e, x_out);
}

/* envelope centile (after ARS_DONE) */
double stb_ars_centile(stb_ars_t *e, double q) {
  trial_t p;
  invert_cdf(e, q / 100.0, &p);
  return p.x;
}

/* ---- the reference's entry points: the machine driven by the caller's callback ----------------- */
int arms(double *xinit, int ninit, double *xl, double *xr, double (*myfunc)(double x, void *mydata), void *mydata,
         double *convex, int npoint, int dometrop, double *xprev, double *xsamp, int nsamp, double *qcent,
         double *xcent, int ncent, int *neval) {
  stb_ars_t *e;
  double x = 0;
  int i, rc;
  for (i = 0; i < ncent; i++)
    if (qcent[i] < 0.0 || qcent[i] > 100.0) return 1005;
  /* (the reference checks ninit and npoint before it allocates, lib/arms.c:286-296) */
  if (ninit < 3) return 1001;
  if (npoint < 2 * ninit + 1) return 1002;
  e = stb_ars_new(npoint);
  if (!e) return 1006;
  *neval = 0;
  rc = stb_ars_begin(e, xinit, ninit, *xl, *xr, *convex, dometrop, *xprev, xsamp, nsamp, NULL, NULL, &x);
  while (rc == ARS_NEED) rc = stb_ars_feed(e, myfunc(x, mydata), &x);
  *neval = e->neval;
  if (rc == ARS_DONE)
    for (i = 0; i < ncent; i++) xcent[i] = stb_ars_centile(e, qcent[i]);
  stb_ars_free(e);
  return rc;
}

int arms_simple(int ninit, double *xl, double *xr, double (*myfunc)(double x, void *mydata), void *mydata,
                int dometrop, double *xprev, double *xsamp) {
  double convex = 1.0, qcent = 0, xcent = 0;
  int neval = 0, i, err;
  double *xinit;
  if (ninit < 1) return 1001;
  xinit = (double *)malloc((size_t)ninit * sizeof(double));
  if (!xinit) return 1006;
  /* ninit starting points, equally spaced inside the bounds */
  for (i = 0; i < ninit; i++) xinit[i] = *xl + (i + 1.0) * (*xr - *xl) / (ninit + 1.0);
  err = arms(xinit, ninit, xl, xr, myfunc, mydata, &convex, 100, dometrop, xprev, xsamp, 1, &qcent, &xcent, 0, &neval);
  free(xinit);
  return err;
}
