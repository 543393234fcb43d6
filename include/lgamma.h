/*
 * lgamma.h -- differences of lgamma / digamma with small self-filling caches (reference interface:
 * lib/lgamma.h:22-37, implementation lib/lgamma.c:30-240).  Host scalar helpers used by the
 * table-free discount sampler (lib/samplea.c:92-139) and by callers' own likelihood code
 * (test/check.c:197-207).
 */
#ifndef STB_B200_LGAMMA_H
#define STB_B200_LGAMMA_H
#ifdef __cplusplus
extern "C" {
#endif

#define GCACHE 100
struct gcache_s {
  double par;
  double lgpar;
  double cache[GCACHE];
};

void gcache_init(struct gcache_s *lpg, double p);   /* values lgamma(j+p) - lgamma(p) */
double gcache_value(struct gcache_s *lpg, int j);
void pcache_init(struct gcache_s *lpg, double p);   /* values digamma(j+p) - digamma(p) */
double pcache_value(struct gcache_s *lpg, int j);
void qcache_init(struct gcache_s *lpg, double p);   /* values S^{j+1}_{2,p} / S^j_{1,p} */
double qcache_value(struct gcache_s *lpg, int j);

double gammadiff(int N, double alpha, double lga); /* lgamma(N+alpha) - lgamma(alpha); lga = lgamma(alpha) or 0 */
double psidiff(int N, double alpha, double pa);    /* digamma(N+alpha) - digamma(alpha); pa = digamma(alpha) or 0 */

#ifdef __cplusplus
}
#endif
#endif
