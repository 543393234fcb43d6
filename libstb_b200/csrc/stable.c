/*
 * stable.c -- C host layer of the Stirling table engine: the stable.h API over device tables.
 *
 * Mirrors the behaviour of the reference's lib/stable.c (cited per function) but owns no
 * arithmetic on table cells: every cell is produced by the CUDA kernels behind stb_cuda.h.
 * What lives here is the reference's control logic -- bounds clamping, the growth policy,
 * look-up conventions, the asymptote, the S1 cache, reporting -- plus the host mirror that
 * keeps scalar look-ups O(1).
 *
 * Deliberate deviations from the reference (documented in DESIGN.md):
 *  - S_extend: the reference caps the new M at the OLD usedN (lib/stable.c:627-629), which
 *    leaves a request with m > old usedN un-served and reads out of bounds afterwards; here
 *    M is capped at the NEW N, so the request is served.
 *  - S_S1: when a request grows the S1 cache the reference overwrites its own argument with
 *    the grown size and answers for that row instead (lib/stable.c:845-872); here the answer
 *    is always for the row that was asked.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "stable.h"
#include "stb_b200.h"
#include "stb_cuda.h"
#include "yaps.h"

/* device slabs of tables whose maximum extent is at most this many bytes are allocated whole by S_make */
#define PRESIZE_BYTES (64ull << 20)
/* rows per mirror block are chosen so that a block is about this many bytes */
#define MIRROR_BLOCK_BYTES (4u << 20)

struct stb_table_impl {
  stb_dev_t *dev;
  int algo;
  size_t ld; /* elements per row, device and mirror alike */
  /* Host mirror: pinned blocks of blk_rows rows, fetched from the device on first touch.  A refill does
   * not free them: it only starts a new epoch, and a block whose epoch is old is fetched again when a
   * look-up next lands in it -- a caller that alternates S_remake and a few look-ups (the samplers' loop,
   * test/demo.c:467-488) pays for the rows it reads, not for pinning and copying the whole table. */
  unsigned blk_rows, nblk; /* blk_rows is a power of two: a scalar look-up finds its block with a shift and a mask */
  unsigned blk_shift;
  double **blkS, **blkV;
  uint32_t *epochS, *epochV; /* epoch the block's content belongs to (0: never fetched) */
  uint32_t epoch;            /* of the current device content; bumped by every fill */
  /* S_THREADS (lib/stable.h:43, lib/stable.c:572-580): look-ups from several threads, growth
   * serialised.  Growth here refills the table and invalidates the host mirror, so look-ups hold the
   * lock shared while they read and growth holds it exclusively; fetching a mirror block
   * (done under the shared lock) has its own mutex. */
  pthread_rwlock_t rw;
  pthread_mutex_t fetch_mutex;
  int locking;
  /* The strip kernel produces log S^n_1 on the device; the host vector S1 (S_S1, S_S(n,1)) is
   * refreshed from it on first use after a fill, not by every fill: callers that only do batched
   * look-ups (samplea) never pay for the copy. */
  int s1_stale;
  float last_gibbs_ms; /* device time of the most recent stb_ti_gibbs kernel */
};

static void lock(stable_t *sp) { /* exclusive: growth, refill */
  if (sp->impl->locking) pthread_rwlock_wrlock(&sp->impl->rw);
}
static void unlock(stable_t *sp) {
  if (sp->impl->locking) pthread_rwlock_unlock(&sp->impl->rw);
}
static void rlock(stable_t *sp) { /* shared: reading cells */
  if (sp->impl->locking) pthread_rwlock_rdlock(&sp->impl->rw);
}

/* ------------------------------------------------------------------------------------------ */
/* host mirror                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static void mirror_drop(stable_t *sp) {
  struct stb_table_impl *im = sp->impl;
  unsigned b;
  for (b = 0; b < im->nblk; b++) {
    if (im->blkS && im->blkS[b]) stb_cuda_host_free(im->blkS[b]);
    if (im->blkV && im->blkV[b]) stb_cuda_host_free(im->blkV[b]);
  }
  free(im->blkS);
  free(im->blkV);
  free(im->epochS);
  free(im->epochV);
  im->blkS = im->blkV = NULL;
  im->epochS = im->epochV = NULL;
  im->nblk = 0;
}

/* after the device tables were (re)filled: every block is stale; the bookkeeping is rebuilt only when the
 * extent or the row pitch changed (growth).  0 on success */
static int mirror_reset(stable_t *sp) {
  struct stb_table_impl *im = sp->impl;
  const int hasS = (sp->flags & S_STABLE) != 0, hasV = (sp->flags & S_UVTABLE) != 0;
  const size_t ld = stb_cuda_table_ld(im->dev);
  unsigned blk_rows = (unsigned)(MIRROR_BLOCK_BYTES / (ld * sizeof(double)));
  unsigned nblk;
  unsigned blk_shift = 0;
  if (blk_rows < 1) blk_rows = 1;
  while ((2u << blk_shift) <= blk_rows) blk_shift++;
  blk_rows = 1u << blk_shift;
  nblk = (sp->usedN + blk_rows - 1) / blk_rows;
  if (++im->epoch == 0) im->epoch = 1;
  if (ld == im->ld && blk_rows == im->blk_rows && nblk == im->nblk && (!hasS || im->blkS) && (!hasV || im->blkV))
    return 0; /* same geometry: the pinned blocks stay, their epochs are old now */
  mirror_drop(sp);
  im->ld = ld;
  im->blk_rows = blk_rows;
  im->blk_shift = blk_shift;
  im->nblk = nblk;
  if (hasS && (!(im->blkS = (double **)calloc(nblk, sizeof(double *))) || !(im->epochS = (uint32_t *)calloc(nblk, sizeof(uint32_t)))))
    return 1;
  if (hasV && (!(im->blkV = (double **)calloc(nblk, sizeof(double *))) || !(im->epochV = (uint32_t *)calloc(nblk, sizeof(uint32_t)))))
    return 1;
  return 0;
}

/* slow path of a read: bring one block of rows across (again); NULL on failure */
static double *mirror_fetch(stable_t *sp, int which, unsigned b) {
  struct stb_table_impl *im = sp->impl;
  double **tab = which == STB_TAB_S ? im->blkS : im->blkV;
  uint32_t *ep = which == STB_TAB_S ? im->epochS : im->epochV;
  double *blk;
  if (im->locking) pthread_mutex_lock(&im->fetch_mutex);
  blk = tab[b];
  if (!blk || __atomic_load_n(&ep[b], __ATOMIC_ACQUIRE) != im->epoch) {
    unsigned row0 = b * im->blk_rows, rows = im->blk_rows;
    if (row0 + rows > sp->usedN) rows = sp->usedN - row0;
    if (!blk) blk = (double *)stb_cuda_host_alloc((size_t)im->blk_rows * im->ld * sizeof(double));
    if (blk && stb_cuda_read_rows(im->dev, which, row0, rows, blk)) {
      stb_cuda_host_free(blk);
      blk = NULL;
    }
    tab[b] = blk;
    /* publish only after the block is complete: readers are lock-free */
    if (blk) __atomic_store_n(&ep[b], im->epoch, __ATOMIC_RELEASE);
  }
  if (im->locking) pthread_mutex_unlock(&im->fetch_mutex);
  return blk;
}

/* cell (n,m), 1<=m<=n<=usedN, m<=usedM */
static double cell(stable_t *sp, int which, unsigned n, unsigned m) {
  struct stb_table_impl *im = sp->impl;
  double v;
  rlock(sp);
  {
    double **tab = which == STB_TAB_S ? im->blkS : im->blkV;
    const uint32_t *ep = which == STB_TAB_S ? im->epochS : im->epochV;
    const unsigned b = (n - 1) >> im->blk_shift;
    double *blk = tab[b];
    if (__atomic_load_n(&ep[b], __ATOMIC_ACQUIRE) != im->epoch && !(blk = mirror_fetch(sp, which, b)))
      yaps_quit("stable: device read failed: %s\n", stb_cuda_last_error());
    v = blk[(size_t)((n - 1) & (im->blk_rows - 1)) * im->ld + (m - 1)];
  }
  unlock(sp);
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* fill                                                                                        */
/* ------------------------------------------------------------------------------------------ */
/*
 * Fill the device tables for rows<=N, cols<=M at discount a and refresh S1 + mirror.
 * Counterpart of S_remake_part (lib/stable.c:321-547); bounds are published last (:543-545).
 */
static int fill(stable_t *sp, double a, unsigned startN, unsigned startM, unsigned N, unsigned M) {
  struct stb_table_impl *im = sp->impl;
  const int hasS = (sp->flags & S_STABLE) != 0;
  const int mirror_order = im->algo == STB_FILL_MIRROR;
  unsigned n;
  sp->a = a;
  sp->lga = lgamma(1.0 - a);
  if (mirror_order || !hasS) {
    /* S1 running sum in the reference's order, lib/stable.c:338-348 */
    sp->S1[0] = 0;
    for (n = 2; n <= N; n++) sp->S1[n - 1] = sp->S1[n - 2] + log(n - 1 - a);
  }
  if (stb_cuda_fill(im->dev, a, startN, startM, N, M, im->algo, mirror_order ? sp->S1 : NULL)) return 1;
  im->s1_stale = hasS && !mirror_order;
  /* a changed: forget the lazily computed tail of S1, lib/stable.c:350-353 */
  for (n = N + 1; n <= sp->usedN1; n++) sp->S1[n - 1] = 0;
  sp->usedN = N;
  sp->usedM = M;
  return mirror_reset(sp);
}

stable_t *S_make(unsigned initN, unsigned initM, unsigned maxN, unsigned maxM, double a, uint32_t flags) {
  stable_t *sp;
  struct stb_table_impl *im;
  /* lib/stable.c:118-129 */
  if (maxM < 10) maxM = 10;
  if (maxN < maxM) maxN = maxM;
  if (initM < 10) initM = 10;
  if (initN < initM) initN = initM;
  if (initN > maxN) initN = maxN;
  if (initM > maxM) initM = maxM;
  /* lib/stable.c:131-132 */
  if ((flags & S_STABLE) == 0 && (flags & S_UVTABLE) == 0) return NULL;

  sp = (stable_t *)calloc(1, sizeof *sp);
  im = (struct stb_table_impl *)calloc(1, sizeof *im);
  if (!sp || !im) {
    free(sp);
    free(im);
    return NULL;
  }
  sp->impl = im;
  sp->flags = flags;
  sp->maxN = maxN;
  sp->maxM = maxM;
  sp->usedN = initN;
  sp->usedM = initM;
  sp->usedN1 = initN;
  pthread_rwlock_init(&im->rw, NULL);
  pthread_mutex_init(&im->fetch_mutex, NULL);
  im->locking = (flags & S_THREADS) != 0;
  im->algo = (flags & S_MIRROR_ORDER) ? STB_FILL_MIRROR : STB_FILL_LINEAR;
  {
    const char *s = getenv("STB_FILL_ALGO");
    if (s && !strcmp(s, "mirror")) im->algo = STB_FILL_MIRROR;
    if (s && !strcmp(s, "linear")) im->algo = STB_FILL_LINEAR;
  }
  sp->S1 = (double *)calloc(initN, sizeof(double));
  im->dev = stb_cuda_table_create((flags & S_STABLE) != 0, (flags & S_UVTABLE) != 0, (flags & S_FLOAT) != 0);
  if (im->dev) {
    /* a table whose MAXIMUM extent is small gets all of it at once: it never reallocates when it grows */
    const uint64_t es = (flags & S_FLOAT) ? 4 : 8, ntab = ((flags & S_STABLE) != 0) + ((flags & S_UVTABLE) != 0);
    const uint64_t full = (uint64_t)maxN * (((uint64_t)maxM + 31) / 32 * 32) * es * ntab;
    const int whole = full <= PRESIZE_BYTES;
    if (stb_cuda_table_reserve(im->dev, whole ? maxN : initN, whole ? maxM : initM, 0)) {
      stb_cuda_table_destroy(im->dev);
      im->dev = NULL;
    }
  }
  if (!sp->S1 || !im->dev || fill(sp, a, 0, 0, initN, initM)) {
    if (stb_cuda_last_error()[0]) yaps_message("S_make: %s\n", stb_cuda_last_error());
    S_free(sp);
    return NULL;
  }
  if (flags & S_VERBOSE) S_report(sp, stderr);
  return sp;
}

void S_tag(stable_t *S, char *tag) {
  free(S->tag);
  S->tag = (char *)malloc(strlen(tag) + 1);
  if (S->tag) strcpy(S->tag, tag);
}

int S_remake(stable_t *sp, double a) {
  int ret = fill(sp, a, 0, 0, sp->usedN, sp->usedM);
  if (!ret && (sp->flags & S_VERBOSE)) S_report(sp, stderr);
  return ret;
}

/*
 * Growth, lib/stable.c:564-815: a request for (N,M) grows each dimension by at least 10% and
 * at least 50, capped by the maxima.  Returns non-zero when memory (host or device) ran out.
 */
static int extend(stable_t *sp, unsigned Nreq, unsigned Mreq) {
  struct stb_table_impl *im = sp->impl;
  long N = (long)Nreq + 1, M = (long)Mreq + 1;
  int result = 0;
  lock(sp);
  if (N < (long)sp->usedN && M < (long)sp->usedM) goto done; /* someone else already grew it */
  if (N < (long)sp->usedN) N = sp->usedN;
  if (N > (long)sp->maxN) N = sp->maxN;
  if (N > (long)sp->usedN) {
    if (N < sp->usedN * 1.1) N = (long)(sp->usedN * 1.1);
    if (N < (long)sp->usedN + 50) N = (long)sp->usedN + 50;
    if (N > (long)sp->maxN) N = sp->maxN;
  }
  if (M < (long)sp->usedM) M = sp->usedM;
  if (N < M) M = N;
  if (M > (long)sp->maxM) M = sp->maxM;
  if (M > (long)sp->usedM) {
    if (M < sp->usedM * 1.1) M = (long)(sp->usedM * 1.1);
    if (M < (long)sp->usedM + 50) M = (long)sp->usedM + 50;
    if (M > (long)sp->maxM) M = sp->maxM;
    if (M > N) M = N; /* deviation: the reference caps at the old usedN */
  }
  if (N == (long)sp->usedN && M == (long)sp->usedM) goto done;
  if (N > (long)sp->usedN1) {
    double *s1 = (double *)realloc(sp->S1, sizeof(double) * (size_t)N);
    if (!s1) {
      result = 1;
      goto done;
    }
    memset(s1 + sp->usedN1, 0, sizeof(double) * (size_t)(N - sp->usedN1));
    sp->S1 = s1;
    sp->usedN1 = (unsigned)N;
  }
  /* the whole extent is refilled, so nothing of the old slab is kept; capacity doubles (within the table's
   * maxima), so a table grown in 10 % steps -- lib/stable.c:590-601 -- reallocates a handful of times */
  if (stb_cuda_table_grow(im->dev, (unsigned)N, (unsigned)M, sp->maxN, sp->maxM) ||
      fill(sp, sp->a, sp->usedN, sp->usedM, (unsigned)N, (unsigned)M))
    result = 1;
done:
  unlock(sp);
  return result;
}

/* bring S1[0 .. usedN-1] up to date with the last fill (see s1_stale) */
static void s1_sync(stable_t *sp) {
  struct stb_table_impl *im = sp->impl;
  if (!__atomic_load_n(&im->s1_stale, __ATOMIC_ACQUIRE)) return;
  lock(sp);
  if (im->s1_stale) {
    if (stb_cuda_read_s1(im->dev, sp->usedN, sp->S1)) yaps_quit("stable: device read failed: %s\n", stb_cuda_last_error());
    __atomic_store_n(&im->s1_stale, 0, __ATOMIC_RELEASE);
  }
  unlock(sp);
}

/* lib/stable.c:822-873 */
double S_S1(stable_t *sp, unsigned n) {
  if (n == 0) return -HUGE_VAL;
  if (!sp->S1) return -HUGE_VAL;
  s1_sync(sp);
  if (n > sp->usedN) {
    double v;
    if (n > sp->maxN) {
      if (!(sp->flags & S_ASYMPT)) return -HUGE_VAL;
      /* past the bound the closed form is exact for m==1 */
      return lgamma(n - sp->a) - sp->lga;
    }
    lock(sp);
    if (n > sp->usedN1) {
      unsigned want = n;
      double *s1;
      if (want < sp->usedN1 * 1.1) want = (unsigned)(sp->usedN1 * 1.1);
      if (want < sp->usedN1 + 50) want = sp->usedN1 + 50;
      if (want > sp->maxN) want = sp->maxN;
      s1 = (double *)realloc(sp->S1, sizeof(double) * want);
      if (!s1) {
        unlock(sp);
        return -HUGE_VAL;
      }
      memset(s1 + sp->usedN1, 0, sizeof(double) * (want - sp->usedN1));
      sp->S1 = s1;
      sp->usedN1 = want;
    }
    if (sp->S1[n - 1] == 0) {
      if (sp->S1[n - 2] == 0)
        sp->S1[n - 1] = lgamma(n - sp->a) - sp->lga;
      else
        sp->S1[n - 1] = sp->S1[n - 2] + log(n - 1 - sp->a);
    }
    v = sp->S1[n - 1];
    unlock(sp);
    return v;
  }
  {
    double v;
    rlock(sp); /* growth may move the vector */
    v = sp->S1[n - 1];
    unlock(sp);
    return v;
  }
}

/* lib/stable.c:875-883 */
double S_U(stable_t *sp, unsigned n, unsigned m) {
  if (m == 1) return n - sp->a;
  if (m <= 1) yaps_quit("Bad constraints in S_U(%s,%u,%u)\n", sp->tag, n, m);
  return n - m * sp->a + 1 / S_V(sp, n, m);
}

/* lib/stable.c:885-897 */
double S_UV(stable_t *sp, unsigned n, unsigned m) {
  double SV;
  if (m == 1) return -HUGE_VAL;
  if (m == n + 1) return 1; /* S^n_n == 1 */
  if (m == n) return (n + 1.0) / (n - 1.0);
  SV = S_V(sp, n, m);
  return (n - m * sp->a) * SV + 1.0;
}

/* the V asymptote, lib/stable.c:905-911 */
static double v_asympt(stable_t *sp, unsigned n, unsigned m) {
  if (sp->a > 0) return (1.0 - pow(n, -sp->a)) / sp->a / (m - 1);
  {
    double ln = log(n);
    return ln / (m - 1) * exp(lgamma(1 + (m - 2) / ln) - lgamma(1 + (m - 1) / ln));
  }
}

/* lib/stable.c:900-939 */
double S_V(stable_t *sp, unsigned n, unsigned m) {
  if ((sp->flags & S_UVTABLE) == 0) return 0;
  if (m + 1 >= sp->usedM || n + 1 >= sp->usedN) {
    if (n > sp->maxN || m > sp->maxM) {
      if (n > sp->maxN && (sp->flags & S_ASYMPT)) return v_asympt(sp, n, m);
      if (sp->flags & S_QUITONBOUND) {
        if (sp->tag)
          yaps_quit("S_V(%u,%u,%lf) tagged '%s' hit bounds (%u,%u)\n", n, m, sp->a, sp->tag, sp->maxN,
                    sp->maxM);
        else
          yaps_quit("S_V(%u,%u,%lf) hit bounds\n", n, m, sp->a);
      } else
        return 0;
    }
    if (extend(sp, n + 1, m + 1)) yaps_quit("S_extend() out of memory\n");
  }
  if (m < 2 || n < m) return 0;
  return cell(sp, STB_TAB_V, n, m);
}

/* lib/stable.c:941-974 */
double S_S(stable_t *sp, unsigned N, unsigned T) {
  if ((sp->flags & S_STABLE) == 0) return -HUGE_VAL;
  if (N == T) return 0;
  if (T == 1) return S_S1(sp, N);
  if (N < T || T == 0) return -HUGE_VAL;
  if (T > sp->usedM || N > sp->usedN) {
    if (N > sp->maxN || T > sp->maxM) {
      if (N > sp->maxN && (sp->flags & S_ASYMPT)) return S_asympt(sp, N, T);
      if (sp->flags & S_QUITONBOUND) {
        if (sp->tag)
          yaps_quit("S_S(%u,%u,%lf) tagged '%s' hit bounds\n", N, T, sp->a, sp->tag);
        else
          yaps_quit("S_S(%u,%u,%lf) hit bounds\n", N, T, sp->a);
      } else
        return -HUGE_VAL;
    }
    if (extend(sp, N + 1, T + 1)) yaps_quit("S_extend() out of memory\n");
  }
  return cell(sp, STB_TAB_S, N, T);
}

/* lib/stable.c:980-1023 */
void S_free(stable_t *sp) {
  if (!sp) return;
  if (sp->impl) {
    mirror_drop(sp);
    stb_cuda_table_destroy(sp->impl->dev);
    pthread_rwlock_destroy(&sp->impl->rw);
    pthread_mutex_destroy(&sp->impl->fetch_mutex);
    free(sp->impl);
  }
  free(sp->tag);
  free(sp->S1);
  free(sp);
}

/* lib/stable.c:1025-1055: same one-line format, including the trailing blank line */
void S_report(stable_t *sp, FILE *fp) {
  const char *s = (sp->flags & S_STABLE) ? "+S" : "", *uv = (sp->flags & S_UVTABLE) ? "+U/V" : "";
  const char *ty = (sp->flags & S_FLOAT) ? "float" : "double";
  uint64_t bytes = stb_cuda_table_bytes(sp->impl->dev) + (uint64_t)sp->usedN1 * sizeof(double);
  sp->memalloced = (uint32_t)bytes; /* the reference keeps a uint32 too (lib/stable.h:105) */
  if (fp) {
    if (sp->tag)
      fprintf(fp, "S-table '%s': ", sp->tag);
    else
      fprintf(fp, "S-table: ");
    fprintf(fp, "a=%lf, N=%u/%u, M=%u/%u, %s%s %s", sp->a, sp->usedN, sp->maxN, sp->usedM, sp->maxM, s, uv, ty);
    fprintf(fp, " mem=%uk\n", sp->memalloced / 1024);
    fprintf(fp, "\n");
  } else {
    if (sp->tag)
      yaps_message("S-table '%s': ", sp->tag);
    else
      yaps_message("S-table: ");
    yaps_message("a=%lf, N=%u/%u, M=%u/%u, %s%s %s", sp->a, sp->usedN, sp->maxN, sp->usedM, sp->maxM, s, uv, ty);
    yaps_message(" mem=%uk", sp->memalloced / 1024);
    yaps_message("\n");
  }
}

/*
 * lib/stable.c:1057-1084.  a>0: Hutter's form  Gamma(n) / (Gamma(1-a) Gamma(m) a^{m-1} n^a)
 * times (1-n^{-a})^{m-1}; a==0: Hwang's expansion for Stirling numbers of the first kind.
 */
double S_asympt(stable_t *sp, unsigned n, unsigned m) {
  if (sp->a == 0) {
    double ln = log(n);
    return lgamma(n) + (m - 1) * log(ln) - lgamma(m) - lgamma(1 + (m - 1) / ln);
  } else {
    double prod = 0;
    double la1 = lgamma(1.0 - sp->a);
    double aln = sp->a * log((double)n);
    double np = pow(n, -sp->a);
    prod += lgamma((double)n) - la1 - lgamma((double)m) - (m - 1.0) * log(sp->a) - aln;
    if (np < 1e-5)
      prod -= (m - 1) * np * (1 + np * (0.5 + np / 3.0));
    else
      prod += (m - 1) * log(1.0 - np);
    return prod;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* batched extensions (stb_b200.h)                                                             */
/* ------------------------------------------------------------------------------------------ */
static int batch(stable_t *sp, int which, const uint32_t *n, const uint32_t *m, double *out, size_t count,
                 int on_device) {
  if (!sp || !sp->impl) return 1;
  if (which == STB_TAB_S && !(sp->flags & S_STABLE)) return 1;
  if (which != STB_TAB_S && !(sp->flags & S_UVTABLE)) return 1;
  if (!on_device) {
    /* grow once to cover the whole batch, like the scalar calls would one by one */
    unsigned maxn = 0, maxm = 0;
    size_t i;
    for (i = 0; i < count; i++) {
      if (n[i] >= m[i] && n[i] <= sp->maxN && m[i] <= sp->maxM) {
        if (n[i] > maxn) maxn = n[i];
        if (m[i] > maxm) maxm = m[i];
      }
    }
    if (maxn > sp->usedN || maxm > sp->usedM)
      if (extend(sp, maxn > sp->usedN ? maxn : sp->usedN, maxm > sp->usedM ? maxm : sp->usedM)) return 1;
  }
  {
    /* exclusive: the call uses the handle's stream and staging buffers, and a growth in another
     * thread (S_THREADS) would move the table under the kernel */
    int rc;
    lock(sp);
    rc = stb_cuda_gather(sp->impl->dev, which, sp->a, sp->usedN, sp->usedM, n, m, out, count, on_device);
    unlock(sp);
    return rc;
  }
}

int stb_S_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_TAB_S, n, m, out, count, 0);
}
int stb_V_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_TAB_V, n, m, out, count, 0);
}
int stb_U_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_GATHER_U, n, m, out, count, 0);
}
int stb_UV_batch(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_GATHER_UV, n, m, out, count, 0);
}
int stb_U_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_GATHER_U, n, m, out, count, 1);
}
int stb_UV_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_GATHER_UV, n, m, out, count, 1);
}
int stb_S_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_TAB_S, n, m, out, count, 1);
}
int stb_V_batch_device(stable_t *sp, const uint32_t *n, const uint32_t *m, double *out, size_t count) {
  return batch(sp, STB_TAB_V, n, m, out, count, 1);
}

/* samplea2's partition sampler over this table (stb_b200.h); lib/samplea.c:290-321 */
int stb_partition_sample(stable_t *sp, double a, const uint32_t *n, const uint16_t *t, const double *logu,
                         const uint32_t *off, size_t count, uint16_t *m_out, size_t n_m, int exact) {
  unsigned maxn = 0, maxt = 0;
  size_t i;
  if (!sp || !sp->impl || !(sp->flags & S_STABLE)) {
    stb_cuda_set_error("stb_partition_sample: no S table", 0);
    return 1;
  }
  for (i = 0; i < count; i++) {
    if (t[i] < 2 || t[i] >= n[i] || (size_t)off[i] + t[i] - 1 > n_m) {
      stb_cuda_set_error("stb_partition_sample: node needs 1 < t < n and room for t-1 sizes", 0);
      return 1;
    }
    if (n[i] > maxn) maxn = n[i];
    if (t[i] > maxt) maxt = t[i];
  }
  if (maxn > sp->maxN || maxt > sp->maxM) {
    /* S_S would answer -inf (or the asymptote) here and the reference would sample from garbage */
    stb_cuda_set_error("stb_partition_sample: counts beyond the table's maximum size", 0);
    return 1;
  }
  if (maxn > sp->usedN || maxt > sp->usedM)
    if (extend(sp, maxn > sp->usedN ? maxn : sp->usedN, maxt > sp->usedM ? maxt : sp->usedM)) return 1;
  {
    int rc;
    lock(sp); /* exclusive, like the batched look-ups */
    rc = stb_cuda_partition(sp->impl->dev, a, n, t, logu, off, count, m_out, n_m, exact);
    unlock(sp);
    return rc;
  }
}

/*
 * Table-indicator Gibbs sweeps on the device (stb_b200.h; the per-token update of test/demo.c:405-434).
 * The table is grown first to cover every (n, t+1) a sweep can look up, so that the kernel never needs the host.
 */
int stb_ti_gibbs(stable_t *sp, double bpar, size_t R, const uint32_t *tok_off, const uint32_t *tok_dish, const float *H,
                 uint32_t D, const uint32_t *n, uint16_t *t, uint32_t *T, uint64_t *rng, int shared_stream, int sweeps) {
  unsigned maxn = 0;
  size_t j, c;
  uint32_t i;
  if (!sp || !sp->impl || !(sp->flags & S_UVTABLE)) {
    stb_cuda_set_error("stb_ti_gibbs: the table has no V (make it with S_UVTABLE)", 0);
    return 1;
  }
  if (!R || sweeps < 1) return 0;
  for (j = 0; j < R; j++) {
    uint32_t Tj = 0;
    if (tok_off[j + 1] < tok_off[j]) {
      stb_cuda_set_error("stb_ti_gibbs: token offsets must not decrease", 0);
      return 1;
    }
    for (i = 0; i < D; i++) {
      const uint32_t nn = n[j * (size_t)D + i], tt = t[j * (size_t)D + i];
      if (tt > nn || (nn > 0 && tt == 0)) {
        stb_cuda_set_error("stb_ti_gibbs: counts need 1 <= t <= n wherever n > 0", 0);
        return 1;
      }
      if (nn > maxn) maxn = nn;
      Tj += tt;
    }
    if (Tj != T[j]) {
      stb_cuda_set_error("stb_ti_gibbs: T[j] is not the sum of t[j][.]", 0);
      return 1;
    }
    for (c = tok_off[j]; c < tok_off[j + 1]; c++)
      if (tok_dish[c] >= D || n[j * (size_t)D + tok_dish[c]] == 0) {
        stb_cuda_set_error("stb_ti_gibbs: a token names a dish the restaurant does not serve", 0);
        return 1;
      }
  }
  /* a look-up is V(n, t+1) with t+1 <= n: rows up to max n, columns up to min(max n, maxM) */
  if (maxn > sp->maxN) {
    stb_cuda_set_error("stb_ti_gibbs: counts beyond the table's maximum size", 0);
    return 1;
  }
  {
    unsigned wantM = maxn < sp->maxM ? maxn : sp->maxM;
    if (maxn > sp->usedN || wantM > sp->usedM)
      if (extend(sp, maxn > sp->usedN ? maxn : sp->usedN, wantM > sp->usedM ? wantM : sp->usedM)) return 1;
  }
  {
    int rc;
    float ms = 0;
    lock(sp);
    rc = stb_cuda_ti_gibbs(sp->impl->dev, sp->usedN, sp->usedM, sp->a, bpar, R, tok_off, tok_dish, H, D, n, t, T, rng,
                           shared_stream, sweeps, &ms);
    sp->impl->last_gibbs_ms = ms;
    unlock(sp);
    return rc;
  }
}
double stb_last_gibbs_ms(const stable_t *sp) { return sp && sp->impl ? sp->impl->last_gibbs_ms : 0; }

/* ------------------------------------------------------------------------------------------ */
/* discount sweep (stb_b200.h)                                                                 */
/* ------------------------------------------------------------------------------------------ */
struct stb_sweep {
  stb_sweep_dev_t *dev;
  float last_ms;
};

stb_sweep_t *stb_sweep_create(unsigned N, unsigned M, uint32_t flags) {
  stb_sweep_t *w = (stb_sweep_t *)calloc(1, sizeof *w);
  if (!w) return NULL;
  w->dev = stb_cuda_sweep_create(N, M, (flags & S_FLOAT) != 0);
  if (!w->dev) {
    free(w);
    return NULL;
  }
  return w;
}
int stb_sweep_set_pairs(stb_sweep_t *w, const uint32_t *n, const uint32_t *m, size_t npairs) {
  return w ? stb_cuda_sweep_set_pairs(w->dev, n, m, npairs) : 1;
}
int stb_sweep_run(stb_sweep_t *w, const double *a, size_t na, double *gather_out, double *sum_out,
                  double *lastrow_out) {
  return w ? stb_cuda_sweep_run(w->dev, a, na, gather_out, sum_out, lastrow_out, &w->last_ms) : 1;
}
/* the batched partition step of samplea2 (psample_batch.c; declared in psample_core.h) */
int stb_sweep_set_nodes(stb_sweep_t *w, const uint32_t *n, const uint16_t *t, const uint32_t *draw, size_t count,
                        const uint32_t *hbase, unsigned hbins) {
  return w ? stb_cuda_sweep_set_nodes(w->dev, n, t, draw, count, hbase, hbins) : 1;
}
int stb_sweep_partition(stb_sweep_t *w, const double *a, size_t na, const uint64_t *x0, int exact, uint32_t *hist_out,
                        float *ms) {
  return w ? stb_cuda_sweep_partition(w->dev, a, na, x0, exact, hist_out, ms) : 1;
}
int stb_sweep_hist_eval(stb_sweep_t *w, const double *x, const int *chain, size_t cnt, double *out, float *ms) {
  return w ? stb_cuda_sweep_hist_eval(w->dev, x, chain, cnt, out, ms) : 1;
}
double stb_sweep_last_fill_ms(const stb_sweep_t *w) { return w->last_ms; }
int stb_sweep_tables_in_flight(const stb_sweep_t *w) { return stb_cuda_sweep_tables_in_flight(w->dev); }
int stb_sweep_tables_per_launch(const stb_sweep_t *w) { return stb_cuda_sweep_tables_per_launch(w->dev); }
void stb_sweep_free(stb_sweep_t *w) {
  if (!w) return;
  stb_cuda_sweep_destroy(w->dev);
  free(w);
}

int stb_extend(stable_t *sp, unsigned N, unsigned M) {
  if (!sp || N > sp->maxN || M > sp->maxM) return 1;
  if (N <= sp->usedN && M <= sp->usedM) return 0;
  return extend(sp, N > sp->usedN ? N : sp->usedN, M > sp->usedM ? M : sp->usedM);
}

int stb_read_rows(stable_t *sp, int which_V, unsigned n0, unsigned nrows, double *dst) {
  if (!sp || n0 < 1 || n0 + nrows - 1 > sp->usedN) return 1;
  if (which_V ? !(sp->flags & S_UVTABLE) : !(sp->flags & S_STABLE)) return 1;
  return stb_cuda_read_rows(sp->impl->dev, which_V ? STB_TAB_V : STB_TAB_S, n0 - 1, nrows, dst);
}

double stb_last_fill_ms(const stable_t *sp) { return stb_cuda_last_fill_ms(sp->impl->dev); }
double stb_last_partition_ms(const stable_t *sp) { return stb_cuda_last_partition_ms(sp->impl->dev); }
const void *stb_device_table(const stable_t *sp, int which_V, size_t *ld) {
  if (ld) *ld = stb_cuda_table_ld(sp->impl->dev);
  return stb_cuda_table_ptr(sp->impl->dev, which_V ? STB_TAB_V : STB_TAB_S);
}
int stb_device_count(void) { return stb_cuda_device_count(); }
const char *stb_last_error(void) { return stb_cuda_last_error(); }
