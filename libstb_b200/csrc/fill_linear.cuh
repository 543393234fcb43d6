/*
 * fill_linear.cuh -- the throughput path of the table fill (sm_100a).
 *
 * What it replaces: the two double loops of S_remake_part, lib/stable.c:356-388 (log S) and
 * :451-482 (V), plus the S1 prefix :338-348.
 *
 * Formulation.  The reference iterates in log space, S' = logadd(log(c)+S_up, S_left), which
 * puts an exp->log chain (~50 dependent FP64 instructions) on the row-to-row critical path.
 * Here the SAME recurrence  S^n_m = (n-1-m a) S^{n-1}_m + S^{n-1}_{m-1}  runs in the linear
 * domain on scaled values  S = x * 2^E  (x: double, E: integer kept per lane), so the
 * critical path per row is ONE DFMA; the logarithm is taken once per STORED cell, off the
 * critical path, by separate warps:  log S = log(x) + E ln2.  V^n_m = S^n_m / S^n_{m-1} is the
 * ratio of two neighbouring scaled values, one division per stored cell.  Scaling by powers of
 * two is exact, so results do not depend on how columns are partitioned or when a lane
 * renormalises: fills are bit-reproducible across launch geometries.
 * Agreement with the reference: <= ~1e-14 relative on log S and V (tests/).
 *
 * Geometry.  Columns are cut into warp-strips of 32*K columns (lane l owns K adjacent columns).
 * A producer warp walks its strip down the rows in a diagonal wavefront: at step t lane l
 * computes row rs+1+t-l, so the left neighbour's value it needs (one row up, one column left)
 * was finished two steps earlier and the warp shuffle that fetches it is off the dependent
 * chain.  Producers drop raw x values into a shared-memory ring; consumer warps pick up whole
 * rows, take log / divide, and write each row segment to HBM exactly once with coalesced
 * 256-byte stores.  Warp-strips hand their last column to the right neighbour through small
 * rings: in shared memory inside a CTA, through an L2-resident global ring between CTAs
 * (loader / flusher helper warps, release/acquire counters).  All CTAs are co-resident
 * (cooperative launch, <= 1 CTA per SM).  No tensor cores: nothing here is a contraction.
 *
 * Roofline: 8 B (4 B float) written per cell, 0 B read; ~12 FP64-pipe instructions per S cell
 * (3 recurrence + 9 log), ~+8 per V cell.  HBM-write bound on B200 for the FP64 table.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "fill_common.cuh"

namespace stb {

constexpr int LIN_B = 8;          // steps per batch; a lane renormalises once per batch
constexpr int LIN_RBI = 64;       // boundary ring entries between two strips of one CTA (power of two)
constexpr int LIN_RBX = 256;      // boundary ring entries at a CTA edge (fed/drained through L2)
constexpr int LIN_RBG = 2048;     // boundary ring entries in global memory per CTA boundary
constexpr int LIN_CONS = 2;       // consumer warps per warp-strip
constexpr int LIN_CHUNK = 32;     // least rows a loader/flusher moves per round trip
constexpr int LIN_LOGTAB = LOGTAB_N;
constexpr long long LIN_WATCHDOG = FILL_WATCHDOG;

struct __align__(16) BndEntry {
  double x;  // value of the strip's last column at this row, in units of 2^E
  int elow;  // low 32 bits of that lane's E
  int pad;
};

struct RingCtl {
  int written;  // entries for rows <= written are valid
  int taken;    // the reader is done with rows <= taken
};

struct LinParams {
  void *tabS, *tabV;
  double *s1;
  unsigned long long ld;
  double a;
  int N, M;
  int nws;  // warp-strips in this launch
  int G;    // warp-strips per CTA
  BndEntry *gring;
  int *gwritten, *gtaken;
  int *abort_flag;
  const LogTabEntry *logtab;
};

/*
 * Warp-collective wait until *ctr >= need (all 32 lanes call it with the same arguments; the
 * counter load is one broadcast access, so every lane sees the same value and the loop stays
 * converged).  Returns false if the fill was aborted (watchdog or another role's failure);
 * every role then drains out of the kernel, so a protocol bug can never hang the GPU.
 *
 * Ordering inside a CTA: the counters and the data they guard live in shared memory, whose
 * accesses the SM performs in issue order -- ptxas itself lowers ld.acquire.cta.shared to a
 * plain LDS and mbarrier.arrive.release to a bare SYNCS -- so the shared-memory protocol uses
 * volatile accesses plus compiler barriers and no MEMBAR (a MEMBAR.SC.CTA per hand-off costs
 * more than the eight recurrence steps it guards).  The global-memory hand-off between CTAs
 * uses real release/acquire.
 */
template <bool GLOBAL, int SLEEP>
__device__ __forceinline__ bool wait_ge(const int *ctr, int need, int *abort_flag, int &cached) {
  if (cached >= need) return true;
  // global counters are polled relaxed; one acquire fence follows the successful read
  int v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_vol(ctr);
  if (v < need) {
    const long long t0 = clock64();
    unsigned spins = 0;
    do {
      if (SLEEP) __nanosleep(SLEEP);
      v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_vol(ctr);
      if ((++spins & 1023u) == 0) {
        const int bad = ld_vol(abort_flag) || (clock64() - t0 > LIN_WATCHDOG);
        if (__any_sync(0xffffffffu, bad)) {
          if ((threadIdx.x & 31) == 0) atomicExch(abort_flag, 1);
          return false;
        }
      }
    } while (v < need);
  }
  cached = v;
  if (GLOBAL)
    fence_acq_rel_gpu();
  else
    asm volatile("" ::: "memory");
  return true;
}

// ---------------------------------------------------------------------------------------------
// shared memory carve-up (per CTA)
// ---------------------------------------------------------------------------------------------
template <int K, bool HAS_V, int RS>
struct LinSmem {
  static constexpr int W = 32 * K;               // columns per warp-strip
  static constexpr int NJ = RS / LIN_B * 2;      // batches of E kept
  static constexpr size_t xring_bytes = (size_t)RS * W * 8;
  static constexpr size_t yring_bytes = HAS_V ? (size_t)RS * 32 * 8 : 0;
  static constexpr size_t ering_bytes = (size_t)NJ * 32 * 8;
  static constexpr size_t strip_bytes = xring_bytes + yring_bytes + ering_bytes;
  static constexpr size_t ringi_bytes = (size_t)LIN_RBI * sizeof(BndEntry);
  static constexpr size_t ringx_bytes = (size_t)LIN_RBX * sizeof(BndEntry);
  // layout: logtab | 2 edge rings | (G-1) interior rings | G strips | control words
  __host__ __device__ static size_t total(int G) {
    return (size_t)LIN_LOGTAB * sizeof(LogTabEntry) + 16 + 2 * ringx_bytes + (size_t)(G - 1) * ringi_bytes +
           (size_t)G * strip_bytes + (size_t)(G + 1) * sizeof(RingCtl) +
           (size_t)G * (1 + LIN_CONS) * sizeof(int) + 64;
  }
};

// ---------------------------------------------------------------------------------------------
// producer: the recurrence, one warp per warp-strip
// ---------------------------------------------------------------------------------------------
template <int K, bool HAS_V, int RS, bool EDGE>
__device__ __forceinline__ void producer_batch(
    double (&x)[K], double &yin, int &elow, const double (&ma)[K], double &nm1, int n_lane, int rs,
    int N, int lane, bool left_seed, bool has_right, const BndEntry *ring_in, int mask_in,
    BndEntry *ring_out, int mask_out, double *xring, double *yring) {
  constexpr int W = 32 * K;
#pragma unroll
  for (int i = 0; i < LIN_B; i++) {
    const int n = n_lane + i;  // row this lane computes in this step
    const bool act = EDGE ? (n > rs && n <= N) : true;
    // --- neighbour exchange: the state BEFORE this step's update ---
    double s = shfl_up_d(x[K - 1]);
    int sE = __shfl_up_sync(0xffffffffu, elow, 1);
    if (lane == 0) {
      if (left_seed) {
        s = 0.0;
        sE = elow;
      } else {
        int rn = n < N ? n : N;
        const BndEntry e = ring_in[rn & mask_in];
        s = e.x;
        sE = e.elow;
      }
    }
    // --- the recurrence ---
    double xn[K];
    xn[0] = fma(nm1 - ma[0], x[0], yin);
#pragma unroll
    for (int k = 1; k < K; k++) xn[k] = fma(nm1 - ma[k], x[k], x[k - 1]);
    if (act) {
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = xn[k];
    }
    // left value for the next row, rescaled to this lane's exponent
    yin = s * pow2i(sE - elow);
    nm1 += 1.0;
    if (act) {
      const int slot = n & (RS - 1);
      double *xr = xring + (size_t)slot * W + lane * K;
      if (K == 1) {
        xr[0] = x[0];
      } else if (K == 2) {
        *reinterpret_cast<double2 *>(xr) = make_double2(x[0], x[1]);
      } else {
#pragma unroll
        for (int k = 0; k < K; k += 2) *reinterpret_cast<double2 *>(xr + k) = make_double2(x[k], x[k + 1]);
      }
      if (HAS_V) yring[slot * 32 + lane] = yin;
      if (has_right && lane == 31) {
        BndEntry e;
        e.x = x[K - 1];
        e.elow = elow;
        e.pad = 0;
        ring_out[n & mask_out] = e;
      }
    }
  }
}

template <int K, bool HAS_V, int RS>
__device__ void producer_warp(const LinParams &P, int ws, int lane, bool left_seed, bool has_right,
                              const BndEntry *ring_in, int mask_in, RingCtl *ctl_in, BndEntry *ring_out,
                              int mask_out, RingCtl *ctl_out, double *xring, double *yring, double *ering,
                              int *prod_done, const int *cfin) {
  constexpr int W = 32 * K;
  constexpr int NJ = LinSmem<K, HAS_V, RS>::NJ;
  const int N = P.N;
  const int m_first = ws * W + 1;
  const int rs = m_first - 1;  // every column of this strip is still zero at row rs (m > n)
  const int bfirst = (rs + 1) >> 3;

  double x[K], ma[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    x[k] = 0.0;
    ma[k] = (double)(m_first + lane * K + k) * P.a;
  }
  long long E = 0;
  int elow = 0;
  double nm1 = (double)(rs - lane);  // (n-1) for n = rs+1-lane at step 0
  double yin = 0.0;
  if (lane == 0) {
    if (left_seed) {
      yin = (rs == 0) ? 1.0 : 0.0;  // S^0_0 = 1 seeds S^1_1 = 1
    }
  }
  int c_in = ctl_in ? rs - 1 : 0, c_out = 0, c_cons = 0;
  if (!left_seed) {
    // initial left value: the boundary entry of row rs
    if (!wait_ge<false, 0>(&ctl_in->written, rs, P.abort_flag, c_in)) return;
    if (lane == 0) {
      const BndEntry e = ring_in[rs & mask_in];
      yin = e.x * pow2i(e.elow - elow);
    }
  }
  c_out = rs - 1;
  if (has_right) {
    // publish the (all-zero) initial state of the last column at row rs
    if (lane == 31) {
      BndEntry e;
      e.x = 0.0;
      e.elow = 0;
      e.pad = 0;
      ring_out[rs & mask_out] = e;
      asm volatile("" ::: "memory");
      st_vol(&ctl_out->written, rs);
    }
    __syncwarp();
  }

  const int Ttot = N - rs + 31;  // steps until lane 31 has done row N
  for (int t0 = 0; t0 < Ttot; t0 += LIN_B) {
    const int top0 = rs + 1 + t0 + (LIN_B - 1);  // highest row lane 0 touches in this batch
    // ---- flow control (uniform across the warp) ----
    if (!left_seed) {
      int need = top0 < N ? top0 : N;
      if (!wait_ge<false, 0>(&ctl_in->written, need, P.abort_flag, c_in)) return;
    }
    if (has_right) {
      // lane 31 will write rows up to top0-31; slot reuse needs row-RB taken
      int need = top0 - 31 - (mask_out + 1);
      if (!wait_ge<false, 0>(&ctl_out->taken, need, P.abort_flag, c_out)) return;
    }
    {
      // xring slot reuse: rows <= top0-RS must have been consumed by every consumer warp
      const int need = top0 - RS;
      if (c_cons < need) {
        const long long tw = clock64();
        unsigned spins = 0;
        for (;;) {
          int mn = ld_vol(cfin);
#pragma unroll
          for (int c = 1; c < LIN_CONS; c++) {
            int v = ld_vol(cfin + c);
            mn = v < mn ? v : mn;
          }
          c_cons = (bfirst + mn + LIN_CONS) * LIN_B - 1;
          if (c_cons >= need) break;
          if ((++spins & 1023u) == 0) {
            const int bad = ld_vol(P.abort_flag) || (clock64() - tw > LIN_WATCHDOG);
            if (__any_sync(0xffffffffu, bad)) {
              if (lane == 0) atomicExch(P.abort_flag, 1);
              return;
            }
          }
        }
        asm volatile("" ::: "memory");
      }
    }
    // ---- renormalise (uniform step) ----
    {
      const int hi = __double2hiint(x[0]);
      int e = ((hi >> 20) & 0x7ff) - 1023;
      if (x[0] == 0.0) e = 0;
      const double sc = pow2i(-e);
#pragma unroll
      for (int k = 0; k < K; k++) x[k] *= sc;
      yin *= sc;
      E += e;
      elow = (int)E;
      ering[((t0 >> 3) & (NJ - 1)) * 32 + lane] = (double)E;
    }
    // ---- eight steps ----
    const int n_lane = rs + 1 + t0 - lane;
    const bool edge = (t0 < 32) || (top0 > N);
    if (edge)
      producer_batch<K, HAS_V, RS, true>(x, yin, elow, ma, nm1, n_lane, rs, N, lane, left_seed, has_right,
                                         ring_in, mask_in, ring_out, mask_out, xring, yring);
    else
      producer_batch<K, HAS_V, RS, false>(x, yin, elow, ma, nm1, n_lane, rs, N, lane, left_seed,
                                          has_right, ring_in, mask_in, ring_out, mask_out, xring, yring);
    // ---- publish ----
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 31) {
      int done = top0 - 31;
      if (done > N) done = N;
      if (done > rs) {
        st_vol(prod_done, done);
        if (has_right) st_vol(&ctl_out->written, done);
      }
    }
    if (lane == 0 && !left_seed) {
      int tk = top0 < N ? top0 : N;
      st_vol(&ctl_in->taken, tk);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// consumer: log / divide / store, LIN_CONS warps per warp-strip
// ---------------------------------------------------------------------------------------------
template <int K, bool HAS_S, bool HAS_V, int RS, typename OutT>
__device__ void consumer_warp(const LinParams &P, int ws, int cidx, int lane, const double *xring,
                              const double *yring, const double *ering, const int *prod_done, int *cfin,
                              const LogTabEntry *logtab) {
  constexpr int W = 32 * K;
  constexpr int NJ = LinSmem<K, HAS_V, RS>::NJ;
  constexpr double EBIAS = 4503601774854144.0;  // 2^52 + 2^31
  const int N = P.N, M = P.M;
  const int m_first = ws * W + 1;
  const int rs = m_first - 1;
  const int bfirst = (rs + 1) >> 3;
  const int blast = N >> 3;
  OutT *tabS = (OutT *)P.tabS;
  OutT *tabV = (OutT *)P.tabV;
  int c_done = rs;
  for (int bb = cidx; bfirst + bb <= blast; bb += LIN_CONS) {
    const int nb = (bfirst + bb) << 3;
    int top = nb + LIN_B - 1;
    if (top > N) top = N;
    if (!wait_ge<false, 0>(prod_done, top, P.abort_flag, c_done)) return;
    const int slot0 = nb & (RS - 1);  // the block's eight ring rows are contiguous
#pragma unroll
    for (int kk = 0; kk < K; kk++) {
      const int col = lane + 32 * kk;
      const int m = m_first + col;
      const int pl = col / K;
      const int kq = col % K;
      const bool full = (nb > rs) && (nb + LIN_B - 1 <= N) && (m <= nb) && (m <= M);
      if (full) {
        // ---- steady state: eight rows, no predicates, loads first so they overlap ----
        double xv[LIN_B];
#pragma unroll
        for (int i = 0; i < LIN_B; i++) xv[i] = xring[(size_t)(slot0 + i) * W + col];
        OutT *pS = HAS_S ? tabS + (size_t)(nb - 1) * P.ld + (size_t)(m - 1) : nullptr;
        OutT *pV = HAS_V ? tabV + (size_t)(nb - 1) * P.ld + (size_t)(m - 1) : nullptr;
        if (HAS_S) {
          // lane pl computed row n at step n-rs-1+pl; its exponent changes every eighth step
          const int toff = nb - rs - 1 + pl;
          const int j0 = toff >> 3, thr = 8 - (toff & 7);
          const double Ea = ering[(j0 & (NJ - 1)) * 32 + pl] - EBIAS;
          const double Eb = ering[((j0 + 1) & (NJ - 1)) * 32 + pl] - EBIAS;
          double v[LIN_B];
#pragma unroll
          for (int i = 0; i < LIN_B; i++) v[i] = log_scaled(xv[i], (i >= thr) ? Eb : Ea, logtab);
#pragma unroll
          for (int i = 0; i < LIN_B; i++) st_out(pS + (size_t)i * P.ld, v[i]);
          if (m == 1) {
#pragma unroll
            for (int i = 0; i < LIN_B; i++) P.s1[nb - 1 + i] = v[i];
          }
        }
        if (HAS_V && m >= 2) {
          double den[LIN_B];
#pragma unroll
          for (int i = 0; i < LIN_B; i++)
            den[i] = (kq == 0) ? yring[(slot0 + i) * 32 + pl] : xring[(size_t)(slot0 + i) * W + col - 1];
#pragma unroll
          for (int i = 0; i < LIN_B; i++) st_out(pV + (size_t)i * P.ld, div_pos(xv[i], den[i]));
        }
      } else if (m <= M) {
        // ---- edges: first rows of the strip (triangle), last partial block ----
        for (int i = 0; i < LIN_B; i++) {
          const int n = nb + i;
          if (n > N || n <= rs || m > n) continue;
          const int slot = n & (RS - 1);
          const double xv = xring[(size_t)slot * W + col];
          const size_t off = (size_t)(n - 1) * P.ld + (size_t)(m - 1);
          if (HAS_S) {
            const int j = (n - rs - 1 + pl) >> 3;
            const double v = log_scaled(xv, ering[(j & (NJ - 1)) * 32 + pl] - EBIAS, logtab);
            st_out(tabS + off, v);
            if (m == 1) P.s1[n - 1] = v;
          }
          if (HAS_V && m >= 2) {
            const double den = (kq == 0) ? yring[slot * 32 + pl] : xring[(size_t)slot * W + col - 1];
            st_out(tabV + off, div_pos(xv, den));
          }
        }
      }
    }
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 0) st_vol(cfin + cidx, bb);
  }
}

// ---------------------------------------------------------------------------------------------
// loader / flusher: inter-CTA hand-off of the boundary column through an L2-resident ring.
// Both move EVERYTHING that is available per round trip, so a helper that falls behind
// amortises its global-memory latency over more rows and catches up by itself.
// ---------------------------------------------------------------------------------------------
__device__ void loader_warp(const LinParams &P, int cta, int rs0, int lane, BndEntry *ring0, RingCtl *ctl0) {
  // reads the ring written by CTA cta-1; feeds the CTA's input edge ring
  const BndEntry *g = P.gring + (size_t)(cta - 1) * LIN_RBG;
  const int last = P.N;  // rows rs0..N are published by the left CTA
  int c_w = 0, c_t = rs0 - 1, pub = rs0 - 1;
  if (lane == 0) st_release_gpu(P.gtaken + (cta - 1), rs0 - 1);
  for (int next = rs0; next <= last;) {
    int top = next + LIN_CHUNK - 1;
    if (top > last) top = last;
    if (!wait_ge<true, 0>(P.gwritten + (cta - 1), top, P.abort_flag, c_w)) return;
    if (!wait_ge<false, 0>(&ctl0->taken, top - LIN_RBX, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (hi > c_t + LIN_RBX) hi = c_t + LIN_RBX;
    for (int base = next; base <= hi; base += 128) {
      BndEntry e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int row = base + 32 * u + lane;
        if (row <= hi) e[u] = g[row & (LIN_RBG - 1)];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int row = base + 32 * u + lane;
        if (row <= hi) ring0[row & (LIN_RBX - 1)] = e[u];
      }
    }
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 0) {
      st_vol(&ctl0->written, hi);
      if (hi - pub >= LIN_RBG / 4 || hi == last) {
        st_release_gpu(P.gtaken + (cta - 1), hi);
        pub = hi;
      }
    }
    pub = __shfl_sync(0xffffffffu, pub, 0);
    next = hi + 1;
  }
}

__device__ void flusher_warp(const LinParams &P, int cta, int rsl, int lane, const BndEntry *ringG,
                             RingCtl *ctlG) {
  // drains the last strip's output edge ring into the global ring read by CTA cta+1
  BndEntry *g = P.gring + (size_t)cta * LIN_RBG;
  const int last = P.N;
  int c_w = rsl - 1, c_t = 0;
  for (int next = rsl; next <= last;) {
    int top = next + LIN_CHUNK - 1;
    if (top > last) top = last;
    if (!wait_ge<false, 0>(&ctlG->written, top, P.abort_flag, c_w)) return;
    if (!wait_ge<true, 0>(P.gtaken + cta, top - LIN_RBG, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (hi > c_t + LIN_RBG) hi = c_t + LIN_RBG;
    for (int row = next + lane; row <= hi; row += 32) g[row & (LIN_RBG - 1)] = ringG[row & (LIN_RBX - 1)];
    __syncwarp();
    if (lane == 0) {
      st_release_gpu(P.gwritten + cta, hi);  // release orders the whole warp's stores (syncwarp above)
      st_vol(&ctlG->taken, hi);
    }
    next = hi + 1;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int K, bool HAS_S, bool HAS_V, int RS, typename OutT>
__global__ void __launch_bounds__(1024, 1) fill_linear_kernel(const LinParams P) {
  using SM = LinSmem<K, HAS_V, RS>;
  constexpr int W = 32 * K;
  extern __shared__ __align__(16) unsigned char smem[];
  const int G = P.G;
  const int cta = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  LogTabEntry *logtab = reinterpret_cast<LogTabEntry *>(smem);
  unsigned char *p = smem + (((size_t)LIN_LOGTAB * sizeof(LogTabEntry) + 15) & ~(size_t)15);
  BndEntry *ring_in_edge = reinterpret_cast<BndEntry *>(p);
  p += SM::ringx_bytes;
  BndEntry *ring_out_edge = reinterpret_cast<BndEntry *>(p);
  p += SM::ringx_bytes;
  BndEntry *rings_int = reinterpret_cast<BndEntry *>(p);  // ring g (1<=g<G) at rings_int + (g-1)*RBI
  p += (size_t)(G - 1) * SM::ringi_bytes;
  unsigned char *strips = p;
  p += (size_t)G * SM::strip_bytes;
  RingCtl *ctl = reinterpret_cast<RingCtl *>(p);
  p += (size_t)(G + 1) * sizeof(RingCtl);
  int *prod_done = reinterpret_cast<int *>(p);
  p += (size_t)G * sizeof(int);
  int *cfin = reinterpret_cast<int *>(p);

  const int ws0 = cta * G;
  int nloc = P.nws - ws0;  // strips handled by this CTA
  if (nloc > G) nloc = G;
  const bool last_cta = (ws0 + nloc >= P.nws);

  for (int i = threadIdx.x; i < LIN_LOGTAB; i += blockDim.x) logtab[i] = P.logtab[i];
  if (threadIdx.x == 0) {
    for (int g = 0; g <= G; g++) {
      // ring g is written by strip g-1 (or the loader) and read by strip g (or the flusher)
      int rs_w = (ws0 + (g > 0 ? g - 1 : 0)) * W;
      int rs_r = (ws0 + (g < nloc ? g : nloc - 1)) * W;
      if (g == 0) rs_w = rs_r;
      ctl[g].written = rs_w - 1;
      ctl[g].taken = rs_r - 1;
    }
    for (int g = 0; g < G; g++) {
      prod_done[g] = (ws0 + g) * W;
      for (int c = 0; c < LIN_CONS; c++) cfin[g * LIN_CONS + c] = c - LIN_CONS;
    }
  }
  __syncthreads();

  if (warp < G) {
    const int g = warp;
    if (g < nloc) {
      const int ws = ws0 + g;
      unsigned char *sb = strips + (size_t)g * SM::strip_bytes;
      double *xring = reinterpret_cast<double *>(sb);
      double *yring = reinterpret_cast<double *>(sb + SM::xring_bytes);
      double *ering = reinterpret_cast<double *>(sb + SM::xring_bytes + SM::yring_bytes);
      const bool left_seed = (ws == 0);
      const bool has_right = !(last_cta && g == nloc - 1);
      // ring g feeds strip g, ring g+1 takes its last column; the CTA's first and last rings are
      // the large edge rings, the ones in between are small
      const BndEntry *rin = (g == 0) ? ring_in_edge : rings_int + (size_t)(g - 1) * LIN_RBI;
      const int min_ = (g == 0) ? LIN_RBX - 1 : LIN_RBI - 1;
      BndEntry *rout = (g == nloc - 1) ? ring_out_edge : rings_int + (size_t)g * LIN_RBI;
      const int mout = (g == nloc - 1) ? LIN_RBX - 1 : LIN_RBI - 1;
      producer_warp<K, HAS_V, RS>(P, ws, lane, left_seed, has_right, rin, min_, ctl + g, rout, mout,
                                  ctl + g + 1, xring, yring, ering, prod_done + g, cfin + g * LIN_CONS);
    }
  } else if (warp < G + G * LIN_CONS) {
    const int g = (warp - G) / LIN_CONS, c = (warp - G) % LIN_CONS;
    if (g < nloc) {
      unsigned char *sb = strips + (size_t)g * SM::strip_bytes;
      const double *xring = reinterpret_cast<const double *>(sb);
      const double *yring = reinterpret_cast<const double *>(sb + SM::xring_bytes);
      const double *ering = reinterpret_cast<const double *>(sb + SM::xring_bytes + SM::yring_bytes);
      consumer_warp<K, HAS_S, HAS_V, RS, OutT>(P, ws0 + g, c, lane, xring, yring, ering, prod_done + g,
                                               cfin + g * LIN_CONS, logtab);
    }
  } else if (warp == G + G * LIN_CONS) {
    if (cta > 0) loader_warp(P, cta, ws0 * W, lane, ring_in_edge, ctl);
  } else {
    if (!last_cta) flusher_warp(P, cta, (ws0 + nloc - 1) * W, lane, ring_out_edge, ctl + nloc);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct LinearState {
  BndEntry *gring;
  int *gctr;  // [2*cap] written | taken, then abort flag at [2*cap]
  int cap;    // CTA boundaries the buffers can serve
  LogTabEntry *logtab;
};

struct LinearFillArgs {
  void *tabS, *tabV;
  double *s1;
  int is_float;
  size_t ld;
  double a;
  unsigned startN, startM, N, M;
  int num_sms;
};

inline void linear_state_free(LinearState *st) {
  cudaFree(st->gring);
  cudaFree(st->gctr);
  cudaFree(st->logtab);
  st->gring = NULL;
  st->gctr = NULL;
  st->logtab = NULL;
  st->cap = 0;
}

inline size_t linear_state_bytes(const LinearState *st) {
  return st->cap ? (size_t)st->cap * LIN_RBG * sizeof(BndEntry) + (2 * (size_t)st->cap + 1) * sizeof(int) +
                       LIN_LOGTAB * sizeof(LogTabEntry)
                 : 0;
}

inline cudaError_t linear_state_prepare(LinearState *st, int nctas) {
  cudaError_t e;
  if (!st->logtab) {
    LogTabEntry h[LIN_LOGTAB];
    for (int i = 0; i < LIN_LOGTAB; i++) {
      double c = 1.0 + (double)i / 256.0;
      double inv = 1.0 / c;
      h[i].inv_c = inv;
      h[i].log_c = (double)(-logl((long double)inv));
    }
    h[0].inv_c = 1.0;
    h[0].log_c = 0.0;
    if ((e = cudaMalloc(&st->logtab, sizeof h)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(st->logtab, h, sizeof h, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  }
  if (nctas > st->cap) {
    cudaFree(st->gring);
    cudaFree(st->gctr);
    st->gring = NULL;
    st->gctr = NULL;
    st->cap = 0;
    if ((e = cudaMalloc(&st->gring, (size_t)nctas * LIN_RBG * sizeof(BndEntry))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->gctr, (2 * (size_t)nctas + 1) * sizeof(int))) != cudaSuccess) return e;
    st->cap = nctas;
  }
  return cudaSuccess;
}

template <int K, bool HAS_S, bool HAS_V, int RS, typename OutT>
inline cudaError_t launch_linear(const LinParams &P, int nctas, cudaStream_t stream) {
  using SM = LinSmem<K, HAS_V, RS>;
  const size_t smem = SM::total(P.G);
  auto kern = fill_linear_kernel<K, HAS_S, HAS_V, RS, OutT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int threads = 32 * (P.G * (1 + LIN_CONS) + 2);
  void *args[] = {(void *)&P};
  return cudaLaunchCooperativeKernel((void *)kern, dim3(nctas), dim3(threads), args, smem, stream);
}

template <int K, int RS>
inline cudaError_t dispatch_linear(const LinParams &P, int nctas, bool hasS, bool hasV, bool is_float,
                                   cudaStream_t stream) {
  if (is_float) {
    if (hasS && hasV) return launch_linear<K, true, true, RS, float>(P, nctas, stream);
    if (hasS) return launch_linear<K, true, false, RS, float>(P, nctas, stream);
    return launch_linear<K, false, true, RS, float>(P, nctas, stream);
  }
  if (hasS && hasV) return launch_linear<K, true, true, RS, double>(P, nctas, stream);
  if (hasS) return launch_linear<K, true, false, RS, double>(P, nctas, stream);
  return launch_linear<K, false, true, RS, double>(P, nctas, stream);
}

/* shared-memory budget a CTA may ask for on sm_100a (227 KB) minus a little slack */
constexpr size_t LIN_SMEM_BUDGET = 225 * 1024;

template <int K, bool HAS_V, int RS>
inline int lin_gmax() {
  int g = 0;
  while (g < 10 && LinSmem<K, HAS_V, RS>::total(g + 1) <= LIN_SMEM_BUDGET &&
         32 * ((g + 1) * (1 + LIN_CONS) + 2) <= 1024)
    g++;
  return g;
}

/*
 * Enqueue the fill on `stream`.  Returns 0, or non-zero with a message in err.
 * Geometry: smallest K in {1,2,4} whose warp-strips fit <= one CTA per SM.
 */
inline int linear_fill(LinearState *st, const LinearFillArgs &A, cudaStream_t stream, cudaEvent_t ev_end,
                       char *err, size_t errlen) {
  const bool hasS = A.tabS != NULL, hasV = A.tabV != NULL;
  if (A.N >= 0x7fffff00u || A.M > A.N) {
    snprintf(err, errlen, "linear_fill: unsupported extent N=%u M=%u", A.N, A.M);
    return -1;
  }
  int K = 0, G = 0, nws = 0;
  const int ks[3] = {1, 2, 4};
  int force_k = 0;
  if (const char *s = getenv("STB_LINEAR_K")) force_k = atoi(s);
  for (int i = 0; i < 3 && !K; i++) {
    int k = ks[i];
    if (force_k && k != force_k) continue;
    int gmax = k == 1 ? (hasV ? lin_gmax<1, true, 128>() : lin_gmax<1, false, 128>())
               : k == 2 ? (hasV ? lin_gmax<2, true, 128>() : lin_gmax<2, false, 128>())
                        : (hasV ? lin_gmax<4, true, 64>() : lin_gmax<4, false, 64>());
    int w = (int)((A.M + 32u * k - 1) / (32u * k));
    int g = (w + A.num_sms - 1) / A.num_sms;
    if (g <= gmax) {
      K = k;
      G = g;
      nws = w;
    }
  }
  if (!K) {
    snprintf(err, errlen, "linear_fill: M=%u needs more than one pass (not supported yet)", A.M);
    return -1;
  }
  const int nctas = (nws + G - 1) / G;
  cudaError_t e = linear_state_prepare(st, nctas);
  if (e == cudaSuccess) e = cudaMemsetAsync(st->gctr, 0, (2 * (size_t)st->cap + 1) * sizeof(int), stream);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "linear_fill: %s", cudaGetErrorString(e));
    return (int)e;
  }
  LinParams P;
  P.tabS = A.tabS;
  P.tabV = A.tabV;
  P.s1 = A.s1;
  P.ld = A.ld;
  P.a = A.a;
  P.N = (int)A.N;
  P.M = (int)A.M;
  P.nws = nws;
  P.G = G;
  P.gring = st->gring;
  P.gwritten = st->gctr;
  P.gtaken = st->gctr + st->cap;
  P.abort_flag = st->gctr + 2 * st->cap;
  P.logtab = st->logtab;
  if (K == 1)
    e = dispatch_linear<1, 128>(P, nctas, hasS, hasV, A.is_float != 0, stream);
  else if (K == 2)
    e = dispatch_linear<2, 128>(P, nctas, hasS, hasV, A.is_float != 0, stream);
  else
    e = dispatch_linear<4, 64>(P, nctas, hasS, hasV, A.is_float != 0, stream);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "linear_fill launch (K=%d G=%d ctas=%d): %s", K, G, nctas, cudaGetErrorString(e));
    return (int)e;
  }
  cudaEventRecord(ev_end, stream);
  // the abort flag is read back once the stream has drained
  int flag = 0;
  e = cudaMemcpyAsync(&flag, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "linear_fill (K=%d G=%d ctas=%d): %s", K, G, nctas, cudaGetErrorString(e));
    return (int)e;
  }
  if (flag) {
    snprintf(err, errlen, "linear_fill: pipeline watchdog fired (K=%d G=%d ctas=%d)", K, G, nctas);
    return -2;
  }
  return 0;
}

}  // namespace stb
