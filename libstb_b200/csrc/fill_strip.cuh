/*
 * fill_strip.cuh -- the throughput path of the table fill (sm_100a).
 *
 * What it replaces: the two double loops of S_remake_part, lib/stable.c:356-388 (log S) and
 * :451-482 (V), plus the S1 prefix :338-348.
 *
 * Formulation.  The reference iterates in log space, S' = logadd(log(c)+S_up, S_left), which
 * puts an exp->log chain (~50 dependent FP64 instructions) on the row-to-row critical path.
 * Here the SAME recurrence  S^n_m = (n-1-m a) S^{n-1}_m + S^{n-1}_{m-1}  runs in the linear
 * domain on scaled values  S = x * 2^E  (fill_common.cuh), so the critical path per row is ONE
 * DFMA; the logarithm is taken once per STORED cell by other warps.  V^n_m = S^n_m / S^n_{m-1}
 * is the ratio of two neighbouring scaled values, one division per stored cell.
 * Agreement with the reference: <= ~1e-14 relative on log S and V (tests/).
 *
 * Geometry.  The columns of a table are cut into strips of C = L*K columns, one CTA (one SM)
 * per strip, all strips of a launch co-resident (cooperative launch).  Inside a CTA:
 *   - ONE producer warp runs the recurrence: lane l owns K adjacent columns and walks down the
 *     rows in a diagonal wavefront (at step t lane l computes row t-l of the strip), so the
 *     left neighbour's value it needs was finished two steps earlier and the warp shuffle that
 *     fetches it is off the dependent chain.  Every 8 steps the lanes renormalise (uniformly).
 *     Raw x values go to a shared-memory ring indexed by row.
 *   - Consumer warps claim 8-row batches of the ring (dynamic, atomic counter), take
 *     log / divide and write each row segment to HBM exactly once with coalesced 256-byte
 *     stores.  Producer -> consumer and back is one mbarrier per batch slot (full / empty);
 *     waiting warps sleep in mbarrier.try_wait instead of spinning on shared memory.
 *   - The strip's last column goes to the right-hand neighbour CTA in batches of 8 rows with
 *     ONE exponent per batch: strips start their batch counters with a phase shift chosen so
 *     that the sender's and the receiver's batches line up, which makes the receiver's scale
 *     factor a per-batch constant.  A flusher warp copies finished batches to an L2-resident
 *     global ring (release), the neighbour's loader warp brings them into its shared memory
 *     (acquire), so no fence sits on the recurrence's critical path.
 * Several independent tables (a discount sweep) can share one launch: blockIdx.x / P selects
 * the table, blockIdx.x % P the strip.  No tensor cores: nothing here is a contraction.
 *
 * Roofline: 8 B (4 B float) written per cell, 0 B read; per S cell 2 FP64-pipe instructions of
 * recurrence + 9 of logarithm (+8 per V cell).  HBM-write bound on B200 for the FP64 table.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "fill_common.cuh"

namespace stb {

constexpr int ST_B = 8;        // steps per producer batch == rows per consumer batch
constexpr int ST_NBR = 32;     // boundary ring (batches) in shared memory
constexpr int ST_NBG = 256;    // boundary ring (batches) in global memory, per strip boundary
constexpr int ST_NJ = 32;      // batches of per-lane exponents kept
constexpr int ST_WARPS = 16;   // warps per CTA
#ifndef ST_CONS_SLEEP
#define ST_CONS_SLEEP 40       // ns a consumer sleeps between polls of its batch barrier
#endif
constexpr int ST_PRODUCER = 0, ST_LOADER = 4, ST_FLUSHER = 8;  // warp roles; all others consume

struct StripTable {  // one table of a launch
  void *tabS, *tabV;
  double *s1;
  double a;
};

struct StripParams {
  const StripTable *tables;  // device array, one per table in the launch
  unsigned long long ld;     // elements per table row
  int N, M;
  int C;       // columns per strip (= L*K)
  int L;       // producer lanes in use
  int P;       // strips per table
  double *gx;  // [boundaries][ST_NBG*8]
  int *ge;     // [boundaries][ST_NBG]
  int *gwritten, *gtaken;  // [boundaries]
  int *abort_flag;
  const LogTabEntry *logtab;
  int ncons;       // consumer warps in use (tuning knob; at most NB-1)
  long long *dbg;  // STB_PROFILE_PRODUCER builds: [ctas][8] cycle counters of the producer's phases
};

#ifdef STB_PROFILE_PRODUCER
#define ST_TICK(var) const long long var = clock64()
#define ST_ACC(slot, t1, t0) dbgacc[slot] += (t1) - (t0)
#else
#define ST_TICK(var)
#define ST_ACC(slot, t1, t0)
#endif

template <int K, bool HAS_V>
struct StripCfg {
  static constexpr int CP = 32 * K;  // row pitch of the x ring (doubles)
  // batch slots: as many as fit beside the other shared-memory users, at most 16
  static constexpr int ROW_BYTES = (CP + (HAS_V ? 32 : 0)) * 8;
  static constexpr int NB_FIT = (int)((227 * 1024 - 20 * 1024) / ROW_BYTES / ST_B) - 1;
  static constexpr int NB = NB_FIT > 16 ? 16 : NB_FIT;
  static constexpr int RS = NB * ST_B;  // ring rows; rows RS..RS+7 duplicate rows 0..7
};

// ---- shared memory -----------------------------------------------------------------------------
template <int K, bool HAS_V>
struct StripSmem {
  using Cfg = StripCfg<K, HAS_V>;
  LogTabEntry logtab[LOGTAB_N + 1];
  // raw x values indexed by producer STEP (not by row): step u of the strip sits in ring row
  // u % RS, lane l's K columns at [l*K, l*K+K).  Row r of producer lane l was made at step
  // u = r + l + phi.  Rows RS..RS+7 repeat rows 0..7 so that a consumer's eight consecutive
  // steps never wrap.
  double xring[(Cfg::RS + ST_B) * Cfg::CP];
  double yring[HAS_V ? (Cfg::RS + ST_B) * 32 : 2];
  double ering[ST_NJ * 32];           // (double)E - LOG_EBIAS per (batch, lane)
  double in_x[ST_NBR * ST_B];         // boundary from the left strip
  double out_x[ST_NBR * ST_B];        // boundary for the right strip
  int in_e[ST_NBR], out_e[ST_NBR];
  unsigned long long full[Cfg::NB], empty[Cfg::NB];
  int in_written, in_taken, out_written, out_taken;
  int next_q;
};

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

/*
 * Warp-collective blocking wait; false when the fill was aborted (watchdog / another role).
 * Polls with the non-blocking test_wait: the blocking try_wait parks the warp in hardware and was
 * measured to wake late (the hand-off then costs microseconds, not cycles).  SLEEP (ns) between
 * polls keeps waiting consumer warps from eating the issue slots of working ones.
 */
template <int SLEEP>
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity, int *abort_flag) {
  if (mbar_test_wait(bar, parity)) return true;
  const long long t0 = clock64();
  unsigned spins = 0;
  for (;;) {
    if (SLEEP) __nanosleep(SLEEP);
    if (mbar_test_wait(bar, parity)) return true;
    if ((++spins & 1023u) == 0) {
      const int bad = ld_vol(abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
      if (__any_sync(0xffffffffu, bad)) {
        if ((threadIdx.x & 31) == 0) atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
}

/*
 * Warp-collective wait until *ctr >= need (shared-memory or global counter).  Shared-memory
 * counters and the data they guard are accessed in issue order by the SM, so volatile accesses
 * plus compiler barriers suffice inside a CTA; the global hand-off uses release/acquire.
 */
template <bool GLOBAL, int SLEEP>
__device__ __forceinline__ bool ctr_wait(const int *ctr, int need, int *abort_flag, int &cached) {
  if (cached >= need) return true;
  int v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_vol(ctr);
  if (v < need) {
    const long long t0 = clock64();
    unsigned spins = 0;
    do {
      if (SLEEP) __nanosleep(SLEEP);
      v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_vol(ctr);
      if ((++spins & 1023u) == 0) {
        const int bad = ld_vol(abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
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

// ---- strip geometry shared by the roles ------------------------------------------------------------
struct StripGeom {
  int rs;      // columns rs+1 .. rs+C ; rows n = rs+1 .. N are r = 0 .. R-1
  int R;       // rows of the strip
  int phi;     // batch phase: step t sits at position (t+phi)&7 of batch (t+phi)>>3
  int nbatch;  // producer batches
  int QT;      // consumer row batches
  int D;       // row batch q is complete after producer batch q+D
  int delta;   // producer batch p reads the left strip's batch p+delta
};

__device__ __forceinline__ int strip_phi(int strip, int L, int C) {
  return (int)(((long long)strip * (long long)(L - 1 + C)) & 7);
}

__device__ __forceinline__ StripGeom strip_geom(const StripParams &P, int strip) {
  StripGeom g;
  g.rs = strip * P.C;
  g.R = P.N - g.rs;
  g.phi = strip_phi(strip, P.L, P.C);
  g.nbatch = ((g.R - 1 + P.L - 1 + g.phi) >> 3) + 1;
  g.QT = ((g.R - 1) >> 3) + 1;
  g.D = (g.phi + P.L + 6) >> 3;
  g.delta = strip > 0 ? (P.L - 1 + P.C + strip_phi(strip - 1, P.L, P.C) - g.phi) >> 3 : 0;
  return g;
}

// ---- producer ----------------------------------------------------------------------------------------
/*
 * Eight recurrence steps.  x[k]: the lane's K columns; the coefficient of column m_k in row n is
 * (n-1) - m_k a, formed per step from nm1 = n-1 and ma[k] = m_k a exactly like that (one rounding,
 * independent of where a strip starts, so every geometry produces the same bits); yin: the left neighbour's value one row up, in this lane's units;
 * scn: the per-batch factor that brings the neighbour's units to this lane's.  Everything is
 * stored unconditionally: rows that do not exist (before the strip's first row, past N) land in
 * ring slots the consumers never read as valid.
 */
template <int K, bool HAS_V, bool DUP, int CP, int RS>
__device__ __forceinline__ void strip_steps(double (&x)[K], const double (&ma)[K], double &nm1, double &yin,
                                            const double scn,
                                            const double (&bnd)[ST_B], const bool lane0, const bool write_out,
                                            double *__restrict__ xr, double *__restrict__ yr,
                                            double *__restrict__ outp) {
#pragma unroll
  for (int i = 0; i < ST_B; i++) {
    // the left neighbour's last column BEFORE this step's update: its value one row up from
    // the row this lane makes in the NEXT step
    double s = shfl_up_d(x[K - 1]);
    if (lane0) s = bnd[i];
#pragma unroll
    for (int k = K - 1; k >= 1; k--) x[k] = fma(nm1 - ma[k], x[k], x[k - 1]);
    x[0] = fma(nm1 - ma[0], x[0], yin);
    yin = s * scn;
    nm1 += 1.0;
    if (K == 2) {
      *reinterpret_cast<double2 *>(xr + i * CP) = make_double2(x[0], x[1]);
      if (DUP) *reinterpret_cast<double2 *>(xr + (RS + i) * CP) = make_double2(x[0], x[1]);
    } else {
#pragma unroll
      for (int k = 0; k < K; k++) {
        xr[i * CP + k] = x[k];
        if (DUP) xr[(RS + i) * CP + k] = x[k];
      }
    }
    if (HAS_V) {
      yr[i * 32] = yin;
      if (DUP) yr[(RS + i) * 32] = yin;
    }
    if (write_out) outp[i] = x[K - 1];
  }
}

template <int K, bool HAS_V>
__device__ void strip_producer(const StripParams &P, StripSmem<K, HAS_V> &sm, const StripGeom &g, int lane,
                               double a, bool has_left, bool has_right, int jlast) {
  using Cfg = StripCfg<K, HAS_V>;
  constexpr int CP = Cfg::CP, RS = Cfg::RS, NB = Cfg::NB;
  const int L = P.L;

  double x[K], ma[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    x[k] = 0.0;
    ma[k] = (double)(g.rs + 1 + lane * K + k) * a;
  }
  // step u = 0 (t = -phi) is the first of batch 0; lane l is then at row r = -phi - l, n = rs+1+r
  double nm1 = (double)(g.rs - g.phi - lane);  // n - 1
  // S^rs_rs = 1 (S^0_0 = 1 for the first strip) seeds the strip's diagonal; with phi > 0 it
  // arrives through the boundary ring like every other row
  double yin = (lane == 0 && (g.phi == 0 || !has_left)) ? 1.0 : 0.0;
  long long E = 0;
  int elow = 0;
  int c_in = -1, c_out = -1;
  const bool lane0 = (lane == 0);
  const bool write_out = has_right && (lane == L - 1);
  bool slot_free = false;

#ifdef STB_PROFILE_PRODUCER
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  for (int p = 0; p < g.nbatch; ++p) {
    const int slot = p % NB;
    ST_TICK(tk0);
    // ---- flow control (uniform across the warp) ----
    // the ring slot of this batch was probed (non-blocking) one batch ago; consumers are normally
    // far ahead, so the blocking wait and its latency are rarely needed
    if (p >= NB && !slot_free) {
      if (!mbar_wait<0>(&sm.empty[slot], (unsigned)((p / NB) - 1) & 1u, P.abort_flag)) return;
    }
    if (p + 1 >= NB) slot_free = mbar_test_wait(&sm.empty[(p + 1) % NB], (unsigned)(((p + 1) / NB) - 1) & 1u);
    ST_TICK(tk1);
    int jb = p + g.delta;
    if (has_left) {
      if (jb > jlast) jb = jlast;
      if (!ctr_wait<false, 0>(&sm.in_written, jb, P.abort_flag, c_in)) return;
    }
    if (has_right) {
      if (!ctr_wait<false, 0>(&sm.out_taken, p - ST_NBR, P.abort_flag, c_out)) return;
    }
    ST_TICK(tk2);
    // ---- renormalise ----
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
      sm.ering[(p & (ST_NJ - 1)) * 32 + lane] = (double)E - LOG_EBIAS;
    }
    // scale that brings the left neighbour's values into this lane's units, fixed for the batch
    int sE = __shfl_up_sync(0xffffffffu, elow, 1);
    double bnd[ST_B];
    if (has_left) {
      const int bs = jb & (ST_NBR - 1);
      if (lane0) sE = sm.in_e[bs];
      const double2 *bp = reinterpret_cast<const double2 *>(&sm.in_x[bs * ST_B]);
#pragma unroll
      for (int i = 0; i < ST_B / 2; i++) {
        const double2 v = bp[i];
        bnd[2 * i] = v.x;
        bnd[2 * i + 1] = v.y;
      }
    } else {
      if (lane0) sE = elow;
#pragma unroll
      for (int i = 0; i < ST_B; i++) bnd[i] = 0.0;
    }
    const double scn = pow2i(sE - elow);
    if (write_out) sm.out_e[p & (ST_NBR - 1)] = elow;
    double *outp = &sm.out_x[(p & (ST_NBR - 1)) * ST_B];
    double *xr = &sm.xring[slot * ST_B * CP + lane * K];
    double *yr = &sm.yring[HAS_V ? slot * ST_B * 32 + lane : 0];

    ST_TICK(tk3);
    // ---- eight steps ----
    if (slot == 0)
      strip_steps<K, HAS_V, true, CP, RS>(x, ma, nm1, yin, scn, bnd, lane0, write_out, xr, yr, outp);
    else
      strip_steps<K, HAS_V, false, CP, RS>(x, ma, nm1, yin, scn, bnd, lane0, write_out, xr, yr, outp);

    ST_TICK(tk4);
    // ---- publish ----
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane0) {
      const int q = p - g.D;
      if (q >= 0 && q < g.QT) mbar_arrive(&sm.full[q % NB]);
      if (has_left) st_vol(&sm.in_taken, p + g.delta);
    }
    if (write_out) st_vol(&sm.out_written, p);
    ST_TICK(tk5);
    ST_ACC(0, tk1, tk0);  // wait: ring slot free
    ST_ACC(1, tk2, tk1);  // wait: boundary in / out
    ST_ACC(2, tk3, tk2);  // renormalise + batch set-up
    ST_ACC(3, tk4, tk3);  // eight steps
    ST_ACC(4, tk5, tk4);  // publish
  }
#ifdef STB_PROFILE_PRODUCER
  if (lane0 && P.dbg) {
    for (int i = 0; i < 5; i++) P.dbg[(size_t)blockIdx.x * 8 + i] = dbgacc[i];
    P.dbg[(size_t)blockIdx.x * 8 + 5] = g.nbatch;
  }
#endif
  // a partial last row batch never sees its eighth row: release it now that every row exists
  if (lane0) {
    const int qd = g.nbatch - 1 - g.D;
    for (int q = (qd < 0 ? 0 : qd + 1); q < g.QT; q++) mbar_arrive(&sm.full[q % NB]);
  }
}

// ---- consumer ----------------------------------------------------------------------------------------
template <int K, bool HAS_S, bool HAS_V, typename OutT>
__device__ void strip_consumer(const StripParams &P, StripSmem<K, HAS_V> &sm, const StripGeom &g, int lane,
                               const StripTable &tb) {
  using Cfg = StripCfg<K, HAS_V>;
  constexpr int CP = Cfg::CP, NB = Cfg::NB, RS = Cfg::RS;
  const int M = P.M;
  int cvalid = M - g.rs;  // columns of this strip that exist
  if (cvalid > P.C) cvalid = P.C;
  OutT *tabS = (OutT *)tb.tabS;
  OutT *tabV = (OutT *)tb.tabV;
  const bool first_strip = (g.rs == 0);

  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&sm.next_q, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= g.QT) break;
    const int slot = q % NB;
    if (!mbar_wait<ST_CONS_SLEEP>(&sm.full[slot], (unsigned)(q / NB) & 1u, P.abort_flag)) return;
    const int r0 = q * ST_B;
    const bool fast = (r0 >= cvalid - 1) && (r0 + ST_B <= g.R);
    // not unrolled: the body is ~250 instructions; K copies of it overflow the instruction cache
    // and the misses stall every warp of the SM, the producer included
#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
      if (32 * kk >= cvalid) break;
      const int col = lane + 32 * kk;
      const int pl = col / K;
      const int kq = col % K;
      const bool lane_ok = col < cvalid;
      const size_t cell0 = (size_t)(g.rs + r0) * P.ld + (size_t)(g.rs + col);  // row n-1 = rs+r, column m-1
      // producer lane pl made rows r0..r0+7 at steps toff..toff+7 (ring rows ub..ub+7, never
      // wrapping thanks to the duplicate rows); its exponent changes once on the way
      const int toff = r0 + pl + g.phi;
      const int j0 = toff >> 3, thr = ST_B - (toff & 7);
      const int ub = (j0 % NB) * ST_B + (toff & 7);
      const double *xc = &sm.xring[ub * CP + col];
      const double *yc = &sm.yring[HAS_V ? ub * 32 + pl : 0];
      if (fast) {
        double xv[ST_B];
#pragma unroll
        for (int i = 0; i < ST_B; i++) xv[i] = xc[i * CP];
        if (HAS_S) {
          const double Ea = sm.ering[(j0 & (ST_NJ - 1)) * 32 + pl];
          const double Eb = sm.ering[((j0 + 1) & (ST_NJ - 1)) * 32 + pl];
          double v[ST_B];
#pragma unroll
          for (int i = 0; i < ST_B; i++) v[i] = log_scaled(xv[i], (i >= thr) ? Eb : Ea, sm.logtab);
          if (lane_ok) {
            OutT *pS = tabS + cell0;
#pragma unroll
            for (int i = 0; i < ST_B; i++) st_out(pS + (size_t)i * P.ld, v[i]);
            if (first_strip && col == 0) {
#pragma unroll
              for (int i = 0; i < ST_B; i++) tb.s1[r0 + i] = v[i];
            }
          }
        }
        if (HAS_V) {
          double den[ST_B];
#pragma unroll
          for (int i = 0; i < ST_B; i++)
            den[i] = (kq == 0) ? yc[i * 32] : xc[i * CP - 1];
          if (lane_ok && !(first_strip && col == 0)) {
            OutT *pV = tabV + cell0;
#pragma unroll
            for (int i = 0; i < ST_B; i++) st_out(pV + (size_t)i * P.ld, div_pos(xv[i], den[i]));
          }
        }
      } else if (lane_ok) {
        // ---- edges: the strip's triangle (column col exists from row r = col on), last rows ----
        for (int i = 0; i < ST_B; i++) {
          const int r = r0 + i;
          if (r >= g.R || col > r) continue;
          const double xv = xc[i * CP];
          const size_t off = cell0 + (size_t)i * P.ld;
          if (HAS_S) {
            const int j = (r + pl + g.phi) >> 3;
            const double v = log_scaled(xv, sm.ering[(j & (ST_NJ - 1)) * 32 + pl], sm.logtab);
            st_out(tabS + off, v);
            if (first_strip && col == 0) tb.s1[r] = v;
          }
          if (HAS_V && !(first_strip && col == 0)) {
            const double den = (kq == 0) ? yc[i * 32] : xc[i * CP - 1];
            st_out(tabV + off, div_pos(xv, den));
          }
        }
      }
    }
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 0) mbar_arrive(&sm.empty[slot]);
  }
}

// ---- loader / flusher: the strip boundary through an L2-resident ring --------------------------------
template <int K, bool HAS_V>
__device__ void strip_loader(const StripParams &P, StripSmem<K, HAS_V> &sm, const StripGeom &g, int lane,
                             int bidx, int jlast) {
  // boundary bidx is written by the strip to the left; batches delta .. jlast are needed here
  const double *gx = P.gx + (size_t)bidx * (ST_NBG * ST_B);
  const int *ge = P.ge + (size_t)bidx * ST_NBG;
  int c_w = -1, c_t = -1, pub = g.delta - 1;
  if (lane == 0) st_release_gpu(P.gtaken + bidx, g.delta - 1);
  for (int next = g.delta; next <= jlast;) {
    if (!ctr_wait<true, 64>(P.gwritten + bidx, next, P.abort_flag, c_w)) return;
    if (!ctr_wait<false, 64>(&sm.in_taken, next - ST_NBR, P.abort_flag, c_t)) return;
    int hi = c_w < jlast ? c_w : jlast;
    if (hi > c_t + ST_NBR) hi = c_t + ST_NBR;
    // x: 8 doubles per batch, lane i&7 of group i>>3 ; four batches per pass
    for (int b = next; b <= hi; b += 4) {
      const int j = b + (lane >> 3);
      if (j <= hi) sm.in_x[(j & (ST_NBR - 1)) * ST_B + (lane & 7)] = __ldcg(&gx[(size_t)(j & (ST_NBG - 1)) * ST_B + (lane & 7)]);
    }
    for (int j = next + lane; j <= hi; j += 32) sm.in_e[j & (ST_NBR - 1)] = __ldcg(&ge[j & (ST_NBG - 1)]);
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 0) {
      st_vol(&sm.in_written, hi);
      if (hi - pub >= ST_NBG / 4 || hi == jlast) {
        st_release_gpu(P.gtaken + bidx, hi);
        pub = hi;
      }
    }
    pub = __shfl_sync(0xffffffffu, pub, 0);
    next = hi + 1;
  }
}

template <int K, bool HAS_V>
__device__ void strip_flusher(const StripParams &P, StripSmem<K, HAS_V> &sm, const StripGeom &g, int lane,
                              int bidx) {
  double *gx = P.gx + (size_t)bidx * (ST_NBG * ST_B);
  int *ge = P.ge + (size_t)bidx * ST_NBG;
  const int last = g.nbatch - 1;
  int c_w = -1, c_t = -1;
  for (int next = 0; next <= last;) {
    if (!ctr_wait<false, 32>(&sm.out_written, next, P.abort_flag, c_w)) return;
    if (!ctr_wait<true, 64>(P.gtaken + bidx, next - ST_NBG, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (hi > c_t + ST_NBG) hi = c_t + ST_NBG;
    for (int b = next; b <= hi; b += 4) {
      const int j = b + (lane >> 3);
      if (j <= hi) gx[(size_t)(j & (ST_NBG - 1)) * ST_B + (lane & 7)] = sm.out_x[(j & (ST_NBR - 1)) * ST_B + (lane & 7)];
    }
    for (int j = next + lane; j <= hi; j += 32) ge[j & (ST_NBG - 1)] = sm.out_e[j & (ST_NBR - 1)];
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      st_release_gpu(P.gwritten + bidx, hi);
      st_vol(&sm.out_taken, hi);
    }
    next = hi + 1;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
template <int K, bool HAS_S, bool HAS_V, typename OutT>
__global__ void __launch_bounds__(ST_WARPS * 32, 1) fill_strip_kernel(const StripParams P) {
  using SM = StripSmem<K, HAS_V>;
  using Cfg = StripCfg<K, HAS_V>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &sm = *reinterpret_cast<SM *>(smem_raw);
  const int table = blockIdx.x / P.P, strip = blockIdx.x % P.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const StripGeom g = strip_geom(P, strip);
  const StripTable tb = P.tables[table];
  const bool has_left = strip > 0, has_right = strip + 1 < P.P;

  for (int i = threadIdx.x; i < LOGTAB_N; i += blockDim.x) sm.logtab[i] = P.logtab[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NB; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    sm.in_written = -1;
    sm.in_taken = -1;
    sm.out_written = -1;
    sm.out_taken = -1;
    sm.next_q = 0;
  }
  __syncthreads();

  // last batch the strip to the left produces
  const int jlast = has_left ? strip_geom(P, strip - 1).nbatch - 1 : 0;
  if (warp == ST_PRODUCER) {
    strip_producer<K, HAS_V>(P, sm, g, lane, tb.a, has_left, has_right, jlast);
  } else if (warp == ST_LOADER) {
    if (has_left) strip_loader<K, HAS_V>(P, sm, g, lane, table * P.P + strip - 1, jlast);
  } else if (warp == ST_FLUSHER) {
    if (has_right) strip_flusher<K, HAS_V>(P, sm, g, lane, table * P.P + strip);
  } else {
    // a waiter may be at most one phase ahead of its mbarrier: fewer claimants than batch slots
    // consumers stay off the producer's SM sub-partition (warp % 4 == 0): it keeps its issue slots
    // and FP64 pipe to itself
    const int cidx = (warp >> 2) * 3 + (warp & 3) - 1;
    if ((warp & 3) != 0 && cidx < Cfg::NB - 1 && cidx < P.ncons) strip_consumer<K, HAS_S, HAS_V, OutT>(P, sm, g, lane, tb);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct StripState {
  double *gx;
  int *ge;
  int *gctr;  // [2*cap] written | taken, then the abort flag at [2*cap]
  int cap;    // strip boundaries the buffers can serve
  LogTabEntry *logtab;
  StripTable *tables;  // device array
  int tables_cap;
  int ncons;       // consumer warps in use (tuning knob; at most NB-1)
  long long *dbg;  // STB_PROFILE_PRODUCER builds only
};

inline void strip_state_free(StripState *st) {
  cudaFree(st->gx);
  cudaFree(st->ge);
  cudaFree(st->gctr);
  cudaFree(st->logtab);
  cudaFree(st->tables);
  cudaFree(st->dbg);
  memset(st, 0, sizeof *st);
}

inline size_t strip_state_bytes(const StripState *st) {
  return (size_t)st->cap * ST_NBG * (ST_B * sizeof(double) + sizeof(int)) +
         (st->cap ? (2 * (size_t)st->cap + 1) * sizeof(int) : 0) + (st->logtab ? LOGTAB_N * sizeof(LogTabEntry) : 0) +
         (size_t)st->tables_cap * sizeof(StripTable);
}

inline cudaError_t strip_state_prepare(StripState *st, int nbound, int ntables) {
  cudaError_t e;
  if (!st->logtab) {
    LogTabEntry h[LOGTAB_N];
    logtab_host(h);
    if ((e = cudaMalloc(&st->logtab, sizeof h)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(st->logtab, h, sizeof h, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  }
  if (nbound > st->cap) {
    cudaFree(st->gx);
    cudaFree(st->ge);
    cudaFree(st->gctr);
    st->gx = NULL;
    st->ge = NULL;
    st->gctr = NULL;
    st->cap = 0;
    if ((e = cudaMalloc(&st->gx, (size_t)nbound * ST_NBG * ST_B * sizeof(double))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->ge, (size_t)nbound * ST_NBG * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->gctr, (2 * (size_t)nbound + 1) * sizeof(int))) != cudaSuccess) return e;
    st->cap = nbound;
  }
  if (ntables > st->tables_cap) {
    cudaFree(st->tables);
    st->tables = NULL;
    st->tables_cap = 0;
    if ((e = cudaMalloc(&st->tables, (size_t)ntables * sizeof(StripTable))) != cudaSuccess) return e;
    st->tables_cap = ntables;
  }
  return cudaSuccess;
}

template <int K, bool HAS_S, bool HAS_V, typename OutT>
inline cudaError_t launch_strip(const StripParams &P, int nctas, cudaStream_t stream) {
  const size_t smem = sizeof(StripSmem<K, HAS_V>);
  auto kern = fill_strip_kernel<K, HAS_S, HAS_V, OutT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void *args[] = {(void *)&P};
  return cudaLaunchCooperativeKernel((void *)kern, dim3(nctas), dim3(ST_WARPS * 32), args, smem, stream);
}

template <int K>
inline cudaError_t dispatch_strip(const StripParams &P, int nctas, bool hasS, bool hasV, bool is_float,
                                  cudaStream_t stream) {
  if (is_float) {
    if (hasS && hasV) return launch_strip<K, true, true, float>(P, nctas, stream);
    if (hasS) return launch_strip<K, true, false, float>(P, nctas, stream);
    return launch_strip<K, false, true, float>(P, nctas, stream);
  }
  if (hasS && hasV) return launch_strip<K, true, true, double>(P, nctas, stream);
  if (hasS) return launch_strip<K, true, false, double>(P, nctas, stream);
  return launch_strip<K, false, true, double>(P, nctas, stream);
}

struct StripPlan {
  int K, L, C, P;  // columns per lane, lanes, columns per strip, strips per table
};

/*
 * Geometry for tables of M columns when `slots` CTAs are available per table.  A strip narrower
 * than ~96 columns is bound by the latency of one recurrence step, not by throughput, so strips
 * are at least that wide (fewer hand-offs for small tables); beyond that the columns are spread
 * over all slots.  K is the smallest of {1,2,3,5,7} that reaches the width (odd K and K=2 store
 * to the x ring without bank conflicts); the lane count is trimmed so that strip edges fall on
 * 32-byte sectors of the table.
 */
inline bool strip_plan(unsigned M, int slots, size_t elem_size, StripPlan *pl) {
  static const int ks[5] = {1, 2, 3, 5, 7};
  int force_k = 0, force_l = 0;
  if (const char *s = getenv("STB_STRIP_K")) force_k = atoi(s);
  if (const char *s = getenv("STB_STRIP_L")) force_l = atoi(s);
  if (slots < 1) slots = 1;
  unsigned want = (M + (unsigned)slots - 1) / (unsigned)slots;  // columns per strip with every slot in use
  if (want < 96) want = 96;
  if (want > M) want = M;
  for (int i = 0; i < 5; i++) {
    const int k = ks[i];
    if (force_k ? k != force_k : 32u * (unsigned)k < want) continue;
    int L = (int)((want + (unsigned)k - 1) / (unsigned)k);
    if (L > 32) L = 32;
    while (L < 32 && ((size_t)L * k * elem_size) % 32 != 0) L++;
    if (force_l) L = force_l;
    pl->K = k;
    pl->L = L;
    pl->C = L * k;
    pl->P = (int)((M + (unsigned)pl->C - 1) / (unsigned)pl->C);
    if (pl->P > slots) continue;
    return true;
  }
  return false;
}

struct StripFillArgs {
  const StripTable *tables;  // HOST array of ntables entries
  int ntables;
  int has_S, has_V, is_float;
  size_t ld;
  unsigned N, M;
  int num_sms;
};

/*
 * Enqueue the fill of `ntables` tables of identical extent on `stream` and wait for it.
 * Returns 0, or non-zero with a message in err.  ev_end is recorded right after the last kernel.
 */
inline int strip_fill(StripState *st, const StripFillArgs &A, cudaStream_t stream, cudaEvent_t ev_end, char *err,
                      size_t errlen) {
  if (A.N >= 0x7fffff00u || A.M > A.N || A.M < 1) {
    snprintf(err, errlen, "strip_fill: unsupported extent N=%u M=%u", A.N, A.M);
    return -1;
  }
  // tables per launch: as many as fit when each gets at least enough CTAs for K=7 strips
  StripPlan pl;
  int per_launch = 1;
  if (A.ntables > 1) {
    // widest strips first: fewest CTAs per table, most tables in flight
    int pmin = (int)((A.M + 32u * 7u - 1) / (32u * 7u));
    per_launch = A.num_sms / (pmin > 0 ? pmin : 1);
    if (per_launch < 1) per_launch = 1;
    if (per_launch > A.ntables) per_launch = A.ntables;
  }
  if (!strip_plan(A.M, A.num_sms / per_launch, A.is_float ? 4 : 8, &pl)) {
    snprintf(err, errlen, "strip_fill: M=%u needs more than one pass over the columns (not supported)", A.M);
    return -1;
  }
  cudaError_t e = strip_state_prepare(st, per_launch * pl.P, A.ntables);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(st->tables, A.tables, (size_t)A.ntables * sizeof(StripTable), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "strip_fill: %s", cudaGetErrorString(e));
    return (int)e;
  }
  StripParams P;
  P.ld = A.ld;
  P.N = (int)A.N;
  P.M = (int)A.M;
  P.C = pl.C;
  P.L = pl.L;
  P.P = pl.P;
  P.gx = st->gx;
  P.ge = st->ge;
  P.gwritten = st->gctr;
  P.gtaken = st->gctr + st->cap;
  P.abort_flag = st->gctr + 2 * st->cap;
  P.logtab = st->logtab;
  P.ncons = 64;
  if (const char *s = getenv("STB_STRIP_CONS")) P.ncons = atoi(s);
  P.dbg = NULL;
#ifdef STB_PROFILE_PRODUCER
  if (!st->dbg) cudaMalloc(&st->dbg, 256 * 8 * sizeof(long long));
  cudaMemsetAsync(st->dbg, 0, 256 * 8 * sizeof(long long), stream);
  P.dbg = st->dbg;
#endif
  for (int t0 = 0; t0 < A.ntables && e == cudaSuccess; t0 += per_launch) {
    const int nt = (A.ntables - t0 < per_launch) ? A.ntables - t0 : per_launch;
    P.tables = st->tables + t0;
    // counters start at -1 ("nothing written / taken"), the abort flag at 0
    e = cudaMemsetAsync(st->gctr, 0xFF, 2 * (size_t)st->cap * sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.abort_flag, 0, sizeof(int), stream);
    if (e != cudaSuccess) break;
    const int nctas = nt * pl.P;
    switch (pl.K) {
      case 1: e = dispatch_strip<1>(P, nctas, A.has_S, A.has_V, A.is_float != 0, stream); break;
      case 2: e = dispatch_strip<2>(P, nctas, A.has_S, A.has_V, A.is_float != 0, stream); break;
      case 3: e = dispatch_strip<3>(P, nctas, A.has_S, A.has_V, A.is_float != 0, stream); break;
      case 5: e = dispatch_strip<5>(P, nctas, A.has_S, A.has_V, A.is_float != 0, stream); break;
      default: e = dispatch_strip<7>(P, nctas, A.has_S, A.has_V, A.is_float != 0, stream); break;
    }
    if (e == cudaSuccess) {
      // the abort flag is checked per launch: a later memset must not hide it
      int flag = 0;
      if (t0 + per_launch >= A.ntables) cudaEventRecord(ev_end, stream);
      e = cudaMemcpyAsync(&flag, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
#ifdef STB_PROFILE_PRODUCER
      if (e == cudaSuccess && getenv("STB_PROFILE_PRINT")) {
        static long long h[256 * 8];
        cudaMemcpy(h, st->dbg, sizeof h, cudaMemcpyDeviceToHost);
        const int show[4] = {0, 1, nctas / 2, nctas - 1};
        for (int si = 0; si < 4; si++) {
          const int c = show[si];
          if (c < 0 || c >= nctas || c >= 256 || (si && c == show[si - 1])) continue;
          const double nb = (double)h[c * 8 + 5];
          fprintf(stderr, "cta %3d batches %6.0f cycles/batch: slot-wait %.0f ring-wait %.0f setup %.0f steps %.0f publish %.0f\n",
                  c, nb, h[c * 8 + 0] / nb, h[c * 8 + 1] / nb, h[c * 8 + 2] / nb, h[c * 8 + 3] / nb, h[c * 8 + 4] / nb);
        }
      }
#endif
      if (e == cudaSuccess && flag) {
        snprintf(err, errlen, "strip_fill: pipeline watchdog fired (K=%d L=%d P=%d tables/launch=%d)", pl.K, pl.L,
                 pl.P, per_launch);
        return -2;
      }
    }
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "strip_fill (K=%d L=%d P=%d tables/launch=%d): %s", pl.K, pl.L, pl.P, per_launch,
             cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace stb
