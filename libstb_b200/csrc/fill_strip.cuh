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
 * Geometry.  The columns of a table are cut into strips of Cw = L*K columns.  A CTA (one per SM,
 * all CTAs of a launch co-resident: cooperative launch) owns G adjacent strips.  Per strip:
 *   - ONE producer warp runs the recurrence: lane l owns K adjacent columns and walks down the
 *     rows in a diagonal wavefront (at step t lane l makes row t-l of the strip), so the left
 *     neighbour's value it needs was finished two steps earlier and the warp shuffle that
 *     fetches it is off the dependent chain.  Every 16 steps (a batch) the lanes renormalise.
 *     Raw x values go to a shared-memory ring indexed by STEP, so a step's stores need no
 *     address arithmetic; the G producers of a CTA sit on different SM sub-partitions.
 *   - Consumer warps claim 8-row batches (atomic counter per strip), take log / divide and write
 *     each row segment to HBM exactly once with coalesced 256-byte stores.  Producer -> consumer
 *     is one mbarrier per batch slot; consumer -> producer a generation word per slot.
 *   - The strip's last column goes to the next strip in batches of 16 rows with ONE exponent per
 *     batch: strips start their batch counters with a phase shift chosen so that the sender's
 *     and the receiver's batches line up, which makes the receiver's scale factor a per-batch
 *     constant.  Inside a CTA the hand-off is a shared-memory ring; between CTAs a flusher warp
 *     copies finished batches to an L2-resident global ring (release) and the neighbour's loader
 *     warp brings them into its shared memory (acquire), so no fence sits on the recurrence's
 *     critical path.
 * Several independent tables (a discount sweep) can share one launch: blockIdx.x / ctas_per_table
 * selects the table.  No tensor cores: nothing here is a contraction.
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

constexpr int ST_SH = 4;
constexpr int ST_B = 1 << ST_SH;  // steps per producer batch (one renormalisation, one boundary exponent)
constexpr int ST_RB = 8;          // rows per consumer batch
#ifndef ST_NBR_OVERRIDE
#define ST_NBR_OVERRIDE 16
#endif
constexpr int ST_NBR = ST_NBR_OVERRIDE;  // boundary ring (producer batches) in shared memory
constexpr int ST_NBG = 128;       // boundary ring (producer batches) in global memory, per CTA boundary
constexpr int ST_NJ = 16;         // producer batches of per-lane exponents kept
constexpr int ST_WARPS = 16;      // warps per CTA
constexpr int ST_LOADER = 4, ST_FLUSHER = 8;  // helper warps; producers are warps 0..G-1, the rest consume
#ifndef ST_CONS_SLEEP
#define ST_CONS_SLEEP 40  // ns a consumer sleeps between polls of its batch barrier
#endif

struct StripTable {  // one table of a launch
  void *tabS, *tabV;
  double *s1;
  double a;
};

struct StripParams {
  const StripTable *tables;  // device array, one per table in the launch
  unsigned long long ld;     // elements per table row
  int N, M;
  int C;     // columns per strip (= L*K)
  int L;     // producer lanes in use
  int P;     // strips per table
  int ctas;  // CTAs per table (= ceil(P / G))
  uint4 *gring;  // [boundaries][ST_NBG*(ST_B+1)] flag-in-data entries, zeroed before each launch
  int *gtaken;   // [boundaries] reader progress (back-pressure only)
  int *abort_flag;
  const LogTabEntry *logtab;
  int ncons;       // consumer warps per strip in use (tuning knob)
  int spread;      // consumer warps allowed on each producer's SM sub-partition (0..3)
  long long *dbg;  // STB_PROFILE_PRODUCER builds: [ctas][8] cycle counters of the producer's phases
};

#ifdef STB_PROFILE_PRODUCER
#define ST_TICK(var) const long long var = clock64()
#define ST_ACC(slot, t1, t0) dbgacc[slot] += (t1) - (t0)
#else
#define ST_TICK(var)
#define ST_ACC(slot, t1, t0)
#endif

template <int K, int G, bool HAS_V>
struct StripCfg {
  static constexpr int CP = 32 * K;  // row pitch of the x ring (doubles)
  static constexpr int ROW_BYTES = (CP + (HAS_V ? 32 : 0)) * 8;
  // consumer batch slots per strip: as many as fit beside the other shared-memory users, an
  // even number (a producer batch is two of them), at most 16
  static constexpr int FIXED = 4352 + (G + 1) * (ST_NBR * ST_B * 8 + ST_NBR * 4 + 16) + G * (ST_NJ * 32 * 4 + 512);
  static constexpr int NB_FIT = ((227 * 1024 - FIXED) / G / ROW_BYTES / ST_RB - 1) & ~1;
  static constexpr int NB = NB_FIT > 16 ? 16 : NB_FIT;
  static constexpr int RS = NB * ST_RB;  // ring rows; rows RS..RS+7 duplicate rows 0..7
  static_assert(NB >= 6, "x ring too small for this geometry");
};

// ---- shared memory -----------------------------------------------------------------------------
struct alignas(16) BRing {  // boundary column between two strips, in producer batches
  double x[ST_NBR * ST_B];
  int e[ST_NBR];
  int written;  // batches <= written are valid
  int taken;    // the reader is done with batches <= taken
  int pad[2];
};

template <int K, int G, bool HAS_V>
struct alignas(16) StripSub {  // per strip
  using Cfg = StripCfg<K, G, HAS_V>;
  // raw x values indexed by producer STEP (not by row): step u of the strip sits in ring row
  // u % RS, lane l's K columns at [l*K, l*K+K).  Row r of producer lane l was made at step
  // u = r + l + phi.  Rows RS..RS+7 repeat rows 0..7 so that a consumer's eight consecutive
  // steps never wrap.
  double xring[(Cfg::RS + ST_RB) * Cfg::CP];
  double yring[HAS_V ? (Cfg::RS + ST_RB) * 32 : 2];
  unsigned ering[ST_NJ * 32];  // E + 0x80000000 - 1023 (mod 2^32) per (producer batch, lane): log_scaled_i
  int pad0[2];
  unsigned long long full[Cfg::NB];
  int empty_gen[Cfg::NB];  // slot s was last released by consumer batch q = s + (gen-1)*NB
  int next_q;
  int pad[3];
};

template <int K, int G, bool HAS_V>
struct StripSmem {
  LogTabEntry logtab[LOGTAB_N + 1];
  StripSub<K, G, HAS_V> sub[G];
  BRing ring[G + 1];  // ring g feeds strip g; ring 0 is filled by the loader, ring G drained by the flusher
};

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

/* two adjacent ints with one volatile 8-byte load (p 8-byte aligned) */
__device__ __forceinline__ int2 ld_vol2(const int *p) {
  const long long v = *reinterpret_cast<const volatile long long *>(p);
  return make_int2((int)v, (int)(v >> 32));
}

// ---- strip geometry shared by the roles ------------------------------------------------------------
struct StripGeom {
  int rs;      // columns rs+1 .. rs+C ; rows n = rs+1 .. N are r = 0 .. R-1
  int R;       // rows of the strip
  int phi;     // batch phase: step t sits at position (t+phi) % ST_B of producer batch (t+phi) / ST_B
  int nbatch;  // producer batches
  int QT;      // consumer row batches
  int delta;   // producer batch p reads the left strip's batch p+delta
};

__device__ __forceinline__ int strip_phi(int strip, int L, int C) {
  return (int)(((long long)strip * (long long)(L - 1 + C)) & (ST_B - 1));
}

__device__ __forceinline__ StripGeom strip_geom(const StripParams &P, int strip) {
  StripGeom g;
  g.rs = strip * P.C;
  g.R = P.N - g.rs;
  g.phi = strip_phi(strip, P.L, P.C);
  g.nbatch = ((g.R - 1 + P.L - 1 + g.phi) >> ST_SH) + 1;
  g.QT = (g.R - 1) / ST_RB + 1;
  g.delta = strip > 0 ? (P.L - 1 + P.C + strip_phi(strip - 1, P.L, P.C) - g.phi) >> ST_SH : 0;
  return g;
}

/* consumer batches that are complete once producer batch p is: q <= this (may be -1) */
__device__ __forceinline__ int strip_qdone(int p, int phi, int L) {
  // after batch p every lane has made rows 0 .. rows-1, rows = ST_B*(p+1) - phi - (L-1)
  const int rows = ST_B * p + ST_B - phi - (L - 1);
  return rows >= 0 ? rows / ST_RB - 1 : -1;
}

// ---- producer ----------------------------------------------------------------------------------------
/*
 * NS recurrence steps.  x[k]: the lane's K columns; the coefficient of column m_k in row n is
 * (n-1) - m_k a, formed per step from nm1 = n-1 and ma[k] = m_k a exactly like that (one rounding,
 * independent of where a strip starts, so every geometry produces the same bits); yin: the left
 * neighbour's value one row up, in this lane's units; scn: the per-batch factor that brings the
 * neighbour's units to this lane's.  Everything is stored unconditionally: rows that do not exist
 * (before the strip's first row, past N) land in ring slots the consumers never read as valid.
 */
template <int K, bool HAS_V, bool DUP, int CP, int RS, int NS>
__device__ __forceinline__ void strip_steps(double (&x)[K], const double (&ma)[K], double &nm1, double &yin,
                                            const double scn, const double *__restrict__ bnd,
                                            const bool take_bnd, const bool write_out, double *__restrict__ xr, double *__restrict__ yr,
                                            double *__restrict__ outp) {
#pragma unroll
  for (int i = 0; i < NS; i++) {
    // the left neighbour's last column BEFORE this step's update: its value one row up from
    // the row this lane makes in the NEXT step
    double s = shfl_up_d(x[K - 1]);
    if (take_bnd) s = bnd[i];
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

/* predicated single instructions: no branch, no reconvergence point in the producer's batch loop */
__device__ __forceinline__ void mbar_arrive_if(unsigned long long *bar, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"((unsigned)pred)
      : "memory");
}
__device__ __forceinline__ void st_shared_if(int *p, int v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p st.volatile.shared.s32 [%0], %1;\n\t}" ::"r"(smem_u32(p)),
      "r"(v), "r"((unsigned)pred)
      : "memory");
}

template <int K, int G, bool HAS_V>
__device__ void strip_producer(const StripParams &P, StripSub<K, G, HAS_V> &sb, BRing *rin, BRing *rout,
                               const bool has_left, const bool has_right, const StripGeom &g, int lane, double a,
                               int jlast) {
  using Cfg = StripCfg<K, G, HAS_V>;
  constexpr int CP = Cfg::CP, RS = Cfg::RS, NB = Cfg::NB;
  const int L = P.L;
  // rin / rout always point at a ring of this CTA; without a neighbour on that side the ring is
  // simply unused by anyone else, which keeps the step code free of branches
  BRing *const rin_ = rin;
  BRing *const rout_ = rout;

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
  const bool lane0 = (lane == 0);
  const bool take_bnd = lane0 && has_left;  // lane 0 of the first strip keeps its shuffled value times zero
  const bool write_out = has_right && (lane == L - 1);
  int qsent = -1;                                    // consumer batches released so far
  int rows = ST_B - g.phi - (L - 1);                 // rows 0..rows-1 are complete after the current batch
  int s0 = 0, gen_need = 0;                          // consumer slot pair of the batch, generation it must have
  int q1slot = 0;                                    // slot of consumer batch qsent+1
  int c_in = -1, c_out = has_right ? ld_vol(&rout->taken) : 0;
#ifdef STB_PROFILE_PRODUCER
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long gt_start = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
#endif
  // words fetched ahead of need (mid-batch), so that the checks at the top of a batch rarely load
  int2 gen_next = make_int2(0, 0);

  for (int p = 0; p < g.nbatch; ++p) {
    ST_TICK(tk0);
    // ---- flow control (uniform across the warp) ----
    // producer batch p overwrites the ring rows of consumer slots s0, s0+1 (s0 = 2p % NB); their
    // previous tenants are consumer batches 2p-NB and 2p+1-NB, released when gen >= 2p/NB
    if (gen_next.x < gen_need || gen_next.y < gen_need) {
      const long long t0 = clock64();
      unsigned spins = 0;
      for (;;) {
        const int2 gen = ld_vol2(&sb.empty_gen[s0]);
        if (gen.x >= gen_need && gen.y >= gen_need) break;
        if ((++spins & 1023u) == 0) {
          const int bad = ld_vol(P.abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
          if (__any_sync(0xffffffffu, bad)) {
            if (lane0) atomicExch(P.abort_flag, 1);
            return;
          }
        }
      }
    }
    ST_TICK(tk1);
    int jb = p + g.delta;
    if (jb > jlast) jb = jlast;
    if (has_left) {
      if (!ctr_wait<false, 0>(&rin->written, jb, P.abort_flag, c_in)) return;
    }
    ST_TICK(tk1b);
    if (has_right) {
      if (!ctr_wait<false, 0>(&rout->taken, p - ST_NBR, P.abort_flag, c_out)) return;
    }
    asm volatile("" ::: "memory");
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
      sb.ering[(p & (ST_NJ - 1)) * 32 + lane] = (unsigned)elow + (0x80000000u - 1023u);
    }
    // scale that brings the left neighbour's values into this lane's units, fixed for the batch
    const int bs = jb & (ST_NBR - 1), os = p & (ST_NBR - 1);
    int sE = __shfl_up_sync(0xffffffffu, elow, 1);
    if (lane0) sE = has_left ? rin_->e[bs] : elow;
    // (lane 0 of a table's first strip has no left neighbour: its shuffled value is multiplied by 0)
    const double scn = (lane0 && !has_left) ? 0.0 : pow2i(sE - elow);
    st_shared_if(&rout_->e[os], elow, write_out);
    const double *bnd = &rin_->x[bs * ST_B];
    double *outp = &rout_->x[os * ST_B];
    double *xr = &sb.xring[s0 * ST_RB * CP + lane * K];
    double *yr = &sb.yring[HAS_V ? s0 * ST_RB * 32 + lane : 0];

    ST_TICK(tk3);
    // ---- sixteen steps, eight per consumer slot; only ring rows 0..7 have duplicates ----
    if (s0 == 0)
      strip_steps<K, HAS_V, true, CP, RS, ST_RB>(x, ma, nm1, yin, scn, bnd, take_bnd, write_out, xr, yr, outp);
    else
      strip_steps<K, HAS_V, false, CP, RS, ST_RB>(x, ma, nm1, yin, scn, bnd, take_bnd, write_out, xr, yr, outp);
    // mid-batch: fetch the words the NEXT batch's flow control will look at
    {
      int s1 = s0 + 2;
      if (s1 >= NB) s1 = 0;
      gen_next = ld_vol2(&sb.empty_gen[s1]);
      if (has_left) {
        const int v = ld_vol(&rin->written);
        c_in = v > c_in ? v : c_in;
      }
      if (has_right) {
        const int v = ld_vol(&rout->taken);
        c_out = v > c_out ? v : c_out;
      }
    }
    strip_steps<K, HAS_V, false, CP, RS, ST_RB>(x, ma, nm1, yin, scn, bnd + ST_RB, take_bnd, write_out,
                                                 xr + ST_RB * CP, yr + ST_RB * 32, outp + ST_RB);
    ST_TICK(tk4);
    // ---- publish: at most two consumer batches complete per producer batch ----
    __syncwarp();
    asm volatile("" ::: "memory");
    {
      int qd = rows >= 0 ? (rows >> 3) - 1 : -1;
      if (qd >= g.QT) qd = g.QT - 1;
      int q2slot = q1slot + 1;
      if (q2slot == NB) q2slot = 0;
      mbar_arrive_if(&sb.full[q1slot], lane0 && qsent + 1 <= qd);
      mbar_arrive_if(&sb.full[q2slot], lane0 && qsent + 2 <= qd);
      const int adv = qd - qsent;  // 0, 1 or 2
      if (adv > 0) {
        qsent = qd;
        q1slot += adv;
        if (q1slot >= NB) q1slot -= NB;
      }
      st_shared_if(&rin_->taken, p + g.delta, lane0 && has_left);
      st_shared_if(&rout_->written, p, write_out);
    }
    rows += ST_B;
    s0 += 2;
    if (s0 >= NB) {
      s0 = 0;
      gen_need++;
    }
    ST_TICK(tk5);
    ST_ACC(0, tk1, tk0);   // wait: ring slots free
    ST_ACC(1, tk1b, tk1);  // wait: boundary in
    ST_ACC(2, tk2, tk1b);  // wait: boundary out
    ST_ACC(3, tk3, tk2);   // renormalise + batch set-up
    ST_ACC(4, tk4, tk3);   // the steps
    ST_ACC(5, tk5, tk4);   // publish
  }
#ifdef STB_PROFILE_PRODUCER
  if (lane0 && P.dbg) {
    long long *d = P.dbg + ((size_t)blockIdx.x * G + (threadIdx.x >> 5)) * 8;
    for (int i = 0; i < 5; i++) d[i] = dbgacc[i];
    d[6] = g.nbatch;
    unsigned long long gt_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
    d[7] = (long long)(gt_end - gt_start);
    d[5] = dbgacc[5];
  }
#endif
  // a partial last row batch never sees its eighth row: release it now that every row exists
  if (lane0)
    for (int q = qsent + 1; q < g.QT; q++) mbar_arrive(&sb.full[q % NB]);
}

// ---- consumer ----------------------------------------------------------------------------------------
template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
__device__ void strip_consumer(const StripParams &P, StripSub<K, G, HAS_V> &sb, const LogTabEntry *logtab,
                               const StripGeom &g, int lane, const StripTable &tb) {
  using Cfg = StripCfg<K, G, HAS_V>;
  constexpr int CP = Cfg::CP, NB = Cfg::NB;
  const int M = P.M;
  int cvalid = M - g.rs;  // columns of this strip that exist
  if (cvalid > P.C) cvalid = P.C;
  OutT *tabS = (OutT *)tb.tabS;
  OutT *tabV = (OutT *)tb.tabV;
  const bool first_strip = (g.rs == 0);
  const unsigned ld32 = (unsigned)P.ld;  // 8 rows x ld elements stay far below 2^32

  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&sb.next_q, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= g.QT) break;
    const int slot = q % NB;
    if (!mbar_wait<ST_CONS_SLEEP>(&sb.full[slot], (unsigned)(q / NB) & 1u, P.abort_flag)) return;
    const int r0 = q * ST_RB;
    const bool fast = (r0 >= cvalid - 1) && (r0 + ST_RB <= g.R);
    // not unrolled: the body is ~250 instructions; K copies of it overflow the instruction cache
    // and the misses stall every warp of the SM, the producers included
#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
      if (32 * kk >= cvalid) break;
      const int col = lane + 32 * kk;
      const int pl = col / K;
      const int kq = col % K;
      const bool lane_ok = col < cvalid;
      const size_t cell0 = (size_t)(g.rs + r0) * P.ld + (size_t)(g.rs + col);  // row n-1 = rs+r, column m-1
      // producer lane pl made rows r0..r0+7 at steps toff..toff+7 (ring rows ub..ub+7, never
      // wrapping thanks to the duplicate rows); its exponent changes at most once on the way
      const int toff = r0 + pl + g.phi;
      const int j0 = toff >> ST_SH, thr = ST_B - (toff & (ST_B - 1));
      const int ub = ((toff >> 3) % NB) * ST_RB + (toff & 7);
      const double *xc = &sb.xring[ub * CP + col];
      const double *yc = &sb.yring[HAS_V ? ub * 32 + pl : 0];
      if (fast) {
        double xv[ST_RB];
#pragma unroll
        for (int i = 0; i < ST_RB; i++) xv[i] = xc[i * CP];
        if (HAS_S) {
          const unsigned Ea = sb.ering[(j0 & (ST_NJ - 1)) * 32 + pl];
          const unsigned Eb = sb.ering[((j0 + 1) & (ST_NJ - 1)) * 32 + pl];
          double v[ST_RB];
#pragma unroll
          for (int i = 0; i < ST_RB; i++) v[i] = log_scaled_i(xv[i], (i >= thr) ? Eb : Ea, logtab);
          if (lane_ok) {
            OutT *pS = tabS + cell0;
#pragma unroll
            for (int i = 0; i < ST_RB; i++) st_out(pS + (size_t)((unsigned)i * ld32), v[i]);
            if (first_strip && col == 0) {
#pragma unroll
              for (int i = 0; i < ST_RB; i++) tb.s1[r0 + i] = v[i];
            }
          }
        }
        if (HAS_V) {
          double den[ST_RB];
#pragma unroll
          for (int i = 0; i < ST_RB; i++) den[i] = (kq == 0) ? yc[i * 32] : xc[i * CP - 1];
          if (lane_ok && !(first_strip && col == 0)) {
            OutT *pV = tabV + cell0;
#pragma unroll
            for (int i = 0; i < ST_RB; i++) st_out(pV + (size_t)((unsigned)i * ld32), div_pos(xv[i], den[i]));
          }
        }
      } else if (lane_ok) {
        // ---- edges: the strip's triangle (column col exists from row r = col on), last rows ----
        for (int i = 0; i < ST_RB; i++) {
          const int r = r0 + i;
          if (r >= g.R || col > r) continue;
          const double xv = xc[i * CP];
          const size_t off = cell0 + (size_t)i * P.ld;
          if (HAS_S) {
            const int j = (toff + i) >> ST_SH;
            const double v = log_scaled_i(xv, sb.ering[(j & (ST_NJ - 1)) * 32 + pl], logtab);
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
    if (lane == 0) st_vol(&sb.empty_gen[slot], q / NB + 1);
  }
}

// ---- loader / flusher: the CTA boundary through an L2-resident ring -----------------------------------
/*
 * Flag-in-data hand-off (the "LL" idea of collective libraries): every double travels as one
 * 16-byte store {lo, seq, hi, seq}; each 8-byte half is written atomically, so a reader that sees
 * seq in both halves has the data -- no fence, no separate counter, one L2 round trip of latency.
 * A batch is ST_B such entries plus one for the exponent; seq = batch index + 1, and the ring is
 * zeroed before every launch.  The only counter left is the reader's progress (back-pressure).
 */
constexpr int ST_GE = ST_B + 1;  // 16-byte entries per batch in the global ring

__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_v4(uint4 *p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ void strip_loader(const StripParams &P, BRing *ring, int delta, int lane, int bidx, int jlast) {
  // boundary bidx is written by the CTA to the left; batches delta .. jlast are needed here
  const uint4 *g = P.gring + (size_t)bidx * (ST_NBG * ST_GE);
  constexpr int NP = 4;  // batches polled per pass (loads in flight per lane)
  int c_t = -1, pub = delta - 1;
  long long t0 = clock64();
  unsigned idle = 0;
  for (int next = delta; next <= jlast;) {
    if (!ctr_wait<false, 64>(&ring->taken, next - ST_NBR, P.abort_flag, c_t)) return;
    int lim = c_t + ST_NBR;  // last batch the shared ring has room for
    if (lim > jlast) lim = jlast;
    uint4 v[NP];
#pragma unroll
    for (int u = 0; u < NP; u++) {
      const int j = next + u;
      v[u] = make_uint4(0, 0, 0, 0);
      if (j <= lim && lane < ST_GE) v[u] = ld_volatile_v4(&g[(size_t)(j & (ST_NBG - 1)) * ST_GE + lane]);
    }
    int got = 0;
#pragma unroll
    for (int u = 0; u < NP; u++) {
      const int j = next + u;
      const unsigned seq = (unsigned)(j + 1);
      const bool ok = (j <= lim) && (lane >= ST_GE || (v[u].y == seq && v[u].w == seq));
      if (got == u && __all_sync(0xffffffffu, ok)) {
        if (lane < ST_B)
          ring->x[(j & (ST_NBR - 1)) * ST_B + lane] = __hiloint2double((int)v[u].z, (int)v[u].x);
        else if (lane == ST_B)
          ring->e[j & (ST_NBR - 1)] = (int)v[u].x;
        got = u + 1;
      }
    }
    if (got) {
      __syncwarp();
      asm volatile("" ::: "memory");
      const int hi = next + got - 1;
      if (lane == 0) {
        st_vol(&ring->written, hi);
        if (hi - pub >= ST_NBG / 4 || hi == jlast) st_vol(P.gtaken + bidx, hi);
      }
      if (hi - pub >= ST_NBG / 4) pub = hi;
      next = hi + 1;
      idle = 0;
      t0 = clock64();
    } else if ((++idle & 1023u) == 0) {
      const int bad = ld_vol(P.abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) atomicExch(P.abort_flag, 1);
        return;
      }
    }
  }
}

__device__ void strip_flusher(const StripParams &P, BRing *ring, int last, int lane, int bidx) {
  uint4 *g = P.gring + (size_t)bidx * (ST_NBG * ST_GE);
  int c_w = -1, c_t = -1;
  for (int next = 0; next <= last;) {
    if (!ctr_wait<false, 32>(&ring->written, next, P.abort_flag, c_w)) return;
    if (!ctr_wait<true, 64>(P.gtaken + bidx, next - ST_NBG, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (hi > c_t + ST_NBG) hi = c_t + ST_NBG;
    for (int j = next; j <= hi; j++) {
      const unsigned seq = (unsigned)(j + 1);
      if (lane < ST_B) {
        const double x = ring->x[(j & (ST_NBR - 1)) * ST_B + lane];
        st_volatile_v4(&g[(size_t)(j & (ST_NBG - 1)) * ST_GE + lane],
                       make_uint4((unsigned)__double2loint(x), seq, (unsigned)__double2hiint(x), seq));
      } else if (lane == ST_B) {
        st_volatile_v4(&g[(size_t)(j & (ST_NBG - 1)) * ST_GE + lane],
                       make_uint4((unsigned)ring->e[j & (ST_NBR - 1)], seq, 0u, seq));
      }
    }
    __syncwarp();
    asm volatile("" ::: "memory");
    if (lane == 0) st_vol(&ring->taken, hi);
    next = hi + 1;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
__global__ void __launch_bounds__(ST_WARPS * 32, 1) fill_strip_kernel(const StripParams P) {
  using SM = StripSmem<K, G, HAS_V>;
  using Cfg = StripCfg<K, G, HAS_V>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &sm = *reinterpret_cast<SM *>(smem_raw);
  const int table = blockIdx.x / P.ctas, cta = blockIdx.x % P.ctas;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip0 = cta * G;
  int nloc = P.P - strip0;  // strips of this CTA
  if (nloc > G) nloc = G;
  const StripTable tb = P.tables[table];
  const bool cta_left = cta > 0, cta_right = strip0 + nloc < P.P;

  for (int i = threadIdx.x; i < LOGTAB_N; i += blockDim.x) sm.logtab[i] = P.logtab[i];
  if (threadIdx.x < G) {
    auto &sb = sm.sub[threadIdx.x];
    for (int s = 0; s < Cfg::NB; s++) {
      mbar_init(&sb.full[s], 1);
      sb.empty_gen[s] = 0;
    }
    sb.next_q = 0;
  }
  if (threadIdx.x <= G) {
    // ring g feeds strip strip0+g, whose first batch reads batch delta: earlier batches count as taken
    const int rg = threadIdx.x, st = strip0 + rg;
    sm.ring[rg].written = -1;
    sm.ring[rg].taken = (st > 0 && st < P.P) ? strip_geom(P, st).delta - 1 : -1;
  }
  __syncthreads();

  if (warp < G) {
    // producers: warp g on SM sub-partition g
    const int gi = warp;
    if (gi < nloc) {
      const int strip = strip0 + gi;
      const StripGeom g = strip_geom(P, strip);
      const int jlast = strip > 0 ? strip_geom(P, strip - 1).nbatch - 1 : 0;
      strip_producer<K, G, HAS_V>(P, sm.sub[gi], &sm.ring[gi], &sm.ring[gi + 1], strip > 0, strip + 1 < P.P, g, lane,
                                  tb.a, jlast);
    }
  } else if (warp == ST_LOADER) {
    if (cta_left) {
      const StripGeom g = strip_geom(P, strip0);
      const int jlast = strip_geom(P, strip0 - 1).nbatch - 1;
      strip_loader(P, &sm.ring[0], g.delta, lane, table * P.ctas + cta - 1, jlast);
    }
  } else if (warp == ST_FLUSHER) {
    if (cta_right) {
      const StripGeom g = strip_geom(P, strip0 + nloc - 1);
      strip_flusher(P, &sm.ring[nloc], g.nbatch - 1, lane, table * P.ctas + cta);
    }
  } else {
    // consumers: dealt round-robin to the CTA's strips.  A waiter may be at most one phase ahead
    // of its mbarrier, so a strip gets fewer claimants than it has batch slots.  Consumers fill
    // the sub-partitions without a producer; a producer's own sub-partition takes only `spread`
    // of them, so that the recurrence keeps most of its issue slots and FP64 pipe.
    // consumer index c: warps on the free sub-partitions first, then rows 1..spread of the
    // producers' sub-partitions (minus the two helper warps)
    const int sp = warp & 3, row = warp >> 2;
    int c;
    if (sp >= G) {
      c = row * (4 - G) + (sp - G);
    } else {
      if (G < 4 && row > P.spread) return;
      c = 4 * (4 - G) + (row - 1) * G + sp - (warp > ST_LOADER) - (warp > ST_FLUSHER);
    }
    const int gi = c % G, ci = c / G;
    if (gi < nloc && ci < Cfg::NB - 1 && ci < P.ncons) {
      const StripGeom g = strip_geom(P, strip0 + gi);
      strip_consumer<K, G, HAS_S, HAS_V, OutT>(P, sm.sub[gi], sm.logtab, g, lane, tb);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct StripState {
  uint4 *gring;
  int *gctr;  // [cap] taken, then the abort flag at [cap]
  int cap;    // CTA boundaries the buffers can serve
  LogTabEntry *logtab;
  StripTable *tables;  // device array
  int tables_cap;
  long long *dbg;  // STB_PROFILE_PRODUCER builds only
};

inline void strip_state_free(StripState *st) {
  cudaFree(st->gring);
  cudaFree(st->gctr);
  cudaFree(st->logtab);
  cudaFree(st->tables);
  cudaFree(st->dbg);
  memset(st, 0, sizeof *st);
}

inline size_t strip_state_bytes(const StripState *st) {
  return (size_t)st->cap * ST_NBG * ST_GE * sizeof(uint4) + (st->cap ? ((size_t)st->cap + 1) * sizeof(int) : 0) + (st->logtab ? LOGTAB_N * sizeof(LogTabEntry) : 0) +
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
    cudaFree(st->gring);
    cudaFree(st->gctr);
    st->gring = NULL;
    st->gctr = NULL;
    st->cap = 0;
    if ((e = cudaMalloc(&st->gring, (size_t)nbound * ST_NBG * ST_GE * sizeof(uint4))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->gctr, ((size_t)nbound + 1) * sizeof(int))) != cudaSuccess) return e;
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

template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
inline cudaError_t launch_strip(const StripParams &P, int nctas, cudaStream_t stream) {
  const size_t smem = sizeof(StripSmem<K, G, HAS_V>);
  auto kern = fill_strip_kernel<K, G, HAS_S, HAS_V, OutT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void *args[] = {(void *)&P};
  return cudaLaunchCooperativeKernel((void *)kern, dim3(nctas), dim3(ST_WARPS * 32), args, smem, stream);
}

template <int K, int G>
inline cudaError_t dispatch_strip(const StripParams &P, int nctas, bool hasS, bool hasV, bool is_float,
                                  cudaStream_t stream) {
  if (is_float) {
    if (hasS && hasV) return launch_strip<K, G, true, true, float>(P, nctas, stream);
    if (hasS) return launch_strip<K, G, true, false, float>(P, nctas, stream);
    return launch_strip<K, G, false, true, float>(P, nctas, stream);
  }
  if (hasS && hasV) return launch_strip<K, G, true, true, double>(P, nctas, stream);
  if (hasS) return launch_strip<K, G, true, false, double>(P, nctas, stream);
  return launch_strip<K, G, false, true, double>(P, nctas, stream);
}

struct StripPlan {
  int K, G, L, C, P, ctas;  // columns per lane, strips per CTA, lanes, columns per strip, strips and CTAs per table
};

/*
 * Geometry for tables of M columns when `slots` CTAs are available per table.  The compiled
 * shapes are (K columns per lane, G strips per CTA) = (1,4), (5,1), (3,2), (7,1): up to 128, 160,
 * 192 and 224 columns per CTA, tried in that order (measured best first); the lane count is then trimmed so that the columns are
 * spread evenly and strip edges fall on 32-byte sectors of the table.
 */
inline bool strip_plan(unsigned M, int slots, size_t elem_size, StripPlan *pl) {
  static const int shapes[4][2] = {{1, 4}, {5, 1}, {3, 2}, {7, 1}};
  int force_k = 0, force_l = 0;
  if (const char *s = getenv("STB_STRIP_K")) force_k = atoi(s);
  if (const char *s = getenv("STB_STRIP_L")) force_l = atoi(s);
  if (slots < 1) slots = 1;
  const unsigned per_cta = (M + (unsigned)slots - 1) / (unsigned)slots;  // columns per CTA with every slot in use
  for (int i = 0; i < 4; i++) {
    const int k = shapes[i][0], gg = shapes[i][1];
    if (force_k ? k != force_k : (unsigned)(32 * k * gg) < per_cta) continue;
    unsigned want = (per_cta + (unsigned)gg - 1) / (unsigned)gg;  // columns per strip
    if (want < 32u) want = M < 32u ? M : 32u;  // never narrower than a warp unless the table is
    int L = (int)((want + (unsigned)k - 1) / (unsigned)k);
    if (L > 32) L = 32;
    if (L < 1) L = 1;
    while (L < 32 && ((size_t)L * k * elem_size) % 32 != 0) L++;
    if (force_l) L = force_l;
    pl->K = k;
    pl->G = gg;
    pl->L = L;
    pl->C = L * k;
    pl->P = (int)((M + (unsigned)pl->C - 1) / (unsigned)pl->C);
    pl->ctas = (pl->P + gg - 1) / gg;
    if (pl->ctas > slots) continue;
    return true;
  }
  return false;
}

/* tables one launch fills side by side: each gets the CTAs its widest shape needs */
inline int strip_tables_per_launch(unsigned M, int num_sms) {
  int pmin = (int)((M + 32u * 7u - 1) / (32u * 7u));
  int per_launch = num_sms / (pmin > 0 ? pmin : 1);
  return per_launch < 1 ? 1 : per_launch;
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
  // the per-lane exponent is carried in 32 bits: N log2 N must stay below 2^31
  if (A.N > 50000000u || A.M > A.N || A.M < 1 || (unsigned long long)A.ld * 8 >= 0xffffffffull) {
    snprintf(err, errlen, "strip_fill: unsupported extent N=%u M=%u", A.N, A.M);
    return -1;
  }
  // tables per launch: as many as fit when each gets the CTAs its widest shape needs
  StripPlan pl;
  int per_launch = 1;
  if (A.ntables > 1) {
    per_launch = strip_tables_per_launch(A.M, A.num_sms);
    if (per_launch > A.ntables) per_launch = A.ntables;
  }
  if (!strip_plan(A.M, A.num_sms / per_launch, A.is_float ? 4 : 8, &pl)) {
    snprintf(err, errlen, "strip_fill: M=%u needs more than one pass over the columns (not supported)", A.M);
    return -1;
  }
  cudaError_t e = strip_state_prepare(st, per_launch * pl.ctas, A.ntables);
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
  P.ctas = pl.ctas;
  P.gring = st->gring;
  P.gtaken = st->gctr;
  P.abort_flag = st->gctr + st->cap;
  P.logtab = st->logtab;
  P.ncons = 64;
  if (const char *s = getenv("STB_STRIP_CONS")) P.ncons = atoi(s);
  P.spread = 1;
  if (const char *s = getenv("STB_STRIP_SPREAD")) P.spread = atoi(s);
  P.dbg = NULL;
#ifdef STB_PROFILE_PRODUCER
  if (!st->dbg) cudaMalloc(&st->dbg, 1024 * 8 * sizeof(long long));
  cudaMemsetAsync(st->dbg, 0, 1024 * 8 * sizeof(long long), stream);
  P.dbg = st->dbg;
#endif
  for (int t0 = 0; t0 < A.ntables && e == cudaSuccess; t0 += per_launch) {
    const int nt = (A.ntables - t0 < per_launch) ? A.ntables - t0 : per_launch;
    P.tables = st->tables + t0;
    // reader progress starts at -1 ("nothing taken"), the abort flag at 0, the ring's flags at 0
    e = cudaMemsetAsync(st->gctr, 0xFF, (size_t)st->cap * sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.abort_flag, 0, sizeof(int), stream);
    if (e == cudaSuccess)
      e = cudaMemsetAsync(st->gring, 0, (size_t)nt * pl.ctas * ST_NBG * ST_GE * sizeof(uint4), stream);
    if (e != cudaSuccess) break;
    const int nctas = nt * pl.ctas;
    const bool hs = A.has_S != 0, hv = A.has_V != 0, fl = A.is_float != 0;
    if (pl.K == 1)
      e = dispatch_strip<1, 4>(P, nctas, hs, hv, fl, stream);
    else if (pl.K == 3)
      e = dispatch_strip<3, 2>(P, nctas, hs, hv, fl, stream);
    else if (pl.K == 5)
      e = dispatch_strip<5, 1>(P, nctas, hs, hv, fl, stream);
    else
      e = dispatch_strip<7, 1>(P, nctas, hs, hv, fl, stream);
    if (e == cudaSuccess) {
      // the abort flag is checked per launch: a later memset must not hide it
      int flag = 0;
      if (t0 + per_launch >= A.ntables) cudaEventRecord(ev_end, stream);
      e = cudaMemcpyAsync(&flag, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
#ifdef STB_PROFILE_PRODUCER
      if (e == cudaSuccess && getenv("STB_PROFILE_PRINT")) {
        static long long h[1024 * 8];
        cudaMemcpy(h, st->dbg, sizeof h, cudaMemcpyDeviceToHost);
        const int show[6] = {0, 1, 2, nctas / 2, nctas - 2, nctas - 1};
        for (int si = 0; si < 6; si++) {
          const int c = show[si];
          if (c < 0 || c >= nctas || (si && c <= show[si - 1])) continue;
          for (int gi = 0; gi < pl.G; gi++) {
            const long long *d = h + ((size_t)c * pl.G + gi) * 8;
            const double nb = (double)d[6];
            if (nb <= 0 || (size_t)c * pl.G + gi >= 1024) continue;
            fprintf(stderr,
                    "cta %3d producer %d batches %6.0f cycles/batch: slot-wait %.0f in-wait %.0f out-wait %.0f setup %.0f "
                    "steps %.0f publish %.0f | busy %.1f us\n",
                    c, gi, nb, d[0] / nb, d[1] / nb, d[2] / nb, d[3] / nb, d[4] / nb, d[5] / nb, d[7] / 1e3);
          }
        }
      }
#endif
      if (e == cudaSuccess && flag) {
        snprintf(err, errlen, "strip_fill: pipeline watchdog fired (K=%d G=%d L=%d P=%d tables/launch=%d)", pl.K, pl.G,
                 pl.L, pl.P, per_launch);
        return -2;
      }
    }
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "strip_fill (K=%d G=%d L=%d P=%d tables/launch=%d): %s", pl.K, pl.G, pl.L, pl.P, per_launch,
             cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace stb
