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
 *     is one progress word (the last batch whose steps are all in the ring); consumer -> producer
 *     a count of finished units per ring slot.
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
#ifndef ST_WARPS_OVERRIDE
#define ST_WARPS_OVERRIDE 16
#endif
constexpr int ST_WARPS = ST_WARPS_OVERRIDE;  // warps per CTA of a kernel that makes S and V
#ifndef ST_WARPS1_OVERRIDE
#define ST_WARPS1_OVERRIDE 20
#endif
constexpr int ST_WARPS1 = ST_WARPS1_OVERRIDE;  // ... of one that makes S or V alone: its consumers need fewer registers, twenty warps fit (15 consumers, 5 per free sub-partition; measured on config 2: S 9.36 -> 9.06 ms on the same box; with S and V the registers of 20 warps spill: 13.7 -> 15.1 ms)
constexpr int ST_MAXW = ST_WARPS1 > ST_WARPS ? ST_WARPS1 : ST_WARPS;
__host__ __device__ constexpr int strip_warps(bool has_s, bool has_v) { return (has_s && has_v) ? ST_WARPS : ST_WARPS1; }
#ifndef ST_LOADER_OVERRIDE
#define ST_LOADER_OVERRIDE 4
#define ST_FLUSHER_OVERRIDE 8
#endif
constexpr int ST_LOADER = ST_LOADER_OVERRIDE, ST_FLUSHER = ST_FLUSHER_OVERRIDE;  // helper warps: they mostly sleep, so they share sub-partition 0 with producer 0 and leave the other three to twelve consumers (measured: 10.4 -> 9.4 ms on config 2 against helpers on sub-partition 3); producers are warps 0..G-1, the rest consume
#ifndef ST_HELPER_SLEEP
#define ST_HELPER_SLEEP 64  // ns the loader / flusher sleep between polls when there is nothing to move
#endif
#ifndef ST_CONS_SLEEP
#define ST_CONS_SLEEP 40  // ns a consumer sleeps between polls of its batch barrier
#endif

struct StripTable {  // one table of a launch
  void *tabS, *tabV;
  double *s1;
  double a;
};

struct StripParams {
  const StripTable *tables;  // device array, one per table of the launch
  int ntables;               // tables of the launch; CTA group q (ctas CTAs) fills tables q, q + ngroups, q + 2 ngroups, ...
  int ngroups;               // CTA groups of the launch (= tables filled side by side)
  unsigned long long ld;     // elements per table row
  int N, M;
  int C;     // columns per strip (= L*K)
  int L;     // producer lanes in use
  int P;     // strips per table
  int ctas;  // CTAs per table (= ceil(P / G))
  uint4 *gring;  // [boundaries][ST_NBG*(ST_B+1)] flag-in-data entries, zeroed before each launch
  int *gtaken;   // [boundaries] reader progress (back-pressure only)
  // Tables wider than one launch can hold are filled in several passes over the columns: a pass
  // covers strips strip_base .. strip_base + ctas*G - 1; the last strip's boundary column is kept
  // for ALL rows in a linear buffer (carry_out: same entries as the ring, batch j at [j*ST_GE],
  // no wrap, no back-pressure) that the first CTA of the next pass reads as carry_in.
  int strip_base;
  const uint4 *carry_in;
  uint4 *carry_out;
  int *abort_flag;
  const double *logtab;  // 8-byte table, LOGTAB_N entries (fill_common.cuh)
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
  static constexpr int FIXED = 256 + LOGTAB_N * LOGTAB_REP8 * 8 + (G + 1) * (ST_NBR * ST_B * 8 + ST_NBR * 4 + 16) + G * (ST_NJ * 32 * 4 + 512);
  static constexpr int NB_FIT = ((227 * 1024 - FIXED) / G / ROW_BYTES / ST_RB - 1) & ~1;
  static constexpr int NB = NB_FIT > 16 ? 16 : NB_FIT;
  static constexpr int RS = NB * ST_RB;  // ring rows; rows RS..RS+7 duplicate rows 0..7
  static_assert(NB >= 6, "x ring too small for this geometry");
  // one strip per CTA and a deep ring: the flusher reads the boundary column from the x ring (strip_producer)
  static constexpr bool FLUSH_FROM_XRING = (G == 1) && (NB >= 16);
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
  unsigned ering[ST_NJ * 32];  // E + 0x80000000 - 1023 (mod 2^32) per (producer batch, lane): log_scaled_r
  int pad0[2];
  int progress;  // last producer batch whose sixteen steps are all in the ring (-1: none yet)
  int padp[3];
  int empty_gen[Cfg::NB];  // units (8 rows x 32 columns) of slot s finished so far: U per tenant batch
  int pad_q;
  int pad[3];
};

template <int K, int G, bool HAS_V>
struct StripSmem {
  double logtab[LOGTAB_N * LOGTAB_REP8];
  StripSub<K, G, HAS_V> sub[G];
  BRing ring[G + 1];  // ring g feeds strip g; ring 0 is filled by the loader, ring G drained by the flusher
};

// ---- waits ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

/*
 * Memory ordering inside a CTA (PTX memory model, cta scope).  Every flag that hands data from one
 * warp to another -- sb.progress, empty_gen[], BRing::written / taken -- is PUBLISHED with release
 * semantics (release_cta(): the warp has synchronised, then one fence.acq_rel.cta in front of the flag
 * store / reduction) and READ with ld.acquire.cta (an ordinary LDS in SASS: shared-memory accesses of a
 * warp execute in order, so acquire costs nothing; release is one MEMBAR.ALL.CTA per 16-row batch in
 * the producer and per 8 x 32 unit in a consumer).  The data itself moves with volatile accesses.
 * Measured cost of the fences on config 2 (B200): S 8.80 -> 9.05 ms, S+V 13.25 -> 13.30, V 9.6 -> 11.0
 * (a consumer's fence also waits for the acknowledgement of the unit's global stores).  Tried instead:
 * the slot release as an mbarrier arrival (release semantics without a MEMBAR): S 9.26, V 10.6 with the
 * producer's fence, 8.78 / 10.1 without -- no better; releasing the slot before the unit's stores: the
 * fence then costs 0.04 ms but the kernel 0.4 ms more.  -DSTB_FENCES=0 builds the round-1 protocol
 * (volatile accesses and compiler barriers only; development only).
 */
#ifndef STB_FENCES
#define STB_FENCES 1
#endif
__device__ __forceinline__ void release_cta() {
#if STB_FENCES
  asm volatile("fence.acq_rel.cta;" ::: "memory");
#else
  asm volatile("" ::: "memory");
#endif
}
__device__ __forceinline__ int ld_acquire_shared(const int *p) {
  int v;
#if STB_FENCES
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
#else
  v = *(const volatile int *)p;
#endif
  return v;
}

/*
 * Warp-collective wait until *ctr >= need (shared-memory counter: acquire at cta scope; global
 * counter: relaxed polls, then an acq_rel fence at gpu scope).
 */
template <bool GLOBAL, int SLEEP>
__device__ __forceinline__ bool ctr_wait(const int *ctr, int need, int *abort_flag, int &cached) {
  if (cached >= need) return true;
  int v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_acquire_shared(ctr);
  if (v < need) {
    const long long t0 = clock64();
    unsigned spins = 0;
    do {
      if (SLEEP) __nanosleep(SLEEP);
      v = GLOBAL ? ld_relaxed_gpu(ctr) : ld_acquire_shared(ctr);
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

// ---- producer ----------------------------------------------------------------------------------------
/*
 * Shared-memory accesses of the recurrence by 32-bit shared address with immediate offsets
 * (volatile: kept in program order, which is what makes a value stored by one lane in step i
 * readable by its neighbour in step i+1 without a warp barrier -- one warp's shared-memory
 * instructions execute in order).
 */
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
  asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v));
}
__device__ __forceinline__ void sts_f64_if(unsigned a, double v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p st.volatile.shared.f64 [%0], %1;\n\t}" ::"r"(a),
      "d"(v), "r"((unsigned)pred));
}
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
/* flag words (acquire at cta scope: what the flag guards is ordered behind the load) */
__device__ __forceinline__ int lds_s32(unsigned a) {
  int v;
#if STB_FENCES
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
#else
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
#endif
  return v;
}
__device__ __forceinline__ int2 lds_v2s32(unsigned a) {
  int2 v;
#if STB_FENCES
  asm volatile("ld.acquire.cta.shared.s32 %0, [%2];\n\tld.acquire.cta.shared.s32 %1, [%2+4];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
#else
  asm volatile("ld.volatile.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
#endif
  return v;
}
__device__ __forceinline__ void sts_s32_if(unsigned a, int v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p st.volatile.shared.s32 [%0], %1;\n\t}" ::"r"(a),
      "r"(v), "r"((unsigned)pred)
      : "memory");
}
/*
 * Eight recurrence steps.  x[k]: the lane's K columns; the coefficient of column m_k in row n is
 * (n-1) - m_k a, formed per step from nm1 = n-1 and ma[k] = m_k a exactly like that (one rounding,
 * independent of where a strip starts, so every geometry produces the same bits); yin: the left
 * neighbour's value one row up, in this lane's units; scn: the per-batch factor that brings the
 * neighbour's units to this lane's.
 *
 * The left neighbour's last column after step i-1 is what this lane needs in step i+1 (the lanes
 * walk a diagonal).  It is read back from where the neighbour stored it for the consumers -- the
 * x ring; lane 0 reads the boundary ring instead -- by ONE load issued at the top of step i and
 * first used at the bottom of step i (yin for step i+1), so that its latency hides behind the
 * step's own arithmetic: the row-to-row critical path is the DFMA, nothing else.  nb_addr is the
 * address of the next such load and advances by nb_stride (a ring row / one boundary entry).
 * The first step of a batch reads the LAST row of the previous batch (first_addr), whose values
 * are still in the neighbour's old units: scn0 = (old neighbour units -> old own units) x (own
 * rescale), a product of two powers of two, converts them exactly.  Everything is stored unconditionally: rows that do not exist (before the
 * strip's first row, past N) land in ring slots the consumers never read as valid.
 */
#ifndef STB_NB_EARLY
#define STB_NB_EARLY 0  // 1: the neighbour load of step i+1 issued in step i, right behind the store that makes it readable (two steps of latency instead of one): measured 1.2 % SLOWER on config 2 (9.35 against 9.24 ms) -- the load is not what holds the producer up
#endif
__device__ __forceinline__ void lds_f64_if(double &v, unsigned a, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p ld.volatile.shared.f64 %0, [%1];\n\t}"
      : "+d"(v)
      : "r"(a), "r"((unsigned)pred));
}
template <int K, bool HAS_V, bool DUP, int CP, int RS, bool FIRST, bool OUT>
__device__ __forceinline__ void strip_steps(double (&x)[K], const double (&ma)[K], double &nm1, double &yin, double &nbp,
                                            const bool lane0, const unsigned first_addr, const double scn0, const double scn,
                                            unsigned &nb_addr,
                                            const unsigned nb_stride, const bool write_out, const unsigned xr,
                                            const unsigned yr, const unsigned outp) {
#pragma unroll
  for (int i = 0; i < ST_RB; i++) {
#if STB_NB_EARLY
    // (the load of this step's neighbour value was issued in the PREVIOUS step, right behind the store that made it
    // readable -- see below; only lane 0, which reads the boundary ring, fetches the first entry of a batch here,
    // once the batch's flow control has made sure that it is there)
    if (FIRST && i == 0) lds_f64_if(nbp, first_addr, lane0);
#else
    double nb;
    if (FIRST && i == 0)
      nb = lds_f64(first_addr);
    else {
      nb = lds_f64(nb_addr);
      nb_addr += nb_stride;
    }
#endif
    // Stores (kept in program order) are placed where their operands are long since ready: columns
    // 0..K-2 of the PREVIOUS row go out now, before this step's arithmetic; only column K-1 -- the
    // one the right-hand neighbour reads back in its next step -- is stored in its own step, and
    // it is the first value the step computes.  (Constant offsets fold into the instructions.)
    if (i > 0) {
#pragma unroll
#ifdef STB_EXP_NOSTS  /* timing experiment (wrong results): the producer without 4 of its 5 stores per row */
      for (int k = 0; k < 0; k++) {
#else
      for (int k = 0; k < K - 1; k++) {
#endif
        sts_f64(xr + ((i - 1) * CP + k) * 8, x[k]);
        if (DUP) sts_f64(xr + ((RS + i - 1) * CP + k) * 8, x[k]);
      }
      // the neighbour's value of the previous row (for V), deferred the same way: stored in its own
      // step it waited for the load and the multiply that made it
      if (HAS_V) {
        sts_f64(yr + (i - 1) * 32 * 8, yin);
        if (DUP) sts_f64(yr + (RS + i - 1) * 32 * 8, yin);
      }
    }
#pragma unroll
    for (int k = K - 1; k >= 1; k--) x[k] = fma(nm1 - ma[k], x[k], x[k - 1]);
    x[0] = fma(nm1 - ma[0], x[0], yin);
    nm1 += 1.0;
    sts_f64(xr + (i * CP + K - 1) * 8, x[K - 1]);
    if (DUP) sts_f64(xr + ((RS + i) * CP + K - 1) * 8, x[K - 1]);
    if (OUT) sts_f64_if(outp + i * 8, x[K - 1], write_out);  // (G == 1: the flusher takes the column from the x ring)
#if STB_NB_EARLY
    // The left neighbour has just stored the value this lane needs in step i+2 (one instruction of the warp ago):
    // the load goes out NOW and is first used at the bottom of the NEXT step -- almost two steps of latency it may
    // take, where a load at the top of the next step had one (under the consumers' shared-memory traffic the
    // producer stalled ~10 cycles a row on it).  What this step hands on (yin, for step i+1) was loaded a step ago.
    const double nbn = lds_f64(nb_addr);
    nb_addr += nb_stride;
    yin = nbp * ((FIRST && i == 0) ? scn0 : scn);
    nbp = nbn;
#else
    yin = nb * ((FIRST && i == 0) ? scn0 : scn);
#endif
  }
  // the last row's remaining columns
  if (HAS_V) {
    sts_f64(yr + (ST_RB - 1) * 32 * 8, yin);
    if (DUP) sts_f64(yr + (RS + ST_RB - 1) * 32 * 8, yin);
  }
#pragma unroll
  for (int k = 0; k < K - 1; k++) {
    sts_f64(xr + ((ST_RB - 1) * CP + k) * 8, x[k]);
    if (DUP) sts_f64(xr + ((RS + ST_RB - 1) * CP + k) * 8, x[k]);
  }
}

/*
 * Slow path of the producer's flow control (rare: the fast path is three register compares on
 * words fetched half a batch earlier).  Waits until the two x-ring slots of the batch are free,
 * the left boundary batch has arrived and the right boundary ring has room; returns false when
 * the fill was aborted.  Not inlined: keeps the batch loop short.
 */
__device__ __noinline__ bool producer_wait(unsigned a_gen, int gen_need, unsigned a_written, int need_in,
                                           unsigned a_taken, int need_out, int *abort_flag) {
  const long long t0 = clock64();
  unsigned spins = 0;
  for (;;) {
    const int2 gen = lds_v2s32(a_gen);
    const int w = lds_s32(a_written), t = lds_s32(a_taken);
    if (gen.x >= gen_need && gen.y >= gen_need && w >= need_in && t >= need_out) return true;
    if ((++spins & 1023u) == 0) {
      const int bad = ld_vol(abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
      if (__any_sync(0xffffffffu, bad)) {
        if ((threadIdx.x & 31) == 0) atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
}

template <int K, int G, bool HAS_V>
__device__ void strip_producer(const StripParams &P, StripSub<K, G, HAS_V> &sb, BRing *rin, BRing *rout,
                               const bool has_left, const bool has_right, const StripGeom &g, int lane, double a,
                               int jlast) {
  using Cfg = StripCfg<K, G, HAS_V>;
  constexpr int CP = Cfg::CP, RS = Cfg::RS, NB = Cfg::NB;
  constexpr int NO_NEED = -0x40000000;  // "nothing to wait for"
  const int L = P.L;
  // rin / rout always point at a ring of this CTA; without a neighbour on that side the ring is
  // simply unused by anyone else (and zero-filled), which keeps the step code free of branches

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
  double nbp = 0.0;  // the neighbour value in flight (STB_NB_EARLY): batch 0 reads "the previous batch's last row", zeros
  const bool lane0 = (lane == 0);
  const bool take_bnd = lane0 && has_left;  // lane 0 of the first strip reads zeros times zero
  const bool write_out = has_right && (lane == L - 1);
  // The strip's last column goes to the right-hand neighbour.  Inside a CTA (G > 1) the producer stores it a
  // second time, into the neighbour's boundary ring.  With one strip per CTA (the shapes the planner picks)
  // nobody in the CTA reads that ring: the flusher takes the column straight from the x ring -- it is one more
  // reader of the ring slots, its progress word (rout->taken) is what the producer's flow control looks at --
  // and the recurrence's step is one predicated store shorter.
  // (only when the x ring is deep enough: the flusher then has NB/2 = 8 batches to follow the producer in, as many
  // as it needs not to hold it back; K = 7 strips have 12 slots and keep the boundary ring -- measured on the
  // config-3 shape: 5 % slower without it)
  constexpr bool OUT = !StripCfg<K, G, HAS_V>::FLUSH_FROM_XRING;
  // shared addresses, formed once
  const unsigned a_xr = smem_u32(&sb.xring[lane * K]);
  const unsigned a_yr = smem_u32(&sb.yring[HAS_V ? lane : 0]);
  const unsigned a_er = smem_u32(&sb.ering[lane]);
  const unsigned a_gen = smem_u32(&sb.empty_gen[0]);
  const unsigned a_prog = smem_u32(&sb.progress);
  const unsigned a_in_x = smem_u32(&rin->x[0]), a_in_e = smem_u32(&rin->e[0]);
  const unsigned a_in_written = smem_u32(&rin->written), a_in_taken = smem_u32(&rin->taken);
  const unsigned a_out_x = smem_u32(&rout->x[0]), a_out_e = smem_u32(&rout->e[0]);
  const unsigned a_out_written = smem_u32(&rout->written), a_out_taken = smem_u32(&rout->taken);
  // where lane l finds its left neighbour's last column: one ring row up, one column to the left
  const unsigned nb_stride = lane0 ? 8u : (unsigned)(CP * 8);

  // exponent state: the scale of batch p+1 is fixed in the MIDDLE of batch p (any power of two is
  // exact; sixteen + eight rows of growth stay far inside the double range), which takes the
  // exponent extraction, the neighbour's exponent and the scale factors off the critical path
  long long E = 0;
  int elow = 0;
  int e_n = 0, elow_n = 0, sE_n = 0;  // predicted for the next batch: exponent step, E, neighbour's E
  double sc_n = 1.0, scn_prev = 0.0;
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a_er), "r"(0x80000000u - 1023u));
  // batch 0 reads "the previous batch's last row": zeros
  sts_f64(a_xr + ((RS - 1) * CP + K - 1) * 8, 0.0);

  int cvalid_p = P.M - g.rs;
  if (cvalid_p > P.C) cvalid_p = P.C;
  const int U = (cvalid_p + 31) >> 5;  // units per consumer batch (see strip_consumer)
  int s0 = 0, gen_need = 0;           // consumer slot pair of the batch, finished units it must show
  int c_in = -1, c_out = has_right ? lds_s32(a_out_taken) : 0;
#ifdef STB_PROFILE_PRODUCER
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long gt_start = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
  const long long ck_start = clock64();
#endif
  // words fetched ahead of need (mid-batch), so that the checks at the top of a batch rarely load
  int2 gen_next = make_int2(0, 0);

  for (int p = 0; p < g.nbatch; ++p) {
    ST_TICK(tk0);
    // ---- flow control (uniform across the warp) ----
    // producer batch p overwrites the ring rows of consumer slots s0, s0+1 (s0 = 2p % NB); their
    // previous tenants are consumer batches 2p-NB and 2p+1-NB, released when U * (2p/NB) units
    // of each slot are finished
    int jb = p + g.delta;
    if (jb > jlast) jb = jlast;
    // batch p overwrites the x-ring rows of batch p - NB/2: the flusher must be done with that one (G == 1);
    // G > 1: the boundary ring holds ST_NBR batches
    const int need_in = has_left ? jb : NO_NEED, need_out = has_right ? p - (OUT ? ST_NBR : NB / 2) : NO_NEED;
#if defined(STB_EXP_PRODONLY) || defined(STB_EXP_NOGENWAIT)  /* timing experiments (wrong results): the producer never waits for ring slots (PRODONLY: there are no consumers) */
    gen_next = make_int2(0x3fffffff, 0x3fffffff);
#endif
    if (gen_next.x < gen_need || gen_next.y < gen_need || c_in < need_in || c_out < need_out) {
#ifdef STB_PROFILE_PRODUCER
      {  // which condition holds the producer up (cycles attributed to the first one found wanting)
        const long long tw0 = clock64();
        const int2 gq = lds_v2s32(a_gen + s0 * 4);
        const int wq = has_left ? lds_s32(a_in_written) : 0, tq = has_right ? lds_s32(a_out_taken) : 0;
        const int why = (gq.x < gen_need || gq.y < gen_need) ? 1 : (wq < need_in ? 7 : (tq < need_out ? 6 : 7));
        if (!producer_wait(a_gen + s0 * 4, gen_need, a_in_written, need_in, a_out_taken, need_out, P.abort_flag)) return;
        if (why != 1) dbgacc[why] += clock64() - tw0;
      }
#else
      if (!producer_wait(a_gen + s0 * 4, gen_need, a_in_written, need_in, a_out_taken, need_out, P.abort_flag)) return;
#endif
      c_in = max(c_in, need_in);
      c_out = max(c_out, need_out);
    }
    ST_TICK(tk2);
#ifdef STB_PROFILE_TRACE
    if (p < 24 && lane0 && P.dbg && blockIdx.x < 4) {  // start of each of the strip's first batches, ns since the kernel's start
      unsigned long long gt_now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_now));
      P.dbg[900 * 8 + blockIdx.x * 24 + p] = (long long)(gt_now - gt_start);
    }
    if (p >= g.nbatch - 24 && lane0 && P.dbg && blockIdx.x < 4) {  // ... and of its last ones
      unsigned long long gt_now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_now));
      P.dbg[900 * 8 + 96 + blockIdx.x * 24 + (p - (g.nbatch - 24))] = (long long)(gt_now - gt_start);
    }
#endif
#ifdef STB_PROFILE_PRODUCER
    if (p == 0) {  // when the strip's first batch starts (its input has arrived), ns since the kernel's start
      unsigned long long gt_now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_now));
      dbgacc[1] = (long long)(gt_now - gt_start);
    }
#endif
    // ---- the batch's scale (fixed half a batch ago) ----
    const int bs = jb & (ST_NBR - 1), os = p & (ST_NBR - 1);
#pragma unroll
    for (int k = 0; k < K; k++) x[k] *= sc_n;
    yin *= sc_n;
    E += e_n;
    elow = elow_n;
    int sE = sE_n;
    if (lane0) sE = has_left ? lds_s32(a_in_e + bs * 4) : elow;
    // (lane 0 of a table's first strip has no left neighbour: what it reads is multiplied by 0)
    const double scn = (lane0 && !has_left) ? 0.0 : pow2i(sE - elow);
    if (OUT) sts_s32_if(a_out_e + os * 4, elow, write_out);
    const unsigned bnd = a_in_x + bs * (ST_B * 8);
    const unsigned outp = a_out_x + os * (ST_B * 8);
    const unsigned xr = a_xr + s0 * (ST_RB * CP * 8);
    const unsigned yr = a_yr + (HAS_V ? s0 * (ST_RB * 32 * 8) : 0);
    // the first step reads the previous batch's last row (lane 0: this batch's first boundary entry)
    const unsigned first_addr = lane0 ? bnd : a_xr + ((s0 == 0 ? RS : s0 * ST_RB) - 1) * (CP * 8) - 8u;
    const double scn0 = lane0 ? scn : scn_prev * sc_n;
    scn_prev = scn;
    unsigned nb_addr = lane0 ? bnd + 8u : xr - 8u;

    ST_TICK(tk3);
    // ---- sixteen steps, eight per consumer slot; only ring rows 0..7 have duplicates ----
    if (s0 == 0)
      strip_steps<K, HAS_V, true, CP, RS, true, OUT>(x, ma, nm1, yin, nbp, lane0, first_addr, scn0, scn, nb_addr, nb_stride,
                                                     write_out, xr, yr, outp);
    else
      strip_steps<K, HAS_V, false, CP, RS, true, OUT>(x, ma, nm1, yin, nbp, lane0, first_addr, scn0, scn, nb_addr, nb_stride,
                                                      write_out, xr, yr, outp);
    // mid-batch: fetch the words the NEXT batch's flow control will look at, fix the next scale
    {
      int s1 = s0 + 2;
      if (s1 >= NB) s1 = 0;
      gen_next = lds_v2s32(a_gen + s1 * 4);
      if (has_left) c_in = max(c_in, lds_s32(a_in_written));
      if (has_right) c_out = max(c_out, lds_s32(a_out_taken));
      const int hi = __double2hiint(x[0]);
      e_n = ((hi >> 20) & 0x7ff) - 1023;
      if (x[0] == 0.0) e_n = 0;
      sc_n = pow2i(-e_n);
      elow_n = (int)(E + e_n);
      sE_n = __shfl_up_sync(0xffffffffu, elow_n, 1);
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a_er + (((p + 1) & (ST_NJ - 1)) * 32) * 4),
                   "r"((unsigned)elow_n + (0x80000000u - 1023u)));
    }
    strip_steps<K, HAS_V, false, CP, RS, false, OUT>(x, ma, nm1, yin, nbp, lane0, 0u, scn, scn, nb_addr, nb_stride, write_out,
                                                xr + ST_RB * CP * 8, yr + (HAS_V ? ST_RB * 32 * 8 : 0),
                                                outp + ST_RB * 8);
    ST_TICK(tk4);
    // ---- publish: ONE word says how far the ring is filled; consumers, flusher and loader work
    // out from it what they may touch ----
    __syncwarp();
    release_cta();
    sts_s32_if(a_prog, p, lane0);
    sts_s32_if(a_in_taken, p + g.delta, take_bnd);
    if (OUT) sts_s32_if(a_out_written, p, write_out);
    s0 += 2;
    if (s0 >= NB) {
      s0 = 0;
      gen_need += U;
    }
    ST_TICK(tk5);
    ST_ACC(0, tk2, tk0);   // flow control
    ST_ACC(3, tk3, tk2);   // batch set-up
    ST_ACC(4, tk4, tk3);   // the steps
    ST_ACC(5, tk5, tk4);   // publish
  }
#ifdef STB_PROFILE_PRODUCER
  if (lane0 && P.dbg) {
    long long *d = P.dbg + ((size_t)blockIdx.x * G + (threadIdx.x >> 5)) * 8;
    for (int i = 0; i < 5; i++) d[i] = dbgacc[i];
    d[3] = dbgacc[3] + (dbgacc[6] << 40);  // (the cycles spent waiting for the flusher travel in the upper bits)
    d[6] = g.nbatch;
    unsigned long long gt_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
    d[7] = (long long)(gt_end - gt_start);
    d[2] = clock64() - ck_start;  // (slot 2, "left boundary", moves to the print below: cycles of the whole producer)
    d[5] = dbgacc[5];
  }
#endif
}


// ---- table stores ---------------------------------------------------------------------------------------
/*
 * The consumers' table stores name the GLOBAL state space (st.global, not a generic store: the
 * load/store unit need not resolve the address space).
 */
__device__ __forceinline__ unsigned long long gaddr(const void *p) { return (unsigned long long)__cvta_generic_to_global(p); }
__device__ __forceinline__ unsigned long long row_addr(unsigned long long base, unsigned pitch_bytes, unsigned i) {
  unsigned long long r;  // base + i * pitch as ONE wide multiply-add (i is a constant after unrolling)
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(pitch_bytes), "r"(i), "l"(base));
  return r;
}
__device__ __forceinline__ void stg(unsigned long long a, double v, double *) {
  asm volatile("st.global.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void stg(unsigned long long a, double v, float *) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(a), "f"((float)v) : "memory");
}

// ---- consumer ----------------------------------------------------------------------------------------
/*
 * The consumers are bound by instruction issue (ncu, profiles/README.md: every instruction of the per-unit code costs
 * about the same, taken branches and dependent address arithmetic more), so the loop over units is kept as lean as
 * it gets: shared addresses are 32-bit words formed once, the slow paths (waiting for the producer, the strip's
 * triangle and last rows) are functions of their own, the table stores are predicated instead of branched around,
 * what only the table's first strip does (the S1 column, no V in column 1) is a template flag, and the slot release
 * is one predicated fence + reduction.
 */
/* one more unit of a ring slot is finished: the warp's ring reads are ordered in front of the count (release, cta scope) */
__device__ __forceinline__ void slot_release(unsigned gen_s, int lane) {
  __syncwarp();
#if STB_FENCES
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p fence.acq_rel.cta;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(gen_s), "r"(lane) : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(gen_s), "r"(lane) : "memory");
#endif
}

/* slow path of a consumer's wait for the producer: polls the progress word; returns what it saw last (< need: the fill was aborted) */
__device__ __noinline__ int consumer_wait(unsigned prog_s, int need, int *abort_flag) {
  const long long t0 = clock64();
  unsigned spins = 0;
  int v;
  for (;;) {
    v = lds_s32(prog_s);
    if (v >= need) break;
    __nanosleep(ST_CONS_SLEEP);
    if ((++spins & 1023u) == 0) {
      const int bad = ld_vol(abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
      if (__any_sync(0xffffffffu, bad)) {
        if ((threadIdx.x & 31) == 0) atomicExch(abort_flag, 1);
        break;
      }
    }
  }
  return v;
}

__device__ __forceinline__ void stg_if(unsigned long long a, double v, bool p, double *) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.f64 [%0], %1;\n\t}" ::"l"(a), "d"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void stg_if(unsigned long long a, double v, bool p, float *) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(a), "f"((float)v), "r"((unsigned)p) : "memory");
}

/* a unit on the strip's triangle (column col exists from row r = col on) or on the table's last rows: rare, not inlined */
template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
__device__ __noinline__ void consumer_edge_unit(const StripParams &P, StripSub<K, G, HAS_V> &sb, const unsigned logtab,
                                                const StripGeom &g, int lane, const StripTable &tb, int q, int kk) {
  using Cfg = StripCfg<K, G, HAS_V>;
  constexpr int CP = Cfg::CP, NB = Cfg::NB;
  int cvalid = P.M - g.rs;
  if (cvalid > P.C) cvalid = P.C;
  OutT *tabS = (OutT *)tb.tabS;
  OutT *tabV = (OutT *)tb.tabV;
  const bool first_strip = (g.rs == 0);
  const int r0 = q * ST_RB;
  const int col = lane + 32 * kk;
  if (col >= cvalid) return;
  const int pl = col / K, kq = col - pl * K;
  const int toff = r0 + pl + g.phi;
  const int ub = ((toff >> 3) % NB) * ST_RB + (toff & 7);
  const volatile double *xc = &sb.xring[ub * CP + col];
  const volatile double *yc = &sb.yring[HAS_V ? ub * 32 + pl : 0];
  const volatile unsigned *er = sb.ering;
  const size_t cell0 = (size_t)(g.rs + r0) * P.ld + (size_t)(g.rs + col);  // row n-1 = rs+r, column m-1
  for (int i = 0; i < ST_RB; i++) {
    const int r = r0 + i;
    if (r >= g.R || col > r) continue;
    const double xv = xc[i * CP];
    const size_t off = cell0 + (size_t)i * P.ld;
    if (HAS_S) {
      const int j = (toff + i) >> ST_SH;
      // S^n_n = 1 (the strip's own diagonal: column col in its row col) is stored as exactly +0.0
      const double v = (col == r) ? 0.0 : log_scaled_r<LOGTAB_REP8>(xv, er[(j & (ST_NJ - 1)) * 32 + pl], logtab);
      st_out(tabS + off, v);
      if (first_strip && col == 0) tb.s1[r] = v;
    }
    if (HAS_V && !(first_strip && col == 0)) {
      const double den = (kq == 0) ? yc[i * 32] : xc[i * CP - 1];
      st_out(tabV + off, div_pos(xv, den));
    }
  }
}

template <int K, int G, bool HAS_S, bool HAS_V, typename OutT, bool FIRST>
__device__ void strip_consumer(const StripParams &P, StripSub<K, G, HAS_V> &sb, const unsigned logtab,
                               const StripGeom &g, int lane, const StripTable &tb, const int ci, const int ncs) {
  using Cfg = StripCfg<K, G, HAS_V>;
  constexpr int CP = Cfg::CP, NB = Cfg::NB;
  int cvalid = P.M - g.rs;  // columns of this strip that exist
  if (cvalid > P.C) cvalid = P.C;
  const unsigned pitchb = (unsigned)P.ld * (unsigned)sizeof(OutT);  // row pitch in bytes (strip_fill checks ld * 8 < 2^32)

#ifdef STB_PROFILE_PRODUCER
  long long cacc[4] = {0, 0, 0, 0};
  long long ct_prev = clock64();
#define ST_CTICK(slot)                  \
  {                                     \
    const long long ct_now = clock64(); \
    cacc[slot] += ct_now - ct_prev;     \
    ct_prev = ct_now;                   \
  }
#else
#define ST_CTICK(slot)
#endif
  // The unit of work is 8 rows x 32 columns, dealt to the strip's consumers round-robin (unit u
  // = q*U + kk goes to consumer u % ncs): no claim traffic at all, and the U column blocks of a
  // row batch are worked on by U warps at once, so the batch's ring slot is free again after ONE
  // block's latency -- a whole batch per warp held it ~K times longer and starved the producer of
  // ring space.  Each finished unit counts the slot's release word up by one.
  const int U = (cvalid + 31) >> 5;
  int q = ci / U, kk = ci - q * U;
  const int dq = ncs / U, dk = ncs - dq * U;
  int prog = -1;  // the producer's progress as last seen
  // addresses the loop works with: shared ones as 32-bit words, the table's as 64-bit global addresses of the
  // strip's cell (r = 0, col = 0)
  const unsigned xring_s = smem_u32(&sb.xring[0]), ering_s = smem_u32(&sb.ering[0]);
  const unsigned prog_s = smem_u32(&sb.progress), gen_s = smem_u32(&sb.empty_gen[0]);
  constexpr unsigned ES = (unsigned)sizeof(OutT);
  const unsigned long long gS0 = HAS_S ? gaddr(tb.tabS) + ((unsigned long long)g.rs * P.ld + (unsigned)g.rs) * ES : 0ull;
  const unsigned long long gV0 = HAS_V ? gaddr(tb.tabV) + ((unsigned long long)g.rs * P.ld + (unsigned)g.rs) * ES : 0ull;
  // rows 8q .. 8q+7 of every lane are in the ring once the producer has finished the batch that
  // holds step 8q+7 + (L-1) + phi (the last lane runs L-1 rows behind the first); a partial last
  // row batch is complete with the last producer batch
  const int c_need = ST_RB - 1 + P.L - 1 + g.phi, last_b = g.nbatch - 1;
  const int r_fast_hi = g.R - ST_RB;  // a unit is off the edges when cvalid <= r0 <= R - 8
  for (; q < g.QT;) {
    ST_CTICK(0);
    const int r0 = q * ST_RB;
    const int p_need = min((r0 + c_need) >> ST_SH, last_b);
#ifndef STB_C_CALL_WAIT  // (the wait inlined measured 1-2 % faster than the call on config 2)
    if (!ctr_wait<false, ST_CONS_SLEEP>(&sb.progress, p_need, P.abort_flag, prog)) return;
#else
    if (prog < p_need) {
      prog = consumer_wait(prog_s, p_need, P.abort_flag);
      if (prog < p_need) return;
    }
#endif
    ST_CTICK(1);  // wait for the rows
    // The strip's first cvalid rows are its triangle (column col exists from row r = col on).  Its units take the
    // same unrolled path as the rest, with one more store predicate per row (col <= r); a unit that lies above the
    // diagonal altogether has nothing to do.  (They used to go through the per-cell edge function: ~100 slow units
    // at the start of EVERY strip, and since no CTA can run faster than its left neighbour, what a strip loses at
    // its start it never makes up -- the lost time ADDS UP along the chain of CTAs: 125 independent 160-column
    // tables side by side fill in 8.20 ms, the 125-CTA chain of config 2 took 9.24.)
    const bool tri = r0 < cvalid;
    if (tri && r0 + ST_RB <= 32 * kk) {
      // (every cell of the unit is above the diagonal)
#ifdef STB_EXP_NOCONS  /* timing experiment (wrong results): consumers release their units without touching them */
    } else if (true) {
#endif
#if !defined(STB_EXP_NOCONS)
    } else if (r0 <= r_fast_hi) {
#else
    } else if (false) {
#endif
      // (tried: a variant for strips whose consumers are a multiple of the column blocks -- every warp keeps its
      // block, the column's geometry hoisted out of the loop: 4 % SLOWER on config 2, 9.44 against 9.04 ms)
      const int col = lane + 32 * kk;
      const int pl = col / K;
      const bool lane_ok = col < cvalid;
      // producer lane pl made rows r0..r0+7 at steps toff..toff+7 (ring rows ub..ub+7, never
      // wrapping thanks to the duplicate rows); its exponent changes at most once on the way
      const int toff = r0 + pl + g.phi;
      const int ub = (NB == 16) ? (toff & (16 * ST_RB - 1)) : ((toff >> 3) % NB) * ST_RB + (toff & 7);
      const unsigned xa = xring_s + (unsigned)(ub * CP + col) * 8u;
      // a 1 the compiler cannot see through, and not loop-invariant either: row i+1 of a unit is stored at (address
      // of row i) + pitch * one, ONE wide multiply-add per row (a constant or hoisted factor becomes an add with
      // carry: two instructions)
      const unsigned one = 1u + ((unsigned)q >> 28);
      double xv[ST_RB];
#pragma unroll
      for (int i = 0; i < ST_RB; i++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(xv[i]) : "r"(xa + (unsigned)(i * CP * 8)));
      if (HAS_S) {
        const unsigned ja = ((unsigned)toff >> ST_SH) & (ST_NJ - 1), jb = (ja + 1) & (ST_NJ - 1);
        const unsigned sh = (unsigned)toff & (ST_B - 1);  // row i is of the next batch when i + sh >= 16
        unsigned Ea, Eb;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Ea) : "r"(ering_s + (ja * 32 + (unsigned)pl) * 4u));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Eb) : "r"(ering_s + (jb * 32 + (unsigned)pl) * 4u));
        double v[ST_RB];
#pragma unroll
        for (int i = 0; i < ST_RB; i++) v[i] = log_scaled_r<LOGTAB_REP8>(xv[i], (i + sh >= (unsigned)ST_B) ? Eb : Ea, logtab);
        unsigned long long ga = row_addr(gS0, pitchb, (unsigned)r0) + (unsigned)col * ES;
        if (!tri) {
#pragma unroll
          for (int i = 0; i < ST_RB; i++) {
            stg_if(ga, v[i], lane_ok, (OutT *)nullptr);
            if (i + 1 < ST_RB) ga = row_addr(ga, pitchb, one);
          }
        } else {
          const int lim = col - r0;  // row i of the unit holds column col when i >= lim
#pragma unroll
          for (int i = 0; i < ST_RB; i++) {
            stg_if(ga, v[i], lane_ok && i >= lim, (OutT *)nullptr);
            if (i + 1 < ST_RB) ga = row_addr(ga, pitchb, one);
          }
        }
        if (FIRST && kk == 0) {  // S1: log S^n_1 in FP64 whatever the table stores
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < ST_RB; i++) tb.s1[r0 + i] = v[i];
          }
        }
      }
      if (HAS_V) {
        const int kq = col - pl * K;
        double den[ST_RB];
        // the cell to the left: the lane's own producer lane made it (same units, one column to the left in the x
        // ring) unless the column is its producer lane's first -- then it is in the y ring, converted to the lane's
        // units by the producer.  Two predicated loads with immediate offsets per row.
        const double *xcm = &sb.xring[ub * CP + col - 1], *ycm = &sb.yring[HAS_V ? ub * 32 + pl : 0];
#pragma unroll
        for (int i = 0; i < ST_RB; i++) den[i] = (kq == 0) ? ycm[i * 32] : xcm[i * CP];
        const bool v_ok = FIRST ? (lane_ok && col != 0) : lane_ok;
        unsigned long long ga = row_addr(gV0, pitchb, (unsigned)r0) + (unsigned)col * ES;
        if (!tri) {
#pragma unroll
          for (int i = 0; i < ST_RB; i++) {
            stg_if(ga, div_pos(xv[i], den[i]), v_ok, (OutT *)nullptr);
            if (i + 1 < ST_RB) ga = row_addr(ga, pitchb, one);
          }
        } else {
          const int lim = col - r0;
#pragma unroll
          for (int i = 0; i < ST_RB; i++) {
            stg_if(ga, div_pos(xv[i], den[i]), v_ok && i >= lim, (OutT *)nullptr);
            if (i + 1 < ST_RB) ga = row_addr(ga, pitchb, one);
          }
        }
      }
    } else {
      consumer_edge_unit<K, G, HAS_S, HAS_V, OutT>(P, sb, logtab, g, lane, tb, q, kk);
    }
    slot_release(gen_s + ((unsigned)q % (unsigned)NB) * 4u, lane);
    ST_CTICK(2);  // the unit
    q += dq;
    kk += dk;
    if (kk >= U) {
      kk -= U;
      q++;
    }
#ifdef STB_PROFILE_PRODUCER
    cacc[3] += 1;
#endif
  }
#ifdef STB_PROFILE_PRODUCER
  if (lane == 0 && P.dbg) {
    long long *d = P.dbg + 1024 * 8 + ((size_t)blockIdx.x * ST_MAXW + (threadIdx.x >> 5)) * 4;
    for (int i = 0; i < 4; i++) d[i] = cacc[i];
  }
#endif
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

/* a boundary in global memory: ring of ST_NBG batches with the reader's progress word, or (mask = ~0,
 * taken = NULL) the linear carry buffer between two passes */
struct GBound {
  uint4 *g;
  unsigned mask;
  int *taken;
};

/*
 * A launch may fill several tables per CTA group, one after the other (the rounds of a sweep).  The global
 * ring of a boundary is then ONE stream over the rounds: batch j of round r travels as entry J = jbase + j,
 * jbase = r * (batches of the writing strip), in slot J mod ST_NBG with flag J + 1, and the reader's progress
 * word counts in J as well -- nothing is reset between rounds, and the writer of round r + 1 may start as soon
 * as the ring has room, while the reader is still finishing round r.
 */
__device__ void strip_loader(const StripParams &P, BRing *ring, int delta, int lane, const GBound gb, int jlast, int jbase) {
  // the boundary is written by the CTA to the left (or by the previous pass); batches delta .. jlast are needed here
  const uint4 *g = gb.g;
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
      if (j <= lim && lane < ST_GE) v[u] = ld_volatile_v4(&g[(size_t)((unsigned)(jbase + j) & gb.mask) * ST_GE + lane]);
    }
    int got = 0;
#pragma unroll
    for (int u = 0; u < NP; u++) {
      const int j = next + u;
      const unsigned seq = (unsigned)(jbase + j + 1);
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
      release_cta();
      const int hi = next + got - 1;
      if (lane == 0) {
        st_vol(&ring->written, hi);
        if (gb.taken && (hi - pub >= ST_NBG / 4 || hi == jlast)) st_release_gpu(gb.taken, jbase + hi);  // back-pressure: the entries were read
      }
      if (hi - pub >= ST_NBG / 4) pub = hi;
      next = hi + 1;
      idle = 0;
      t0 = clock64();
    } else {
      __nanosleep(ST_HELPER_SLEEP);  // nothing new: leave the issue slots to the producer next door
      if ((++idle & 1023u) != 0) continue;
      const int bad = ld_vol(P.abort_flag) || (clock64() - t0 > FILL_WATCHDOG);
      if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) atomicExch(P.abort_flag, 1);
        return;
      }
    }
  }
}

/*
 * One strip per CTA: the boundary column of batch j is in the x ring, where the producer's last lane put it
 * for the consumers -- step j*16 + i sits in ring row (2j % NB)*8 + i, column (L-1)*K + K-1 -- and the batch's
 * exponent in the exponent ring (lane L-1).  The flusher follows the producer's progress word and is a reader
 * of the ring slots like the consumers: its own progress (ring->taken) holds the producer back.
 */
template <int K, int G, bool HAS_V>
__device__ void strip_flusher_xring(const StripParams &P, StripSub<K, G, HAS_V> &sb, BRing *ring, int last, int lane,
                                    const GBound gb, int jbase) {
  using Cfg = StripCfg<K, G, HAS_V>;
  uint4 *g = gb.g;
  const unsigned a_col = smem_u32(&sb.xring[(P.L - 1) * K + K - 1]);
  int c_w = -1, c_t = gb.taken ? -0x40000000 : 0x3fffffff;  // the reader's progress counts over the rounds (see strip_loader); unknown yet
  for (int next = 0; next <= last;) {
    if (!ctr_wait<false, ST_HELPER_SLEEP>(&sb.progress, next, P.abort_flag, c_w)) return;
    if (gb.taken && !ctr_wait<true, 64>(gb.taken, jbase + next - ST_NBG, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (gb.taken && hi > c_t - jbase + ST_NBG) hi = c_t - jbase + ST_NBG;
    for (int j = next; j <= hi; j++) {
      const unsigned seq = (unsigned)(jbase + j + 1);
      if (lane < ST_B) {
        const double x = lds_f64(a_col + (unsigned)((((2 * j) % Cfg::NB) * ST_RB + lane) * Cfg::CP * 8));
        st_volatile_v4(&g[(size_t)((unsigned)(jbase + j) & gb.mask) * ST_GE + lane],
                       make_uint4((unsigned)__double2loint(x), seq, (unsigned)__double2hiint(x), seq));
      } else if (lane == ST_B) {
        const unsigned kb = *(volatile unsigned *)&sb.ering[(j & (ST_NJ - 1)) * 32 + (P.L - 1)];
        st_volatile_v4(&g[(size_t)((unsigned)(jbase + j) & gb.mask) * ST_GE + lane],
                       make_uint4(kb - (0x80000000u - 1023u), seq, 0u, seq));
      }
    }
    __syncwarp();
    release_cta();
    if (lane == 0) st_vol(&ring->taken, hi);
    next = hi + 1;
  }
}

__device__ void strip_flusher(const StripParams &P, BRing *ring, int last, int lane, const GBound gb, int jbase) {
  uint4 *g = gb.g;
  int c_w = -1, c_t = gb.taken ? -0x40000000 : 0x3fffffff;  // the carry buffer holds every batch: nothing to wait for
  for (int next = 0; next <= last;) {
    if (!ctr_wait<false, ST_HELPER_SLEEP>(&ring->written, next, P.abort_flag, c_w)) return;
    if (gb.taken && !ctr_wait<true, 64>(gb.taken, jbase + next - ST_NBG, P.abort_flag, c_t)) return;
    int hi = c_w < last ? c_w : last;
    if (gb.taken && hi > c_t - jbase + ST_NBG) hi = c_t - jbase + ST_NBG;
    for (int j = next; j <= hi; j++) {
      const unsigned seq = (unsigned)(jbase + j + 1);
      if (lane < ST_B) {
        const double x = ring->x[(j & (ST_NBR - 1)) * ST_B + lane];
        st_volatile_v4(&g[(size_t)((unsigned)(jbase + j) & gb.mask) * ST_GE + lane],
                       make_uint4((unsigned)__double2loint(x), seq, (unsigned)__double2hiint(x), seq));
      } else if (lane == ST_B) {
        st_volatile_v4(&g[(size_t)((unsigned)(jbase + j) & gb.mask) * ST_GE + lane],
                       make_uint4((unsigned)ring->e[j & (ST_NBR - 1)], seq, 0u, seq));
      }
    }
    __syncwarp();
    release_cta();
    if (lane == 0) st_vol(&ring->taken, hi);
    next = hi + 1;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
/* shared state of a CTA before it starts a table (the logarithm table is loaded once per launch) */
template <int K, int G, bool HAS_V>
__device__ __forceinline__ void cta_reset(StripSmem<K, G, HAS_V> &sm, const StripParams &P, int strip0) {
  using Cfg = StripCfg<K, G, HAS_V>;
  if (threadIdx.x < G) {
    auto &sb = sm.sub[threadIdx.x];
    for (int s = 0; s < Cfg::NB; s++) sb.empty_gen[s] = 0;
    sb.progress = -1;
  }
  // boundary rings start as zeros: a strip without a left neighbour reads its (unused) ring
  for (int i = threadIdx.x; i < (G + 1) * ST_NBR * ST_B; i += blockDim.x) sm.ring[i / (ST_NBR * ST_B)].x[i % (ST_NBR * ST_B)] = 0.0;
  if (threadIdx.x <= G) {
    // ring g feeds strip strip0+g, whose first batch reads batch delta: earlier batches count as taken
    const int rg = threadIdx.x, st = strip0 + rg;
    sm.ring[rg].written = -1;
    sm.ring[rg].taken = (st > 0 && st < P.P) ? strip_geom(P, st).delta - 1 : -1;
  }
}

/* one table of the CTA: every warp plays its role; a role that gives up (watchdog, abort flag) simply returns */
template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
__device__ __forceinline__ void cta_roles(StripSmem<K, G, HAS_V> &sm, const StripParams &P, const StripTable &tb, int group,
                                          int cta, int strip0, int nloc, int round) {
  using Cfg = StripCfg<K, G, HAS_V>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool cta_left = strip0 > 0, cta_right = strip0 + nloc < P.P;
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
      const int nb_writer = strip_geom(P, strip0 - 1).nbatch;
      GBound gb;
      int jbase = 0;
      if (cta > 0) {
        const int bidx = group * P.ctas + cta - 1;
        gb.g = P.gring + (size_t)bidx * (ST_NBG * ST_GE);
        gb.mask = ST_NBG - 1;
        gb.taken = P.gtaken + bidx;
        jbase = round * nb_writer;
      } else {  // first CTA of a later pass
        gb.g = const_cast<uint4 *>(P.carry_in);
        gb.mask = 0xffffffffu;
        gb.taken = NULL;
      }
      strip_loader(P, &sm.ring[0], g.delta, lane, gb, nb_writer - 1, jbase);
    }
  } else if (warp == ST_FLUSHER) {
    if (cta_right) {
      const StripGeom g = strip_geom(P, strip0 + nloc - 1);
      GBound gb;
      int jbase = 0;
      if (cta + 1 < P.ctas) {
        const int bidx = group * P.ctas + cta;
        gb.g = P.gring + (size_t)bidx * (ST_NBG * ST_GE);
        gb.mask = ST_NBG - 1;
        gb.taken = P.gtaken + bidx;
        jbase = round * g.nbatch;
      } else {  // last CTA of a pass that is not the last
        gb.g = P.carry_out;
        gb.mask = 0xffffffffu;
        gb.taken = NULL;
      }
      if (StripCfg<K, G, HAS_V>::FLUSH_FROM_XRING)
        strip_flusher_xring<K, G, HAS_V>(P, sm.sub[0], &sm.ring[nloc], g.nbatch - 1, lane, gb, jbase);
      else
        strip_flusher(P, &sm.ring[nloc], g.nbatch - 1, lane, gb, jbase);
    }
  } else {
    // consumers: dealt round-robin to the CTA's strips.  Consumers fill the sub-partitions without a
    // producer; a producer's own sub-partition takes only `spread` of them, so that the recurrence
    // keeps most of its issue slots and FP64 pipe.
    // consumer index c: warps on the free sub-partitions first, then rows 1..spread of the
    // producers' sub-partitions (the two helper warps excluded)
    int c = -1, ntot = 0;
    {
      int nfree = 0, nshared = 0, mine = -1;
      bool mine_free = false;
      for (int w = G; w < strip_warps(HAS_S, HAS_V); w++) {
        if (w == ST_LOADER || w == ST_FLUSHER) continue;
        const bool fr = (w & 3) >= G;
        if (!fr && G < 4 && (w >> 2) > P.spread) continue;
        if (w == warp) {
          mine = fr ? nfree : nshared;
          mine_free = fr;
        }
        if (fr)
          nfree++;
        else
          nshared++;
      }
      if (mine >= 0) c = mine_free ? mine : nfree + mine;
      ntot = nfree + nshared;
    }
    if (c < 0) return;
    const int gi = c % G, ci = c / G;
    // consumers serving strip gi: indices gi, gi+G, ... below ntot, at most NB-1 and P.ncons of them
    int ncs = (ntot - gi + G - 1) / G;
    if (ncs > Cfg::NB - 1) ncs = Cfg::NB - 1;
    if (ncs > P.ncons) ncs = P.ncons;
#ifdef STB_EXP_PRODONLY
    if (false) {
#else
    if (gi < nloc && ci < ncs) {
#endif
      const StripGeom g = strip_geom(P, strip0 + gi);
      const unsigned ltab = smem_u32(sm.logtab + (lane & (LOGTAB_REP8 - 1))) - LOGTAB_IDX0 * (LOGTAB_REP8 * 8);
      if (g.rs == 0)
        strip_consumer<K, G, HAS_S, HAS_V, OutT, true>(P, sm.sub[gi], ltab, g, lane, tb, ci, ncs);
      else
        strip_consumer<K, G, HAS_S, HAS_V, OutT, false>(P, sm.sub[gi], ltab, g, lane, tb, ci, ncs);
    }
  }
}

/*
 * CTA group q = blockIdx.x / ctas fills tables q, q + ngroups, q + 2 ngroups, ... one after the other (the
 * rounds of a discount sweep in ONE launch: no launch gap, no memset and no pipeline fill / drain between
 * the tables of a group -- a CTA starts on the next table as soon as its own part of the current one is
 * done, while its neighbours to the right are still finishing).  Between two tables the CTA's warps meet at
 * a barrier and the shared state is reset; the global boundary rings run on (strip_loader).
 */
template <int K, int G, bool HAS_S, bool HAS_V, typename OutT>
__global__ void __launch_bounds__(strip_warps(HAS_S, HAS_V) * 32, 1) fill_strip_kernel(const StripParams P) {
  using SM = StripSmem<K, G, HAS_V>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int stop;
  SM &sm = *reinterpret_cast<SM *>(smem_raw);
  const int group = blockIdx.x / P.ctas, cta = blockIdx.x % P.ctas;
  const int strip0 = P.strip_base + cta * G;  // strips are numbered over the whole table, passes included
  int nloc = P.P - strip0;  // strips of this CTA
  if (nloc > G) nloc = G;
  for (int i = threadIdx.x; i < LOGTAB_N * LOGTAB_REP8; i += blockDim.x) sm.logtab[i] = P.logtab[i / LOGTAB_REP8];
  for (int round = 0, table = group; table < P.ntables; ++round, table += P.ngroups) {
    if (round > 0) {
      // every warp is done with the previous table; did anybody give up?  (one thread looks, all follow it)
      __syncthreads();
      if (threadIdx.x == 0) stop = ld_vol(P.abort_flag);
      __syncthreads();
      if (stop) return;
    }
    cta_reset<K, G, HAS_V>(sm, P, strip0);
    __syncthreads();
    const StripTable tb = P.tables[table];
    cta_roles<K, G, HAS_S, HAS_V, OutT>(sm, P, tb, group, cta, strip0, nloc, round);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct StripState {
  uint4 *gring;
  int *gctr;  // [cap] taken, then the abort flag at [cap]
  int cap;    // CTA boundaries the buffers can serve
  double *logtab;  // device, LOGTAB_N doubles
  StripTable *tables;  // device array
  int tables_cap;
  long long *dbg;  // STB_PROFILE_PRODUCER builds only
  uint4 *carry[2];  // boundary column between two passes over the columns (tables wider than one launch)
  size_t carry_cap; // entries each
};

inline void strip_state_free(StripState *st) {
  cudaFree(st->gring);
  cudaFree(st->gctr);
  cudaFree(st->logtab);
  cudaFree(st->tables);
  cudaFree(st->dbg);
  cudaFree(st->carry[0]);
  cudaFree(st->carry[1]);
  memset(st, 0, sizeof *st);
}

inline size_t strip_state_bytes(const StripState *st) {
  return (size_t)st->cap * ST_NBG * ST_GE * sizeof(uint4) + (st->cap ? ((size_t)st->cap + 1) * sizeof(int) : 0) + (st->logtab ? LOGTAB_N * sizeof(double) : 0) +
         (size_t)st->tables_cap * sizeof(StripTable);
}

inline cudaError_t strip_state_prepare(StripState *st, int nbound, int ntables) {
  cudaError_t e;
  if (!st->logtab) {
    if ((e = cudaMalloc(&st->logtab, LOGTAB_N * sizeof(double))) != cudaSuccess) return e;
    logtab8_build_kernel<<<(LOGTAB_N + 127) / 128, 128>>>(st->logtab);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
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
  return cudaLaunchCooperativeKernel((void *)kern, dim3(nctas), dim3(strip_warps(HAS_S, HAS_V) * 32), args, smem, stream);
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
 * shapes are (K columns per lane, G strips per CTA) = (5,1), (7,1), (1,4), (3,2): up to 160, 224,
 * 128 and 192 columns per CTA, tried in that order (measured best first).
 */
inline bool strip_plan(unsigned M, int slots, size_t elem_size, StripPlan *pl, bool multipass = false) {
  // measured best first: ONE strip per CTA with 5 (up to 160 columns) or 7 (up to 224) columns per
  // lane; the multi-strip shapes are kept for STB_STRIP_K experiments and the geometry-independence test
  static const int shapes[4][2] = {{5, 1}, {7, 1}, {1, 4}, {3, 2}};
  int force_k = 0, force_l = 0;
  if (const char *s = getenv("STB_STRIP_K")) force_k = atoi(s);
  if (const char *s = getenv("STB_STRIP_L")) force_l = atoi(s);
  if (slots < 1) slots = 1;
  const unsigned per_cta = (M + (unsigned)slots - 1) / (unsigned)slots;  // columns per CTA with every slot in use
  for (int i = 0; i < 4; i++) {
    const int k = shapes[i][0], gg = shapes[i][1];
    // several passes over the columns: the widest single-strip shape, however many CTAs it takes
    if (multipass ? (force_k ? k != force_k : !(k == 7 && gg == 1)) : (force_k ? k != force_k : (unsigned)(32 * k * gg) < per_cta))
      continue;
    int L;
    if (gg == 1) {
      // Full warps and as FEW CTAs as the table needs beat an even spread over all SMs: the row rate
      // is the producer's whatever the lane count, the consumers' 32-column units come out full,
      // and every CTA boundary less is one hop less of pipeline lag (config 2: 143 x 140 columns
      // 9.45 ms, 125 x 160 columns 9.16 ms; config 1: 8 CTAs of 4 x 32 columns 1.29 ms, 7 x 160 0.63 ms).
      L = (int)((M + (unsigned)k - 1) / (unsigned)k);
    } else {
      unsigned want = (per_cta + (unsigned)gg - 1) / (unsigned)gg;  // columns per strip
      if (want < 32u) want = M < 32u ? M : 32u;  // never narrower than a warp unless the table is
      L = (int)((want + (unsigned)k - 1) / (unsigned)k);
    }
    if (L > 32) L = 32;
    if (L < 1) L = 1;
    while (L < 32 && ((size_t)L * k * elem_size) % 32 != 0) L++;  // strip edges on 32-byte sectors
    if (force_l) L = force_l;
    pl->K = k;
    pl->G = gg;
    pl->L = L;
    pl->C = L * k;
    pl->P = (int)((M + (unsigned)pl->C - 1) / (unsigned)pl->C);
    pl->ctas = (pl->P + gg - 1) / gg;
    if (pl->ctas > slots && !multipass) continue;
    return true;
  }
  return false;
}

/*
 * Tables one launch fills side by side.  A table of more than one CTA is cut into 160-column strips (K = 5: the
 * 16-slot ring) unless 224-column strips (K = 7: 12 slots, ~136 cycles per row against ~80) fit so many more
 * tables into the launch that they win anyway; measured on config 3 (M = 5000): 4 tables x 32 CTAs 4.54e11 cells/s,
 * 6 x 23 CTAs 4.26e11.  A table that fits ONE 224-column CTA stays there (the samplers' 5001 x 167: 2.9e11 against
 * 2.2e11 as two 84-column CTAs).
 */
inline int strip_tables_per_launch(unsigned M, int num_sms) {
  const int p7 = (int)((M + 32u * 7u - 1) / (32u * 7u)), p5 = (int)((M + 32u * 5u - 1) / (32u * 5u));
  const int t7 = num_sms / (p7 > 0 ? p7 : 1) > 0 ? num_sms / (p7 > 0 ? p7 : 1) : 1;
  const int t5 = num_sms / (p5 > 0 ? p5 : 1) > 0 ? num_sms / (p5 > 0 ? p5 : 1) : 1;
  if (p7 <= 1) return t7;
  return (t5 * 136 >= t7 * 80) ? t5 : t7;
}

struct StripFillArgs {
  const StripTable *tables;  // HOST array of ntables entries
  int ntables;
  int has_S, has_V, is_float;
  size_t ld;
  unsigned N, M;
  int num_sms;
  // NULL: strip_fill waits for the fill and checks the watchdog flag itself.  Otherwise (PINNED host
  // int, and only when the fill is a single launch) the launch is left in flight: the flag is copied
  // there asynchronously and the caller checks it after its own synchronisation -- a sweep queues
  // wave after wave without a host round trip per wave.
  int *async_flag;
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
  int slots = A.num_sms / per_launch;
  if (const char *sv = getenv("STB_STRIP_SLOTS")) {  // testing: fewer CTAs per launch (forces several passes)
    const int v = atoi(sv);
    if (v > 0 && v < slots) slots = v;
  }
  int passes = 1;
  if (!strip_plan(A.M, slots, A.is_float ? 4 : 8, &pl)) {
    // wider than one launch's CTAs can hold: several passes over the columns, the boundary column
    // between two passes kept for all rows in a linear buffer (one table at a time only)
    if (A.ntables > 1 || !strip_plan(A.M, slots, A.is_float ? 4 : 8, &pl, true)) {
      snprintf(err, errlen, "strip_fill: M=%u does not fit %d CTAs per table", A.M, slots);
      return -1;
    }
    passes = (pl.ctas + slots - 1) / slots;
  }
  const int ctas_total = pl.ctas, ctas_pass = passes > 1 ? slots : pl.ctas;
  cudaError_t e = strip_state_prepare(st, per_launch * ctas_pass, A.ntables);
  if (e == cudaSuccess && passes > 1) {
    const size_t need = ((size_t)((A.N + 32u + 32u) >> ST_SH) + 2) * ST_GE;
    if (need > st->carry_cap) {
      cudaFree(st->carry[0]);
      cudaFree(st->carry[1]);
      st->carry[0] = st->carry[1] = NULL;
      st->carry_cap = 0;
      e = cudaMalloc(&st->carry[0], need * sizeof(uint4));
      if (e == cudaSuccess) e = cudaMalloc(&st->carry[1], need * sizeof(uint4));
      if (e == cudaSuccess) st->carry_cap = need;
    }
  }
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
  P.ctas = ctas_pass;
  P.strip_base = 0;
  P.carry_in = NULL;
  P.carry_out = NULL;
  P.gring = st->gring;
  P.gtaken = st->gctr;
  P.abort_flag = st->gctr + st->cap;
  P.logtab = st->logtab;
  P.ncons = 64;
  if (const char *s = getenv("STB_STRIP_CONS")) P.ncons = atoi(s);
  P.spread = 1;  // consumer warps allowed on a producer's sub-partition (measured best: 1)
  if (const char *s = getenv("STB_STRIP_SPREAD")) P.spread = atoi(s);
  P.dbg = NULL;
#ifdef STB_PROFILE_PRODUCER
  if (!st->dbg) cudaMalloc(&st->dbg, (1024 * 8 + 1024 * ST_MAXW * 4) * sizeof(long long));
  cudaMemsetAsync(st->dbg, 0, (1024 * 8 + 1024 * ST_MAXW * 4) * sizeof(long long), stream);
  P.dbg = st->dbg;
#endif
  // ONE launch fills all the tables (CTA group q takes tables q, q + per_launch, ...: the kernel's rounds);
  // only a table wider than a launch's CTAs takes several launches, one per pass over the columns
  for (int pass = 0; pass < passes && e == cudaSuccess; ++pass) {
    const int nt = per_launch;  // CTA groups
    const bool last_launch = pass + 1 == passes;
    P.tables = st->tables;
    P.ntables = A.ntables;
    P.ngroups = nt;
    if (passes > 1) {
      P.strip_base = pass * ctas_pass * pl.G;
      P.ctas = (ctas_total - pass * ctas_pass < ctas_pass) ? ctas_total - pass * ctas_pass : ctas_pass;
      P.carry_in = pass > 0 ? st->carry[(pass - 1) & 1] : NULL;
      P.carry_out = pass + 1 < passes ? st->carry[pass & 1] : NULL;
      if (P.carry_out) e = cudaMemsetAsync(P.carry_out, 0, st->carry_cap * sizeof(uint4), stream);
      if (e != cudaSuccess) break;
    }
    // reader progress starts at -1 ("nothing taken"), the abort flag at 0, the ring's flags at 0
    e = cudaMemsetAsync(st->gctr, 0xFF, (size_t)st->cap * sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.abort_flag, 0, sizeof(int), stream);
    if (e == cudaSuccess)
      e = cudaMemsetAsync(st->gring, 0, (size_t)nt * P.ctas * ST_NBG * ST_GE * sizeof(uint4), stream);
    if (e != cudaSuccess) break;
    const int nctas = nt * P.ctas;
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
      if (last_launch) cudaEventRecord(ev_end, stream);
      if (A.async_flag && passes == 1) {
        e = cudaMemcpyAsync(A.async_flag, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
        break;
      }
      e = cudaMemcpyAsync(&flag, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
#ifdef STB_PROFILE_PRODUCER
      if (e == cudaSuccess && getenv("STB_PROFILE_PRINT")) {
        static long long h[1024 * 8 + 1024 * ST_MAXW * 4];
        cudaMemcpy(h, st->dbg, sizeof h, cudaMemcpyDeviceToHost);
        for (int c : {0, nctas / 2}) {
          for (int w = 0; w < ST_MAXW; w++) {
            const long long *d = h + 1024 * 8 + ((size_t)c * ST_MAXW + w) * 4;
            if (c < 1024 && d[3] > 0)
              fprintf(stderr, "cta %3d consumer warp %2d units %6lld cycles/unit: claim %.0f wait %.0f work %.0f\n", c, w, d[3],
                      (double)d[0] / d[3], (double)d[1] / d[3], (double)d[2] / d[3]);
          }
        }
#ifdef STB_PROFILE_TRACE
        for (int c = 0; c < 4; c++) {
          fprintf(stderr, "cta %d first batches start at (us):", c);
          for (int b = 0; b < 24; b++) fprintf(stderr, " %.2f", h[900 * 8 + c * 24 + b] / 1e3);
          fprintf(stderr, "\n");
          fprintf(stderr, "cta %d last batches start at (us):", c);
          for (int b = 0; b < 24; b++) fprintf(stderr, " %.2f", h[900 * 8 + 96 + c * 24 + b] / 1e3);
          fprintf(stderr, "\n");
        }
#endif
        const int show[10] = {0, 1, 2, 3, nctas / 4, nctas / 2, nctas / 2 + 1, 3 * nctas / 4, nctas - 2, nctas - 1};
        for (int si = 0; si < 10; si++) {
          const int c = show[si];
          if (c < 0 || c >= nctas || (si && c <= show[si - 1])) continue;
          for (int gi = 0; gi < pl.G; gi++) {
            const long long *d = h + ((size_t)c * pl.G + gi) * 8;
            const double nb = (double)d[6];
            if (nb <= 0 || (size_t)c * pl.G + gi >= 1024) continue;
            fprintf(stderr,
                    "cta %3d producer %d batches %6.0f cycles/batch: flow-control %.0f - %.0f - %.0f setup %.0f "
                    "steps %.0f publish %.0f | busy %.1f us\n",
                    c, gi, nb, d[0] / nb, d[1] / nb, d[2] / nb, (d[3] & ((1ll << 40) - 1)) / nb, d[4] / nb, d[5] / nb, d[7] / 1e3);
            fprintf(stderr, "          (first batch started at %.1f us; flusher %.0f; %.0f cycles per batch in all, SM clock %.0f MHz)\n",
                    d[1] / 1e3, (d[3] >> 40) / nb, d[2] / nb, d[2] / (d[7] / 1e3));
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
