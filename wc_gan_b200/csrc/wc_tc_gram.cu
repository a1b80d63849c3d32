// K1 / K4 on the tensor cores (tcgen05.mma, MN-major operands, accumulator in TMEM) for C in {128, 256}:
//
//   Out[slot][i][j] += sum_rows (A[r,i] - ma[i]) * (B[r,j] - mb[j])          (full C x C per slot)
//   asum[slot][i]   += sum_rows (A[r,i] - ma[i]),   bsum[slot][j] += sum_rows (B[r,j] - mb[j])   (optional)
//
// K1 (batch covariance):  A = B = x, shift = pivot  ("self" mode: one converted tile feeds both operands; the
//                         pivot -- mean of the first 32 rows -- is computed by every CTA from the same rows).
// K4 (backward reduce):   A = x - mean, B = dy, one accumulator per class; samples are visited sorted by class.
//
// Each fp32 operand is split into bf16 hi + lo and three products (hi.hi, hi.lo, lo.hi) are accumulated in
// fp32 in TMEM.  The rounding of the split is unbiased, so its effect on a sum over M rows shrinks like
// 1/sqrt(M); the dropped lo.lo term biases the diagonal by ~1e-6 relative ("Accuracy").
//
// One persistent CTA per SM (576 threads) owns a contiguous range of row units.  Three pipelines:
//   warp  16   producer (one thread): 1-D cp.async.bulk of the rows of a step (contiguous in NHWC) into the
//              fp32 staging ring (128 KB in flight per SM)
//   warps 0-15 converters: ld.shared fp32 -> shift -> bf16 hi/lo -> MN-major SWIZZLE_128B operand tiles,
//              column sums on the fly; at a slot change they drain the accumulator (tcgen05.ld -> red.global.add)
//   warp  17   MMA issuer (one thread), owns TMEM
#include "wc_common.cuh"
#include "wc_tc.cuh"
#include <cuda.h>

using namespace wctc;

namespace {

constexpr int KR = 16;                 // rows (K of the MMA) per stage
constexpr int NCONV_WARPS = 16;
constexpr int NCONV = NCONV_WARPS * 32;
constexpr int NTHREADS = (NCONV_WARPS + 2) * 32;
constexpr int UNIT_ROWS = 64;          // a unit never crosses a sample boundary
constexpr int PIVOT_ROWS = 32;

struct TcGramParams {
  const float* A; const float* B;       // [rows][C]
  const float* ma; const float* mb;     // shifts (nullable); self mode: ignored (pivot computed in the kernel)
  int self;                             // 1: B == A (same shift): convert once; only blocks j <= i inside a group are flushed
  int g;                                // self mode: group size (flush only what the finalize kernel reads)
  int N, HW, ups;                       // samples, rows per sample, units per sample
  long long units_per_cta;
  const int32_t* order; const int32_t* slot_of;
  float* out; long long out_slot_stride;          // [slot][C][C]
  float* asum; float* bsum; long long sum_slot_stride;
  float* pivot_out;                     // self mode: the pivot every CTA used (written by CTA 0)
  int nslots;                           // number of [C][C] accumulators behind `out`
  int l2_load;                          // L2 eviction priority of the activation loads (wctc::l2_policy)
  int late_wait;                        // 1: the predecessor on the stream is zero_kernel (see the kernel): only the flush waits
  unsigned long long* dbg; int dbg_flags;   // debug marks (nullptr in normal use)
};

// Shared memory: the fp32 staging ring takes everything the operand ring leaves (the ring is what keeps HBM busy: a slab
// lives from the issue of its bulk load until a converter has it in registers, ~2 us under load, so the bytes in flight
// per SM are ring size / 2 us and have to reach HBM's 44 GB/s per SM).  Self mode stages one operand per step, so its
// operand stages are half as large.  The column-sum scratch and the
// 4 KB boxes of the flush alias a ring that is idle when they are used: the operand ring (every MMA has been committed)
// or, in self mode -- which only ever flushes once, at the end -- the staging ring.
#ifndef WC_GRAM_NOPS_SELF
#define WC_GRAM_NOPS_SELF 3
#endif
#ifndef WC_GRAM_NOPS_CROSS
#define WC_GRAM_NOPS_CROSS 3
#endif
template <int C, bool SELF>
struct GCfg {
  // A pipeline step (one trip through the producer / converter / MMA hand-shakes) costs ~0.4 us of mbarrier round trips
  // whatever its size, so a step carries NSUB sixteen-row MMA slices: 32 KB of rows where shared memory allows it
  // (measured at C = 128, 16-row steps: K1 2.5 TB/s, K4 4.0 TB/s while the apply kernels run at 5.6-6.2 TB/s).
  static constexpr int NSUB = SELF ? (C == 128 ? 4 : 2) : (C == 128 ? 2 : 1);
  static constexpr int RS = KR * NSUB;                     // rows per step
  static constexpr int SLOT = RS * C * 4;                  // one fp32 staging slab
  static constexpr int IMG = KR * C * 2;                   // one bf16 split of one operand, 16 rows: 8 KB (C=256)
  static constexpr int MH = C / 128;                       // M halves
  static constexpr int TMEM_COLS = MH * C;                 // 512 (C=256) / 128
  static constexpr int SUBSTAGE = (SELF ? 2 : 4) * IMG;    // operand images of one 16-row slice
  static constexpr int OPSTAGE = NSUB * SUBSTAGE;
  static_assert(UNIT_ROWS % RS == 0, "a unit is a whole number of steps");
  static constexpr int FLUSH_BYTES = NCONV_WARPS * 4096 + NCONV_WARPS * C * 4;     // boxes + column sums
  static constexpr int NOPS_X = (FLUSH_BYTES + OPSTAGE - 1) / OPSTAGE;             // cross: the flush scratch fits the ring
  static constexpr int NOPS = SELF ? WC_GRAM_NOPS_SELF : (NOPS_X > WC_GRAM_NOPS_CROSS ? NOPS_X : WC_GRAM_NOPS_CROSS);
  static constexpr int BUDGET = 231424 - 1024 - 1024;      // 226 KB per CTA minus alignment slack and static barriers
  static constexpr int NSLOTS_FIT = (BUDGET - NOPS * OPSTAGE) / SLOT;
  static constexpr int NSLOTS = NSLOTS_FIT > 16 ? 16 : NSLOTS_FIT;
  static constexpr int SMEM = NSLOTS * SLOT + NOPS * OPSTAGE + 1024;
  static constexpr int CQ = C / 4, RG = NCONV / CQ;        // pivot: threads per row, row groups
  static_assert(NOPS >= 3, "operand ring too shallow");
  static_assert(!SELF || NSLOTS * SLOT >= FLUSH_BYTES, "self mode: flush scratch lives in the staging ring");
  static_assert(!SELF || NOPS * OPSTAGE >= RG * C * 4, "self mode: pivot scratch lives in the operand ring");
};

struct Step { int sample, row0, nrows, slot; bool flush_before; bool valid; };

struct StepIter {
  const TcGramParams& p;
  long long units_left;
  int sidx, sub, r_in_unit, cur_slot, sample, slot, rs;        // rs: rows per step (a multiple of KR that divides UNIT_ROWS)
  __device__ StepIter(const TcGramParams& p_, long long ub, long long ue, int rs_) : p(p_), units_left(ue - ub), r_in_unit(0), cur_slot(-1), rs(rs_) {
    sidx = (int)(ub / p.ups); sub = (int)(ub % p.ups);    // the only division: next() advances incrementally
    load_sample();
  }
  __device__ void load_sample() {
    sample = p.order ? p.order[sidx] : sidx;
    slot = p.slot_of ? p.slot_of[sample] : 0;
  }
  __device__ Step next() {
    Step s{};
    if (units_left <= 0) { s.valid = false; return s; }
    const int urow0 = sub * UNIT_ROWS;
    const int unrows = min(UNIT_ROWS, p.HW - urow0);
    s.sample = sample;
    s.row0 = urow0 + r_in_unit;
    s.nrows = min(rs, unrows - r_in_unit);
    s.slot = slot;
    s.flush_before = (cur_slot >= 0 && s.slot != cur_slot);
    s.valid = true;
    cur_slot = s.slot;
    r_in_unit += rs;
    if (r_in_unit >= unrows) {
      r_in_unit = 0; --units_left;
      if (++sub == p.ups) { sub = 0; ++sidx; if (units_left > 0) load_sample(); }
    }
    return s;
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int C, bool SELF>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gram_kernel(TcGramParams p, const __grid_constant__ CUtensorMap tm_out) {
  using cfg = GCfg<C, SELF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stg = smem;                                   // fp32 staging ring
  uint8_t* ops = smem + cfg::NSLOTS * cfg::SLOT;         // bf16 operand ring
  __shared__ uint64_t stg_full[cfg::NSLOTS], stg_empty[cfg::NSLOTS], op_full[cfg::NOPS], op_empty[cfg::NOPS], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  uint8_t* scratch = SELF ? stg : ops;                   // flush: 16 boxes of 4 KB, then the column-sum reduction
  float (*sred)[C] = reinterpret_cast<float (*)[C]>(scratch + NCONV_WARPS * 4096);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long total_units = (long long)p.N * p.ups;
  const long long ub = (long long)blockIdx.x * p.units_per_cta;
  const long long ue = min(total_units, ub + p.units_per_cta);
  if (ub >= ue) return;   // uniform per CTA

  if (t == 0) {
    for (int s = 0; s < cfg::NSLOTS; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], NCONV_WARPS / 2); }
    for (int s = 0; s < cfg::NOPS; ++s) { mbar_init(&op_full[s], NCONV_WARPS / 2); mbar_init(&op_empty[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, NCONV_WARPS);
    mbar_fence_init();
    tma_prefetch_desc(&tm_out);
  }
  if (warp == NCONV_WARPS + 1) { tmem_alloc(&tmem_slot, cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_slot;
  constexpr int nops = SELF ? 1 : 2;
  // The kernel right before this one is the one that zeroes the accumulators (zero_kernel; it has itself waited for
  // whatever produced the activations before it let this grid start).  Nothing but the flush touches the accumulators,
  // so only the flush waits for it: the whole main loop runs under the zeroing and its launch latency.  When the zeroing
  // was a memset (unaligned workspace, fine-grained ABI with separate P and s) the wait comes first as everywhere else.
  if (!p.late_wait) wc_pdl_wait();
  if (t == 0) wc_mark(p.dbg, 0);

  if (warp == NCONV_WARPS) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      StepIter iter(p, ub, ue, cfg::RS);
      uint32_t pi = 0;
      const uint64_t pol = l2_policy(p.l2_load);
      for (Step st = iter.next(); st.valid; st = iter.next()) {
        const long long off = ((long long)st.sample * p.HW + st.row0) * C;
        const uint32_t bytes = (uint32_t)st.nrows * C * 4;
        for (int op = 0; op < nops; ++op, ++pi) {
          const int sl = pi % cfg::NSLOTS;
          mbar_wait(&stg_empty[sl], ((pi / cfg::NSLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&stg_full[sl], bytes);
          bulk_g2s_hint(stg + sl * cfg::SLOT, (op == 0 ? p.A : p.B) + off, bytes, &stg_full[sl], pol);
        }
      }
    }
  } else if (warp == NCONV_WARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_full = make_idesc_bf16(128, C, 1, 1);
      // self mode, first M half: only the columns j < 128 are ever read back (the scatter is symmetric)
      const uint32_t idesc_half = make_idesc_bf16(128, 128, 1, 1);
      StepIter iter(p, ub, ue, cfg::RS);
      uint32_t acc = 0, nflush = 0;
      int it = 0;
      for (Step st = iter.next(); st.valid; st = iter.next(), ++it) {
        if (st.flush_before) {
          umma_commit(&acc_full);
          mbar_wait(&acc_empty, nflush & 1);
          tc_fence_after();
          ++nflush;
          acc = 0;
        }
        const int s = it % cfg::NOPS;
        mbar_wait(&op_full[s], (it / cfg::NOPS) & 1);
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < cfg::NSUB; ++sub) {
          if (sub * KR >= st.nrows) break;             // a short step (end of a unit): the remaining slices hold no rows
          const uint32_t a_hi = smem_u32(ops + s * cfg::OPSTAGE + sub * cfg::SUBSTAGE);
          const uint32_t a_lo = a_hi + cfg::IMG;
          const uint32_t b_hi = SELF ? a_hi : a_hi + 2 * cfg::IMG;
          const uint32_t b_lo = b_hi + cfg::IMG;
          const uint64_t dbh = make_desc(b_hi, KR * 128, 1024, LAYOUT_SW128);
          const uint64_t dbl = make_desc(b_lo, KR * 128, 1024, LAYOUT_SW128);
#pragma unroll
          for (int mh = 0; mh < cfg::MH; ++mh) {
            const uint32_t aoff = mh * 2 * (KR * 128);   // M half = two 64-channel blocks
            const uint64_t dah = make_desc(a_hi + aoff, KR * 128, 1024, LAYOUT_SW128);
            const uint64_t dal = make_desc(a_lo + aoff, KR * 128, 1024, LAYOUT_SW128);
            const uint32_t tdm = td + mh * C;
            const uint32_t idesc = (SELF && cfg::MH == 2 && mh == 0) ? idesc_half : idesc_full;
            if (p.dbg_flags & 4) continue;
            umma_bf16(tdm, dah, dbh, idesc, acc);
            umma_bf16(tdm, dah, dbl, idesc, 1);
            umma_bf16(tdm, dal, dbh, idesc, 1);
          }
          acc = 1;
        }
        umma_commit(&op_empty[s]);
      }
      umma_commit(&acc_full);
    }
  } else {
    // ------------------------------------------------------------------ converters
    // Two groups of 8 warps take alternate steps (K4: a group converts both operands of its step), so that the hand-shake
    // latencies of one step run under the conversion of the other.  Within a group a warp owns 2 rows of a sixteen-row
    // slice; a lane owns the 4-channel quads {lane, lane+32} of a row (conflict-free ld.shared.v4 / st.shared.v2).
    const int grp = warp >> 3, wg = warp & 7;
    const uint32_t stg_u32 = smem_u32(stg), ops_u32 = smem_u32(ops);
    constexpr int QPR = C / 128;              // quads per lane per row
    // cross mode: A is shifted by ma and has no column sums, B is not shifted and has them (what K4 needs; the launcher
    // refuses anything else), so a thread keeps one shift and one set of sums whichever operand it converts
    const bool want = SELF ? (p.asum != nullptr) : (p.bsum != nullptr);
    float sum[QPR][4];
    float4 mean[QPR];
#pragma unroll
    for (int j = 0; j < QPR; ++j) { sum[j][0] = sum[j][1] = sum[j][2] = sum[j][3] = 0.f; mean[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (SELF) {
      // pivot = mean of the first rows: identical in every CTA (same rows, same order of operations); scratch: the operand
      // ring, which nobody writes before the last barrier below
      float (*spiv)[C] = reinterpret_cast<float (*)[C]>(ops);
      const int prow = min(PIVOT_ROWS, p.HW);
      {
        const int cq = t % cfg::CQ, rg = t / cfg::CQ;
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < PIVOT_ROWS / cfg::RG; ++k) {
          const int r = rg * (PIVOT_ROWS / cfg::RG) + k;
          if (r < prow) {
            const float4 q4 = __ldg(reinterpret_cast<const float4*>(p.A + (long long)r * C + cq * 4));
            acc4.x += q4.x; acc4.y += q4.y; acc4.z += q4.z; acc4.w += q4.w;
          }
        }
        *reinterpret_cast<float4*>(&spiv[rg][cq * 4]) = acc4;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
      for (int ch = t; ch < C; ch += NCONV) {
        float s = 0.f;
#pragma unroll
        for (int rr = 0; rr < cfg::RG; ++rr) s += spiv[rr][ch];
        s /= (float)prow;
        spiv[0][ch] = s;      // only this thread touches column ch
        if (blockIdx.x == 0 && p.pivot_out) p.pivot_out[ch] = s;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
#pragma unroll
      for (int j = 0; j < QPR; ++j) mean[j] = *reinterpret_cast<const float4*>(&spiv[0][(lane + 32 * j) * 4]);
      asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
    } else {
      const float* m = p.ma;
      if (m) {
#pragma unroll
        for (int j = 0; j < QPR; ++j) mean[j] = *reinterpret_cast<const float4*>(m + (lane + 32 * j) * 4);
      }
    }
    const int r0 = wg * 2;                                        // this warp's rows of a slab: r0, r0 + 1
    const uint32_t ld_off = (uint32_t)r0 * (C * 4) + (uint32_t)lane * 16;
    // MN-major SWIZZLE_128B: quad qd of k-row r -> block qd/16, 16-byte chunk (qd%16)/2 XOR (r%8), half qd%2
    uint32_t st_off[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = r0 + rr;
      st_off[rr] = (uint32_t)(lane >> 4) * (KR * 128) + (uint32_t)r * 128 + (uint32_t)((((lane & 15) >> 1) ^ (r & 7)) << 4) + (uint32_t)(lane & 1) * 8;
    }

    bool dep_waited = !p.late_wait;
    auto flush = [&](int slot, uint32_t parity) {
      // accumulator -> global (vector reductions); column sums -> global
      if (!dep_waited) { wc_pdl_wait(); dep_waited = true; }
      mbar_wait(&acc_full, parity);
      tc_fence_after();
      const int q = warp & 3, part = warp >> 2;            // TMEM lanes 32q..32q+31; a quarter of the columns
      constexpr int CW = C / 4, NCH = CW / 32;             // columns per warp, 32-column chunks per warp
      float* out = p.out + (long long)slot * p.out_slot_stride;
      // CTAs start at different chunks so that at any moment they add into different L2 lines
      const int nsteps = cfg::MH * NCH, rot = (int)(blockIdx.x % nsteps);
      // The accumulator leaves as 32 x 32 boxes: tcgen05.ld -> swizzled 4 KB tile in this warp's slice of the (idle)
      // operand ring -> one TMA reduce-add into the global matrix.  The L2 adds whole 128-byte lines instead of
      // 16-byte vector reductions.  (experiment bit 16: the older red.global.add.v4 path)
      const bool use_tma = !(p.dbg_flags & 16);
      const uint32_t fbuf = smem_u32(scratch) + (uint32_t)warp * 4096u;
      const uint32_t frow = fbuf + (uint32_t)lane * 128u;
#pragma unroll 1
      for (int k = 0; k < nsteps; ++k) {
        const int kk = (k + rot) % nsteps;
        const int mh = kk / NCH, c0 = part * CW + (kk % NCH) * 32;
        const int i = mh * 128 + q * 32 + lane;
        const int i0 = mh * 128 + q * 32;
        int clo = 0, chi = C;                      // self mode: only blocks the finalize kernel reads (j <= i, same group)
        if (SELF) { clo = (i0 / p.g) * p.g; chi = min(((i0 + 31) / p.g + 1) * p.g, i0 + 32); }
        if (c0 >= chi || c0 + 32 <= clo) continue;   // warp-uniform
        float acc[32];
        if (use_tma) {
          if (lane == 0) bulk_wait_group_read<0>();   // the previous box has left the tile
          __syncwarp();
        }
        tmem_ld32(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * C + c0), acc);
        if (use_tma) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(frow + ((j ^ (lane & 7)) << 4)), "f"(acc[4 * j]),
                         "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3]) : "memory");
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.dbg_flags & 2)) {
            tma_reduce_add_3d(&tm_out, c0, i0, slot, fbuf);
            bulk_commit_group();
          }
        } else {
          float* dst = out + (long long)i * C + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) if (!(p.dbg_flags & 2)) red_add_v4(dst + j, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      }
      if (use_tma) {
        // the tiles live in the operand ring: nobody may refill it before every box has been read out
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      if (use_tma) asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
      // column sums: every warp holds partial sums of all channels of its operand
      for (int pass = 0; pass < 2; ++pass) {
        float* gsum = pass == 0 ? p.asum : p.bsum;
        if (!gsum || (pass == 1 && SELF)) continue;
        const bool mine = SELF ? (pass == 0) : (pass == 1);
        asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
#pragma unroll
        for (int j = 0; j < QPR; ++j) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mine) { v = make_float4(sum[j][0], sum[j][1], sum[j][2], sum[j][3]); sum[j][0] = sum[j][1] = sum[j][2] = sum[j][3] = 0.f; }
          *reinterpret_cast<float4*>(&sred[warp][(lane + 32 * j) * 4]) = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
        for (int ch = t; ch < C; ch += NCONV) {
          float s = 0.f;
#pragma unroll
          for (int rr = 0; rr < NCONV_WARPS; ++rr) s += sred[rr][ch];
          atomicAdd(gsum + (long long)slot * p.sum_slot_stride + ch, s);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCONV));     // the scratch is ring memory: nobody converts into it before this
      }
    };

    StepIter iter(p, ub, ue, cfg::RS);
    uint32_t nflush = 0;
    int it = 0, cur_slot = -1;
    for (Step st = iter.next(); st.valid; st = iter.next(), ++it) {
      if (st.flush_before) { flush(cur_slot, nflush & 1); ++nflush; }
      cur_slot = st.slot;
      if ((it & 1) != grp) continue;                 // the groups take alternate steps: two steps are in conversion at a time
      const int s = it % cfg::NOPS;
      // (the operand stage is claimed before the rows are read: the words of a sixteen-row slice then go straight from the
      // conversion into the tile, and a step of several slices needs no more registers than one)
      mbar_wait(&op_empty[s], ((it / cfg::NOPS) & 1) ^ 1);
#pragma unroll
      for (int op = 0; op < (SELF ? 1 : 2); ++op) {
      const bool is_b = !SELF && op == 1;
      const uint32_t pi = SELF ? (uint32_t)it : 2u * it + op;
      const int sl = pi % cfg::NSLOTS;
      mbar_wait(&stg_full[sl], (pi / cfg::NSLOTS) & 1);
#pragma unroll
      for (int sub = 0; sub < cfg::NSUB; ++sub) {
        if (sub * KR >= st.nrows) break;               // warp-uniform: nothing to convert, the MMA issuer skips the slice too
        uint32_t hi[2][QPR][2], lo[2][QPR][2];
        if (!(p.dbg_flags & 8)) {
          const uint32_t src = stg_u32 + sl * cfg::SLOT + sub * (KR * C * 4) + ld_off;
          float4 v[2][QPR];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int j = 0; j < QPR; ++j) v[rr][j] = lds128(src + rr * (C * 4) + j * 512);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const bool ok = (sub * KR + r0 + rr) < st.nrows;      // rows past the end of the step: stale staging, contribute zero
#pragma unroll
            for (int j = 0; j < QPR; ++j) {
              float4 a = v[rr][j];
              if (!is_b) { a.x -= mean[j].x; a.y -= mean[j].y; a.z -= mean[j].z; a.w -= mean[j].w; }
              if (!ok) a = make_float4(0.f, 0.f, 0.f, 0.f);
              if (want && (SELF || is_b)) { sum[j][0] += a.x; sum[j][1] += a.y; sum[j][2] += a.z; sum[j][3] += a.w; }
              split2(a.x, a.y, hi[rr][j][0], lo[rr][j][0]);
              split2(a.z, a.w, hi[rr][j][1], lo[rr][j][1]);
            }
          }
          const uint32_t dst = ops_u32 + s * cfg::OPSTAGE + sub * cfg::SUBSTAGE + op * 2 * cfg::IMG;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int j = 0; j < QPR; ++j) {
              sts64(dst + st_off[rr] + j * 2 * (KR * 128), hi[rr][j][0], hi[rr][j][1]);
              sts64(dst + cfg::IMG + st_off[rr] + j * 2 * (KR * 128), lo[rr][j][0], lo[rr][j][1]);
            }
        }
      }
      // the slab has been consumed: hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&stg_empty[sl]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&op_full[s]);
    }
    if (t == 0) wc_mark(p.dbg, 1);
    flush(cur_slot, nflush & 1);
    if (lane == 0) bulk_wait_group<0>();     // the reduce-adds have been performed, not just read out
    if (t == 0) wc_mark(p.dbg, 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCONV_WARPS + 1) tmem_dealloc(td, cfg::TMEM_COLS);
}

template <int C, bool SELF>
int launch_tc_gram(TcGramParams& p, cudaStream_t st) {
  using cfg = GCfg<C, SELF>;
  p.ups = wc_cdiv(p.HW, UNIT_ROWS);
  p.dbg = wc_debug_buf(); p.dbg_flags = wc_debug_flags();
  long long units = (long long)p.N * p.ups;
  long long ctas = units < WC_NUM_SMS ? units : WC_NUM_SMS;
  p.units_per_cta = (units + ctas - 1) / ctas;
  ctas = (units + p.units_per_cta - 1) / p.units_per_cta;
  static bool attr_set = false;
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(tc_gram_kernel<C, SELF>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM));
    attr_set = true;
  }
  // the accumulators as [slot][C][C], 32 x 32 boxes
  alignas(64) CUtensorMap tm_out;
  const long long sstride = p.out_slot_stride ? p.out_slot_stride : (long long)C * C;
  int rc = wc_make_f32_map3(&tm_out, p.out, C, C, (uint64_t)(p.nslots > 0 ? p.nslots : 1), (uint64_t)C * 4, (uint64_t)sstride * 4, 32, 32);
  if (rc) return rc;
  WC_CUDA(wc_launch(tc_gram_kernel<C, SELF>, (unsigned)ctas, NTHREADS, cfg::SMEM, st, p, tm_out));
  WC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool wc_tc_gram_ok(int Cin, int Cout) { return Cin == Cout && (Cin == 128 || Cin == 256); }

// Sp (C,C) and sums (C) must be zero on entry (the caller memsets them on the same stream).  pivot_out[C] receives
// the shift the kernel used (mean of the first 32 rows).
int wc_tc_gram_self(const float* x, long long M, int C, int g, float* pivot_out, float* Sp, float* sums, int late_wait, cudaStream_t st) {
  TcGramParams p{};
  p.A = x; p.B = x; p.ma = nullptr; p.mb = nullptr; p.self = 1; p.g = g;
  p.N = 1; p.HW = (int)M;
  p.order = nullptr; p.slot_of = nullptr;
  p.out = Sp; p.out_slot_stride = 0;
  p.asum = sums; p.bsum = nullptr; p.sum_slot_stride = 0;
  p.pivot_out = pivot_out; p.nslots = 1; p.late_wait = late_wait;
  p.l2_load = (wc_l2_mode() & 1) ? 1 : 0;      // x is read again by K3 right after the factorisation: keep it in L2
  return C == 256 ? launch_tc_gram<256, true>(p, st) : launch_tc_gram<128, true>(p, st);
}

// P (K,C,C) and s (K,C) must be zero on entry.
int wc_tc_gram_cross(const float* x, const float* dy, int N, int HW, int C, const float* mean, const int32_t* order,
                     const int32_t* slot_of, int K, float* P, float* s, int late_wait, cudaStream_t st) {
  TcGramParams p{};
  p.A = x; p.B = dy; p.ma = mean; p.mb = nullptr; p.self = 0; p.g = C;
  p.N = N; p.HW = HW;
  p.order = order; p.slot_of = slot_of;
  p.out = P; p.out_slot_stride = (long long)C * C;
  p.asum = nullptr; p.bsum = s; p.sum_slot_stride = C;
  p.pivot_out = nullptr; p.nslots = K; p.late_wait = late_wait;
  p.l2_load = (wc_l2_mode() & 8) ? 1 : 0;      // x and dy are read again by K6
  return C == 256 ? launch_tc_gram<256, false>(p, st) : launch_tc_gram<128, false>(p, st);
}
