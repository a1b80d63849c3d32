// K1 / K4 on the tensor cores (tcgen05.mma, MN-major operands, accumulator in TMEM) for C in {128, 256}:
//
//   Out[slot][i][j] += sum_rows (A[r,i] - ma[i]) * (B[r,j] - mb[j])          (full C x C per slot)
//   asum[slot][i]   += sum_rows (A[r,i] - ma[i]),   bsum[slot][j] += sum_rows (B[r,j] - mb[j])   (optional)
//
// K1 (batch covariance):  A = B = x, shift = pivot  ("self" mode: one converted tile feeds both operands).
// K4 (backward reduce):   A = x - mean, B = dy, one accumulator per class; samples are visited sorted by class.
//
// Each fp32 operand is split into bf16 hi + lo and three products (hi.hi, hi.lo, lo.hi) are accumulated in
// fp32 in TMEM.  The rounding of the split is unbiased, so its effect on a sum over M rows shrinks like
// 1/sqrt(M); the dropped lo.lo term biases the diagonal by ~1e-6 relative ("Accuracy").
//
// One persistent CTA per SM (320 threads) owns a contiguous range of row units:
//   warps 0-7  converters: 128-bit coalesced loads -> shift -> bf16 hi/lo -> MN-major SWIZZLE_128B tiles,
//              column sums on the fly; at a slot change they drain the accumulator (tcgen05.ld -> red.global.add)
//   warp  8    MMA issuer (one thread), owns TMEM
#include "wc_common.cuh"
#include "wc_tc.cuh"

using namespace wctc;

namespace {

constexpr int KR = 32;                 // rows (K of the MMA) per stage
constexpr int NCONV_WARPS = 8;
constexpr int NCONV = NCONV_WARPS * 32;
constexpr int NTHREADS = (NCONV_WARPS + 1) * 32;
constexpr int UNIT_ROWS = 64;          // a unit never crosses a sample boundary

struct TcGramParams {
  const float* A; const float* B;       // [rows][C]
  const float* ma; const float* mb;     // shifts (nullable)
  int self;                             // 1: B == A (same shift): convert once; only blocks j <= i inside a group are flushed
  int g;                                // self mode: group size (flush only what the finalize kernel reads)
  int N, HW, ups;                       // samples, rows per sample, units per sample
  long long units_per_cta;
  const int32_t* order; const int32_t* slot_of;
  float* out; long long out_slot_stride;          // [slot][C][C]
  float* asum; float* bsum; long long sum_slot_stride;
};

template <int C>
struct GCfg {
  static constexpr int IMG = KR * C * 2;                   // one bf16 split of one operand: 16 KB (C=256)
  static constexpr int MH = C / 128;                       // M halves
  static constexpr int TMEM_COLS = MH * C;                 // 512 (C=256) / 128
  static constexpr int STAGE_CROSS = 4 * IMG;
  static constexpr int STAGES = 3;
  static constexpr int SMEM = STAGES * STAGE_CROSS + 1024;
  static constexpr int CPR = C / 8;                        // 16-byte chunks per row
  static constexpr int CHUNKS_PER_THREAD = KR * CPR / NCONV;  // 4 (C=256) / 2
  static constexpr int ROWS_PER_PASS = NCONV / CPR;        // 8 / 16
};

struct Step { int sample, row0, nrows, slot; bool flush_before; bool valid; };

struct StepIter {
  const TcGramParams& p;
  long long unit, unit_end;
  int r_in_unit, cur_slot;
  __device__ StepIter(const TcGramParams& p_, long long ub, long long ue) : p(p_), unit(ub), unit_end(ue), r_in_unit(0), cur_slot(-1) {}
  __device__ Step next() {
    Step s{};
    if (unit >= unit_end) { s.valid = false; return s; }
    int sidx = (int)(unit / p.ups), sub = (int)(unit % p.ups);
    int sample = p.order ? p.order[sidx] : sidx;
    int urow0 = sub * UNIT_ROWS;
    int unrows = min(UNIT_ROWS, p.HW - urow0);
    s.sample = sample;
    s.row0 = urow0 + r_in_unit;
    s.nrows = min(KR, unrows - r_in_unit);
    s.slot = p.slot_of ? p.slot_of[sample] : 0;
    s.flush_before = (cur_slot >= 0 && s.slot != cur_slot);
    s.valid = true;
    cur_slot = s.slot;
    r_in_unit += KR;
    if (r_in_unit >= unrows) { r_in_unit = 0; ++unit; }
    return s;
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gram_kernel(TcGramParams p) {
  using cfg = GCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[cfg::STAGES], empty_bar[cfg::STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ float sred[cfg::ROWS_PER_PASS][C];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long total_units = (long long)p.N * p.ups;
  const long long ub = (long long)blockIdx.x * p.units_per_cta;
  const long long ue = min(total_units, ub + p.units_per_cta);
  if (ub >= ue) return;   // uniform per CTA

  if (t == 0) {
    for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(&full_bar[s], NCONV_WARPS); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, NCONV_WARPS);
    mbar_fence_init();
  }
  if (warp == NCONV_WARPS) { tmem_alloc(&tmem_slot, cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_slot;
  const int stage_bytes = p.self ? 2 * cfg::IMG : 4 * cfg::IMG;

  if (warp == NCONV_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, C, 1, 1);
      StepIter iter(p, ub, ue);
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
        const int s = it % cfg::STAGES;
        const uint32_t ph = (it / cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
        const uint32_t a_lo = a_hi + cfg::IMG;
        const uint32_t b_hi = p.self ? a_hi : a_hi + 2 * cfg::IMG;
        const uint32_t b_lo = b_hi + cfg::IMG;
#pragma unroll
        for (int kk = 0; kk < KR / 16; ++kk) {
          const uint64_t dbh = make_desc(b_hi + kk * 2048, KR * 128, 1024, LAYOUT_SW128);
          const uint64_t dbl = make_desc(b_lo + kk * 2048, KR * 128, 1024, LAYOUT_SW128);
#pragma unroll
          for (int mh = 0; mh < cfg::MH; ++mh) {
            const uint32_t aoff = mh * 2 * (KR * 128) + kk * 2048;   // M half = two 64-channel blocks
            const uint64_t dah = make_desc(a_hi + aoff, KR * 128, 1024, LAYOUT_SW128);
            const uint64_t dal = make_desc(a_lo + aoff, KR * 128, 1024, LAYOUT_SW128);
            const uint32_t tdm = td + mh * C;
            umma_bf16(tdm, dah, dbh, idesc, acc);
            umma_bf16(tdm, dah, dbl, idesc, 1);
            umma_bf16(tdm, dal, dbh, idesc, 1);
          }
          acc = 1;
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_full);
    }
  } else {
    // ------------------------------------------------------------------ converters
    const int cc = t % cfg::CPR;              // this thread's 8-channel chunk (fixed)
    const int rsub = t / cfg::CPR;            // row within a pass
    const bool want_a = p.asum != nullptr, want_b = p.bsum != nullptr && !p.self;
    float sa[8], sb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) sa[q] = sb[q] = 0.f;
    float4 ma0 = make_float4(0, 0, 0, 0), ma1 = ma0, mb0 = ma0, mb1 = ma0;
    if (p.ma) { ma0 = *reinterpret_cast<const float4*>(p.ma + cc * 8); ma1 = *reinterpret_cast<const float4*>(p.ma + cc * 8 + 4); }
    if (p.mb) { mb0 = *reinterpret_cast<const float4*>(p.mb + cc * 8); mb1 = *reinterpret_cast<const float4*>(p.mb + cc * 8 + 4); }

    // one "phase" = one operand of one step; loads are prefetched one phase ahead
    // two register buffers: global loads run two phases ahead of the conversion (~64 KB in flight per SM)
    float4 v0[cfg::CHUNKS_PER_THREAD][2], v1[cfg::CHUNKS_PER_THREAD][2];
    auto gload = [&](const Step& st, int operand, float4 (&v)[cfg::CHUNKS_PER_THREAD][2]) {
      const float* src = operand == 0 ? p.A : p.B;
#pragma unroll
      for (int h = 0; h < cfg::CHUNKS_PER_THREAD; ++h) {
        const int r = rsub + h * cfg::ROWS_PER_PASS;
        if (r < st.nrows) {
          const float4* g = reinterpret_cast<const float4*>(src + ((long long)st.sample * p.HW + st.row0 + r) * C + cc * 8);
          v[h][0] = wc_ldg_stream(g);
          v[h][1] = wc_ldg_stream(g + 1);
        } else {
          v[h][0] = v[h][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    uint32_t hi[cfg::CHUNKS_PER_THREAD][4], lo[cfg::CHUNKS_PER_THREAD][4];
    auto convert = [&](const Step& st, int operand, float4 (&v)[cfg::CHUNKS_PER_THREAD][2]) {   // v -> bf16 hi/lo words (+ column sums)
      const float4 m0 = operand == 0 ? ma0 : mb0, m1 = operand == 0 ? ma1 : mb1;
      float* sum = operand == 0 ? sa : sb;
      const bool want = operand == 0 ? want_a : want_b;
#pragma unroll
      for (int h = 0; h < cfg::CHUNKS_PER_THREAD; ++h) {
        const int r = rsub + h * cfg::ROWS_PER_PASS;
        float4 a = v[h][0], b = v[h][1];
        if (r < st.nrows) { a.x -= m0.x; a.y -= m0.y; a.z -= m0.z; a.w -= m0.w; b.x -= m1.x; b.y -= m1.y; b.z -= m1.z; b.w -= m1.w; }
        if (want) { sum[0] += a.x; sum[1] += a.y; sum[2] += a.z; sum[3] += a.w; sum[4] += b.x; sum[5] += b.y; sum[6] += b.z; sum[7] += b.w; }
        split2(a.x, a.y, hi[h][0], lo[h][0]); split2(a.z, a.w, hi[h][1], lo[h][1]);
        split2(b.x, b.y, hi[h][2], lo[h][2]); split2(b.z, b.w, hi[h][3], lo[h][3]);
      }
    };
    auto store = [&](uint8_t* img_hi) {
#pragma unroll
      for (int h = 0; h < cfg::CHUNKS_PER_THREAD; ++h) {
        const int r = rsub + h * cfg::ROWS_PER_PASS;
        const uint32_t off = mnmajor_chunk_off<KR>(cc >> 3, r, cc & 7);
        *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(hi[h][0], hi[h][1], hi[h][2], hi[h][3]);
        *reinterpret_cast<uint4*>(img_hi + cfg::IMG + off) = make_uint4(lo[h][0], lo[h][1], lo[h][2], lo[h][3]);
      }
    };
    auto flush = [&](int slot, uint32_t parity) {
      // accumulator -> global (vector reductions); column sums -> global
      mbar_wait(&acc_full, parity);
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      float* out = p.out + (long long)slot * p.out_slot_stride;
#pragma unroll 1
      for (int mh = 0; mh < cfg::MH; ++mh) {
        const int i = mh * 128 + q * 32 + lane;
        const int i0 = mh * 128 + q * 32;
        int clo = 0, chi = C;                      // self mode: only blocks the finalize kernel reads (j <= i, same group)
        if (p.self) { clo = (i0 / p.g) * p.g; chi = min(((i0 + 31) / p.g + 1) * p.g, i0 + 32); }
#pragma unroll 1
        for (int c0 = half * (C / 2); c0 < (half + 1) * (C / 2); c0 += 32) {
          if (c0 >= chi || c0 + 32 <= clo) continue;   // warp-uniform
          float acc[32];
          tmem_ld32(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * C + c0), acc);
          float* dst = out + (long long)i * C + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      // column sums: reduce the ROWS_PER_PASS threads that share a channel chunk
      for (int pass = 0; pass < 2; ++pass) {
        const bool want = pass == 0 ? want_a : want_b;
        if (!want) continue;
        float* sum = pass == 0 ? sa : sb;
        asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
#pragma unroll
        for (int e = 0; e < 8; ++e) { sred[rsub][cc * 8 + e] = sum[e]; sum[e] = 0.f; }
        asm volatile("bar.sync 1, %0;" ::"n"(NCONV));
        for (int ch = t; ch < C; ch += NCONV) {
          float s = 0.f;
#pragma unroll
          for (int rr = 0; rr < cfg::ROWS_PER_PASS; ++rr) s += sred[rr][ch];
          atomicAdd((pass == 0 ? p.asum : p.bsum) + (long long)slot * p.sum_slot_stride + ch, s);
        }
      }
    };

    struct Phase { Step st; int op; int it; bool valid; };
    StepIter iter(p, ub, ue);
    const int nops = p.self ? 1 : 2;
    Step gen_step = iter.next();
    int gen_op = 0, gen_it = 0;
    auto gen = [&](Phase& ph) {
      ph.valid = gen_step.valid;
      if (!ph.valid) return;
      ph.st = gen_step; ph.op = gen_op; ph.it = gen_it;
      if (++gen_op == nops) { gen_op = 0; gen_step = iter.next(); ++gen_it; }
    };
    uint32_t nflush = 0;
    int cur_slot = gen_step.slot;
    auto process = [&](Phase& ph, float4 (&v)[cfg::CHUNKS_PER_THREAD][2]) {
      const int s = ph.it % cfg::STAGES;
      if (ph.op == 0) {
        if (ph.st.flush_before) { flush(cur_slot, nflush & 1); ++nflush; }
        cur_slot = ph.st.slot;
        mbar_wait(&empty_bar[s], ((ph.it / cfg::STAGES) & 1) ^ 1);
      }
      convert(ph.st, ph.op, v);
      const int op = ph.op;
      gen(ph);                                              // this buffer now belongs to the phase after next
      if (ph.valid) gload(ph.st, ph.op, v);
      store(smem + s * stage_bytes + op * 2 * cfg::IMG);
      if (op == nops - 1) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      }
    };
    Phase ph0, ph1;
    gen(ph0);
    if (ph0.valid) gload(ph0.st, ph0.op, v0);
    gen(ph1);
    if (ph1.valid) gload(ph1.st, ph1.op, v1);
    while (ph0.valid) {
      process(ph0, v0);
      if (!ph1.valid) break;
      process(ph1, v1);
    }
    flush(cur_slot, nflush & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCONV_WARPS) tmem_dealloc(td, cfg::TMEM_COLS);
}

template <int C>
int launch_tc_gram(TcGramParams& p, cudaStream_t st) {
  p.ups = wc_cdiv(p.HW, UNIT_ROWS);
  long long units = (long long)p.N * p.ups;
  long long ctas = units < WC_NUM_SMS ? units : WC_NUM_SMS;
  p.units_per_cta = (units + ctas - 1) / ctas;
  ctas = (units + p.units_per_cta - 1) / p.units_per_cta;
  static bool attr_set = false;
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(tc_gram_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<C>::SMEM));
    attr_set = true;
  }
  tc_gram_kernel<C><<<(unsigned)ctas, NTHREADS, GCfg<C>::SMEM, st>>>(p);
  WC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool wc_tc_gram_ok(int Cin, int Cout) { return Cin == Cout && (Cin == 128 || Cin == 256); }

// Sp (C,C) and sums (C) must be zero on entry (the caller memsets them on the same stream).
int wc_tc_gram_self(const float* x, long long M, int C, int g, const float* pivot, float* Sp, float* sums, cudaStream_t st) {
  TcGramParams p{};
  p.A = x; p.B = x; p.ma = pivot; p.mb = pivot; p.self = 1; p.g = g;
  p.N = 1; p.HW = (int)M;
  p.order = nullptr; p.slot_of = nullptr;
  p.out = Sp; p.out_slot_stride = 0;
  p.asum = sums; p.bsum = nullptr; p.sum_slot_stride = 0;
  return C == 256 ? launch_tc_gram<256>(p, st) : launch_tc_gram<128>(p, st);
}

// P (K,C,C) and s (K,C) must be zero on entry.
int wc_tc_gram_cross(const float* x, const float* dy, int N, int HW, int C, const float* mean, const int32_t* order,
                     const int32_t* slot_of, float* P, float* s, cudaStream_t st) {
  TcGramParams p{};
  p.A = x; p.B = dy; p.ma = mean; p.mb = nullptr; p.self = 0; p.g = C;
  p.N = N; p.HW = HW;
  p.order = order; p.slot_of = slot_of;
  p.out = P; p.out_slot_stride = (long long)C * C;
  p.asum = nullptr; p.bsum = s; p.sum_slot_stride = C;
  return C == 256 ? launch_tc_gram<256>(p, st) : launch_tc_gram<128>(p, st);
}
