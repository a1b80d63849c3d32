// K3 / K6 on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM) for C in {128, 256}.
//
//   out[n][r, :] = sum_seg (A_seg[n][r, :] - mean_seg) . B_seg[slot(n)]  + sign * bias[slot(n)]
//
// fp32 accuracy on bf16 tensor cores: every fp32 operand is split x = hi + lo (two bf16, 16 mantissa
// bits) and the product is accumulated as hi.hi + hi.lo + lo.hi in the fp32 TMEM accumulator ("bf16x3";
// the dropped lo.lo term is 2^-18 relative).
//
// Persistent kernel, one CTA per SM, 128-row x C output tiles handed out round-robin.  Warp roles (448 threads):
//   warp  0     producer (one thread): tensor-map TMA loads of fp32 activation boxes [128 rows x 32 channels]
//               (SWIZZLE_128B, out-of-range rows zero-filled) into a 4-deep staging ring, and 1-D bulk copies of
//               the pre-packed bf16 hi/lo images of the C x C matrices (written by wc_pack_*) into the B ring
//   warp  1     MMA issuer (one thread), owns TMEM: two 128 x C fp32 accumulators, so a tile's epilogue overlaps
//               the next tile's main loop
//   warps 2-5   epilogue: tcgen05.ld -> + bias -> swizzled shared tile -> TMA store (clipped at the sample end)
//   warps 6-13  converters: ld.shared fp32 -> centre -> bf16 hi/lo -> K-major SWIZZLE_64B operand tiles
// Every hand-over is an mbarrier pair (full/empty); smem is freed for the producer by tcgen05.commit.
#include "wc_common.cuh"
#include "wc_tc.cuh"
#include <cuda.h>

using namespace wctc;

namespace {

constexpr int KB = 32;                 // K elements per stage (64-byte bf16 rows, SWIZZLE_64B)
constexpr int TILE_M = 128;
constexpr int NCONV_WARPS = 8, NEPI_WARPS = 4;
constexpr int W_PROD = 0, W_MMA = 1, W_EPI0 = 2, W_CONV0 = W_EPI0 + NEPI_WARPS, W_BPROD = W_CONV0 + NCONV_WARPS;
constexpr int NTHREADS = (W_BPROD + 1) * 32;
constexpr int NEPI = NEPI_WARPS * 32;
constexpr int STG_BYTES = TILE_M * KB * 4;   // one fp32 activation box: 16 KB
constexpr int NEPIBUF = 2;
// ring depths <activation staging, A operand, B operand>: Rings1 is the default, the others are kept for experiments
// (wc_debug_set bits 32768 / 4096)
struct Rings0 { static constexpr int NSTG = 5, NAOP = 2, NBOP = 2; };
struct Rings1 { static constexpr int NSTG = 4, NAOP = 2, NBOP = 3; };
struct Rings2 { static constexpr int NSTG = 6, NAOP = 2, NBOP = 2; };

struct TcSeg {
  const float* mean;     // per-channel shift (nullable)
  const uint8_t* Bp;     // packed operand images: [slot][kb][hi|lo][C x 64 B]
  long long slot_stride; // bytes between slots (0 = shared)
};

struct TcApplyParams {
  TcSeg seg[2];
  int nseg;
  const float* bias; long long bias_slot_stride; float bias_sign;
  int N, HW, tiles_per_sample;
  int tile_rows;                            // rows of the activation a tile covers (<= TILE_M; the MMA always runs M = TILE_M)
  long long total_tiles;
  const int32_t* cls; int K; int32_t* info;
  int relu;                                 // 1: out = max(0, .) in the epilogue
  // whole-layer calls only (the caller knows what the preceding kernel writes).  bit 0: activations, shifts and labels were
  // final before the preceding kernel started, so the activation producer and the converters do not wait for it (only
  // the matrix-image producer and the epilogue do: images and bias are what it writes); bit 1: the images of segment 0
  // are older too (K6: T was packed in the forward pass; only A and c come from the backward chain)
  int early;
  int l2_load[2], l2_store;                 // L2 eviction priority of the activation loads per segment / of the output stores (wctc::l2_policy)
  unsigned long long* dbg; int dbg_flags;   // debug marks / experiment switches (0 in normal use)
};

template <int C, class R>
struct Cfg {
  static constexpr int NSTG = R::NSTG, NAOP = R::NAOP, NBOP = R::NBOP;
  static constexpr int N = C;
  static constexpr int NKB = C / KB;                       // K blocks per segment
  static constexpr int A_IMG = TILE_M * KB * 2;            // 8 KB: one bf16 split of an A tile
  static constexpr int B_IMG = N * KB * 2;                 // 16 KB (C=256) / 8 KB (C=128)
  static constexpr int OFF_AOP = NSTG * STG_BYTES;
  static constexpr int OFF_BOP = OFF_AOP + NAOP * 2 * A_IMG;
  static constexpr int OFF_EPI = OFF_BOP + NBOP * 2 * B_IMG;
  static constexpr int SMEM = OFF_EPI + NEPIBUF * STG_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 2 * N;                  // 512 / 256
};

struct TileInfo { int sample, r0, slot; };

__device__ __forceinline__ TileInfo tile_info(const TcApplyParams& p, long long tile) {
  TileInfo ti;
  // tiles are visited from the end of the tensor: the rows the previous kernel read last are the ones most likely
  // still resident in L2
  const long long tr = p.total_tiles - 1 - tile;
  ti.sample = (int)(tr / p.tiles_per_sample);
  ti.r0 = (int)(tr % p.tiles_per_sample) * p.tile_rows;
  ti.slot = 0;
  if (p.cls) {
    int c = p.cls[ti.sample];
    if (c < 0 || c >= p.K) {
      if (p.info && threadIdx.x == 0 && ti.r0 == 0) atomicMax(&p.info[WC_INFO_BAD_CLASS], ti.sample + 1);
      c = c < 0 ? 0 : p.K - 1;
    }
    ti.slot = c;
  }
  return ti;
}

template <int C, class R>
__global__ void __launch_bounds__(NTHREADS, 1) tc_apply_kernel(TcApplyParams p, const __grid_constant__ CUtensorMap tm_a0,
                                                               const __grid_constant__ CUtensorMap tm_a1,
                                                               const __grid_constant__ CUtensorMap tm_out) {
  using cfg = Cfg<C, R>;
  constexpr int NSTG = cfg::NSTG, NAOP = cfg::NAOP, NBOP = cfg::NBOP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t stg_full[NSTG], stg_empty[NSTG], aop_full[NAOP], aop_empty[NAOP], bop_full[NBOP], bop_empty[NBOP],
      acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int total_kb = p.nseg * cfg::NKB;

  if (t == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], NCONV_WARPS); }
    for (int s = 0; s < NAOP; ++s) { mbar_init(&aop_full[s], NCONV_WARPS); mbar_init(&aop_empty[s], 1); }
    for (int s = 0; s < NBOP; ++s) { mbar_init(&bop_full[s], 1); mbar_init(&bop_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], NEPI_WARPS); }
    mbar_fence_init();
    tma_prefetch_desc(&tm_a0); tma_prefetch_desc(&tm_a1); tma_prefetch_desc(&tm_out);
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_slot, cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_slot;
  const bool early = (p.early & 1) != 0;
  if (!early) wc_pdl_wait();      // barriers, tensor maps and TMEM were set up while the previous kernel drained
  if (t == 0) wc_mark(p.dbg, 0);

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ activation producer
    // (its own thread: the staging ring runs NSTG boxes ahead of the converters, independent of the B ring)
    if (lane == 0) {
      uint32_t ia = 0;
      const uint64_t pol[2] = {l2_policy(p.l2_load[0]), l2_policy(p.l2_load[1])};
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileInfo ti = tile_info(p, tile);
        for (int it = 0; it < total_kb; ++it, ++ia) {
          const int sg = it / cfg::NKB, kb = it % cfg::NKB;
          const int s = ia % NSTG;
          mbar_wait(&stg_empty[s], ((ia / NSTG) & 1) ^ 1);
          if (p.dbg_flags & 1024) mbar_arrive(&stg_full[s]);
          else {
            mbar_arrive_expect_tx(&stg_full[s], (uint32_t)p.tile_rows * KB * 4);
            tma_load_3d_hint(smem + s * STG_BYTES, sg == 0 ? &tm_a0 : &tm_a1, kb * KB, ti.r0, ti.sample, &stg_full[s], pol[sg]);
          }
        }
      }
    }
  } else if (warp == W_BPROD) {
    // ------------------------------------------------------------------ matrix-image producer (L2-resident source)
    if (lane == 0) {
      uint32_t ib = 0;
      bool waited = !early;
      const int wait_seg = (p.early & 2) ? 1 : 0;      // first segment whose images the preceding kernel writes
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileInfo ti = tile_info(p, tile);
        for (int it = 0; it < total_kb; ++it, ++ib) {
          const int sg = it / cfg::NKB, kb = it % cfg::NKB;
          const int s = ib % NBOP;
          if (!waited && sg >= wait_seg) { wc_pdl_wait(); waited = true; }
          const uint8_t* src = p.seg[sg].Bp + (long long)ti.slot * p.seg[sg].slot_stride + (long long)kb * (2 * cfg::B_IMG);
          mbar_wait(&bop_empty[s], ((ib / NBOP) & 1) ^ 1);
          uint8_t* dst = smem + cfg::OFF_BOP + s * 2 * cfg::B_IMG;
          if ((p.dbg_flags & 64) && ib >= NBOP) mbar_arrive(&bop_full[s]);
          else {
            mbar_arrive_expect_tx(&bop_full[s], 2 * cfg::B_IMG);
            bulk_g2s(dst, src, cfg::B_IMG, &bop_full[s]);
            bulk_g2s(dst + cfg::B_IMG, src + cfg::B_IMG, cfg::B_IMG, &bop_full[s]);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TILE_M, cfg::N, 0, 0);
      uint32_t ia = 0, ib = 0, lt = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const uint32_t ab = lt & 1;
        mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = td + ab * cfg::N;
        uint32_t acc = 0;
        for (int it = 0; it < total_kb; ++it, ++ia, ++ib) {
          const int sa = ia % NAOP, sb = ib % NBOP;
          mbar_wait(&aop_full[sa], (ia / NAOP) & 1);
          mbar_wait(&bop_full[sb], (ib / NBOP) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + cfg::OFF_AOP + sa * 2 * cfg::A_IMG);
          const uint32_t a_lo = a_hi + cfg::A_IMG;
          const uint32_t b_hi = smem_u32(smem + cfg::OFF_BOP + sb * 2 * cfg::B_IMG);
          const uint32_t b_lo = b_hi + cfg::B_IMG;
#pragma unroll
          for (int kk = 0; kk < KB / 16; ++kk) {
            const uint64_t dah = make_desc(a_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dal = make_desc(a_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dbh = make_desc(b_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dbl = make_desc(b_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            if (!(p.dbg_flags & 256)) {
              umma_bf16(tacc, dah, dbh, idesc, acc);
              umma_bf16(tacc, dah, dbl, idesc, 1);
              umma_bf16(tacc, dal, dbh, idesc, 1);
            }
            acc = 1;
          }
          umma_commit(&aop_empty[sa]);     // both operand stages are reusable once these MMAs have read them
          umma_commit(&bop_empty[sb]);
        }
        umma_commit(&acc_full[ab]);        // accumulator complete
      }
    }
  } else if (warp >= W_EPI0 && warp < W_CONV0) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int te = t - W_EPI0 * 32;
    const int q = warp & 3;                // TMEM lanes 32q .. 32q+31 are the ones this warp may read
    const int row = q * 32 + lane;         // row of the tile held by this thread
    uint32_t lt = 0, ec = 0;
    const uint64_t pol_st = l2_policy(p.l2_store);
    if (early) wc_pdl_wait();              // the bias is the preceding kernel's; nothing to do here before the first tile anyway
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const TileInfo ti = tile_info(p, tile);
      const uint32_t ab = lt & 1;
      const float* bias = p.bias ? p.bias + (long long)ti.slot * p.bias_slot_stride : nullptr;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < cfg::N; c0 += 32, ++ec) {
        const uint32_t eb = ec & 1;
        // the TMA store issued two chunks ago has finished reading this buffer
        if (te == 0) bulk_wait_group_read<1>();
        asm volatile("bar.sync 2, %0;" ::"n"(NEPI));
        float acc[32];
        tmem_ld32(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * cfg::N + c0), acc);
        if (c0 + 32 == cfg::N) {           // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        const uint32_t ebase = smem_u32(smem + cfg::OFF_EPI + eb * STG_BYTES) + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
          const float s = p.bias_sign;
          float o0 = fmaf(s, bv.x, acc[4 * j]), o1 = fmaf(s, bv.y, acc[4 * j + 1]), o2 = fmaf(s, bv.z, acc[4 * j + 2]), o3 = fmaf(s, bv.w, acc[4 * j + 3]);
          if (p.relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f); }
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ebase + ((j ^ (row & 7)) << 4)), "f"(o0), "f"(o1), "f"(o2), "f"(o3) : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 2, %0;" ::"n"(NEPI));
        if (te == 0 && !(p.dbg_flags & 512)) {
          tma_store_3d_hint(&tm_out, c0, ti.r0, ti.sample, smem + cfg::OFF_EPI + eb * STG_BYTES, pol_st);
          bulk_commit_group();
        }
      }
    }
    if (te == 0) bulk_wait_group<0>();
  } else {
    // ------------------------------------------------------------------ A converters (256 threads)
    // staged box: row r = 128 contiguous bytes, its 16-byte chunk j sits at position j ^ (r & 7) (SWIZZLE_128B).
    // A thread converts 8 channels of two rows (r, r + 64): 2 ld.shared.v4 + 2 st.shared.v4 per row.
    const int tcv = t - W_CONV0 * 32;
    const int crow0 = tcv >> 2, cch = tcv & 3;
    const uint32_t stg_u32 = smem_u32(smem), aop_u32 = smem_u32(smem + cfg::OFF_AOP);
    uint32_t ld_off[2][2], st_off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = crow0 + h * 64;
      ld_off[h][0] = r * 128 + (((2 * cch) ^ (r & 7)) << 4);
      ld_off[h][1] = r * 128 + (((2 * cch + 1) ^ (r & 7)) << 4);
      st_off[h] = kmajor_chunk_off<KB>(r, cch);
    }
    uint32_t ia = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int it = 0; it < total_kb; ++it, ++ia) {
        const int sg = it / cfg::NKB, kb = it % cfg::NKB;
        const int s = ia % NSTG, sa = ia % NAOP;
        float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;       // per-channel shift: 64 bytes per stage, L1-resident
        if (const float* mp = p.seg[sg].mean) {
          m0 = __ldg(reinterpret_cast<const float4*>(mp + kb * KB + cch * 8));
          m1 = __ldg(reinterpret_cast<const float4*>(mp + kb * KB + cch * 8 + 4));
        }
        mbar_wait(&stg_full[s], (ia / NSTG) & 1);
        uint32_t hi[2][4], lo[2][4];
        const bool skipc = p.dbg_flags & 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (skipc) continue;
          float4 a = lds128(stg_u32 + s * STG_BYTES + ld_off[h][0]);
          float4 b = lds128(stg_u32 + s * STG_BYTES + ld_off[h][1]);
          a.x -= m0.x; a.y -= m0.y; a.z -= m0.z; a.w -= m0.w; b.x -= m1.x; b.y -= m1.y; b.z -= m1.z; b.w -= m1.w;
          split2(a.x, a.y, hi[h][0], lo[h][0]);
          split2(a.z, a.w, hi[h][1], lo[h][1]);
          split2(b.x, b.y, hi[h][2], lo[h][2]);
          split2(b.z, b.w, hi[h][3], lo[h][3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_empty[s]);            // the box is in registers: give it back to the producer
        mbar_wait(&aop_empty[sa], ((ia / NAOP) & 1) ^ 1);
        const uint32_t dst = aop_u32 + sa * 2 * cfg::A_IMG;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (skipc) continue;
          sts128(dst + st_off[h], hi[h][0], hi[h][1], hi[h][2], hi[h][3]);
          sts128(dst + cfg::A_IMG + st_off[h], lo[h][0], lo[h][1], lo[h][2], lo[h][3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aop_full[sa]);
      }
    }
  }
  if (t == 0) wc_mark(p.dbg, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(td, cfg::TMEM_COLS);
}

// =====================================================================================================================
// CTA-pair version (cta_group::2).  NOT the default: measured (profiles/r02_pair_apply.txt) it is 3 us faster than the
// single-CTA kernel on warm inputs (K3 24.4 vs 27.4 us, issuer 810 of 1190 cycles per k-block busy issuing MMAs) but no
// faster inside the cold-cache step (K3 33.0 vs 32.3, K6 52.7 vs 50.3 us per graph timeline), where both wait for HBM and
// for the 256-row wave quantisation (256 pair tiles on 74 pairs = 3.46 rounds).  Selected with wc_debug_set bit 2097152.
// =====================================================================================================================
// The single-CTA kernel above re-streams the packed C x C matrices from L2 for every 128-row tile: at C = 256 that is
// 256 KB (K3) / 512 KB (K6) of L2 reads per 128 KB of activations, ~70 % of the kernel's L2 -> SM traffic, and with it the
// kernel sits at the L2 output limit with the tensor pipe 44 / 58 % busy (profiles/r01e_ncu_full_summary.txt).  Here two
// CTAs of a cluster (one TPC) share each MMA: a pair works on 256 rows, every CTA converts its own 128 rows (A operand)
// and loads only HALF of the N rows of each matrix image (B operand); tcgen05.mma.cta_group::2 with M = 256 reads both
// halves from the same shared-memory offsets of the two CTAs.  Matrix traffic from L2 and the MMA's shared-memory reads
// of B are halved; everything on the activation side (TMA boxes, converters, epilogue, TMA stores) stays CTA-local.
//
// Hand-overs across the pair (rank 0 = leader, issues every MMA):
//   peer A / B stage ready  -> a relay thread in the peer (its otherwise idle MMA warp) waits on the peer's LOCAL full barrier
//                              and arrives remotely on the leader's aop_pfull / bop_pfull
//   stage consumed, accumulator complete -> tcgen05.commit multicast: arrives on the same barrier offset in both CTAs
//   peer accumulator drained -> the peer's epilogue warps arrive remotely on the leader's acc_pempty
template <int C>
struct Cfg2 {
  static constexpr int NSTG = 4, NAOP = 3, NBOP = 4;
  static constexpr int N = C;
  static constexpr int NKB = C / KB;
  static constexpr int A_IMG = TILE_M * KB * 2;            // 8 KB: one bf16 split of this CTA's A tile
  static constexpr int B_HALF = (N / 2) * KB * 2;          // 8 KB (C=256) / 4 KB (C=128): this CTA's half of one bf16 split of B
  static constexpr int B_IMG = N * KB * 2;                 // the whole split in the packed buffer
  static constexpr int OFF_AOP = NSTG * STG_BYTES;
  static constexpr int OFF_BOP = OFF_AOP + NAOP * 2 * A_IMG;
  static constexpr int OFF_EPI = OFF_BOP + NBOP * 2 * B_HALF;
  static constexpr int SMEM = OFF_EPI + NEPIBUF * STG_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * N;
};

__device__ __forceinline__ TileInfo tile_info2(const TcApplyParams& p, long long ptile, int rank) {
  TileInfo ti;
  const long long tr = p.total_tiles - 1 - ptile;          // pair tiles, visited from the end of the tensor
  ti.sample = (int)(tr / p.tiles_per_sample);
  ti.r0 = (int)(tr % p.tiles_per_sample) * (2 * TILE_M) + rank * TILE_M;    // rows past the sample end: zero-filled / clipped by TMA
  ti.slot = 0;
  if (p.cls) {
    int c = p.cls[ti.sample];
    if (c < 0 || c >= p.K) {
      if (p.info && threadIdx.x == 0 && ti.r0 == 0) atomicMax(&p.info[WC_INFO_BAD_CLASS], ti.sample + 1);
      c = c < 0 ? 0 : p.K - 1;
    }
    ti.slot = c;
  }
  return ti;
}

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1) tc_apply2_kernel(TcApplyParams p, const __grid_constant__ CUtensorMap tm_a0,
                                                                const __grid_constant__ CUtensorMap tm_a1,
                                                                const __grid_constant__ CUtensorMap tm_out) {
  using cfg = Cfg2<C>;
  constexpr int NSTG = cfg::NSTG, NAOP = cfg::NAOP, NBOP = cfg::NBOP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t stg_full[NSTG], stg_empty[NSTG], aop_full[NAOP], aop_pfull[NAOP], aop_empty[NAOP], bop_full[NBOP],
      bop_pfull[NBOP], bop_empty[NBOP], acc_full[2], acc_empty[2], acc_pempty[2];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = p.nseg * cfg::NKB;

  if (t == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], NCONV_WARPS); }
    for (int s = 0; s < NAOP; ++s) { mbar_init(&aop_full[s], NCONV_WARPS); mbar_init(&aop_pfull[s], 1); mbar_init(&aop_empty[s], 1); }
    for (int s = 0; s < NBOP; ++s) { mbar_init(&bop_full[s], 1); mbar_init(&bop_pfull[s], 1); mbar_init(&bop_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], NEPI_WARPS); mbar_init(&acc_pempty[s], NEPI_WARPS); }
    mbar_fence_init();
    tma_prefetch_desc(&tm_a0); tma_prefetch_desc(&tm_a1); tma_prefetch_desc(&tm_out);
  }
  __syncthreads();
  cluster_sync_all();                  // both CTAs' barriers exist before anything can arrive on them
  if (warp == W_MMA) { tmem_alloc2(&tmem_slot, cfg::TMEM_COLS); tmem_relinquish2(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_slot;
  wc_pdl_wait();
  if (t == 0) wc_mark(p.dbg, 0);

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ activation producer (CTA-local)
    if (lane == 0) {
      uint32_t ia = 0;
      const uint64_t pol[2] = {l2_policy(p.l2_load[0]), l2_policy(p.l2_load[1])};
      for (long long pt = pair; pt < p.total_tiles; pt += npairs) {
        const TileInfo ti = tile_info2(p, pt, rank);
        for (int it = 0; it < total_kb; ++it, ++ia) {
          const int sg = it / cfg::NKB, kb = it % cfg::NKB;
          const int s = ia % NSTG;
          mbar_wait(&stg_empty[s], ((ia / NSTG) & 1) ^ 1);
          mbar_arrive_expect_tx(&stg_full[s], STG_BYTES);
          tma_load_3d_hint(smem + s * STG_BYTES, sg == 0 ? &tm_a0 : &tm_a1, kb * KB, ti.r0, ti.sample, &stg_full[s], pol[sg]);
        }
      }
    }
  } else if (warp == W_BPROD) {
    // ------------------------------------------------------------------ matrix-image producer: this CTA's half of N
    if (lane == 0) {
      uint32_t ib = 0;
      for (long long pt = pair; pt < p.total_tiles; pt += npairs) {
        const TileInfo ti = tile_info2(p, pt, rank);
        for (int it = 0; it < total_kb; ++it, ++ib) {
          const int sg = it / cfg::NKB, kb = it % cfg::NKB;
          const int s = ib % NBOP;
          const uint8_t* src = p.seg[sg].Bp + (long long)ti.slot * p.seg[sg].slot_stride + (long long)kb * (2 * cfg::B_IMG) +
                               (long long)rank * cfg::B_HALF;
          mbar_wait(&bop_empty[s], ((ib / NBOP) & 1) ^ 1);
          uint8_t* dst = smem + cfg::OFF_BOP + s * 2 * cfg::B_HALF;
          mbar_arrive_expect_tx(&bop_full[s], 2 * cfg::B_HALF);
          bulk_g2s(dst, src, cfg::B_HALF, &bop_full[s]);                                   // hi
          bulk_g2s(dst + cfg::B_HALF, src + cfg::B_IMG, cfg::B_HALF, &bop_full[s]);        // lo
        }
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      const uint32_t idesc = make_idesc_bf16(2 * TILE_M, cfg::N, 0, 0);
      uint32_t ia = 0, ib = 0, lt = 0;
      const bool prof = p.dbg != nullptr;
      long long w_loc = 0, w_pa = 0, w_pb = 0, w_acc = 0;
      const long long cstart = prof ? clock64() : 0;
      for (long long pt = pair; pt < p.total_tiles; pt += npairs, ++lt) {
        const uint32_t ab = lt & 1;
        const long long ca = prof ? clock64() : 0;
        mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);               // both epilogues have drained this accumulator
        mbar_wait_cluster(&acc_pempty[ab], ((lt >> 1) & 1) ^ 1);
        w_acc += (prof ? clock64() : 0) - ca;
        tc_fence_after();
        const uint32_t tacc = td + ab * cfg::N;
        uint32_t acc = 0;
        for (int it = 0; it < total_kb; ++it, ++ia, ++ib) {
          const int sa = ia % NAOP, sb = ib % NBOP;
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(&aop_full[sa], (ia / NAOP) & 1);
          mbar_wait(&bop_full[sb], (ib / NBOP) & 1);
          const long long c1 = prof ? clock64() : 0;
          mbar_wait_cluster(&aop_pfull[sa], (ia / NAOP) & 1);
          const long long c2 = prof ? clock64() : 0;
          mbar_wait_cluster(&bop_pfull[sb], (ib / NBOP) & 1);
          const long long c3 = prof ? clock64() : 0;
          w_loc += c1 - c0; w_pa += c2 - c1; w_pb += c3 - c2;
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + cfg::OFF_AOP + sa * 2 * cfg::A_IMG);
          const uint32_t a_lo = a_hi + cfg::A_IMG;
          const uint32_t b_hi = smem_u32(smem + cfg::OFF_BOP + sb * 2 * cfg::B_HALF);
          const uint32_t b_lo = b_hi + cfg::B_HALF;
#pragma unroll
          for (int kk = 0; kk < KB / 16; ++kk) {
            const uint64_t dah = make_desc(a_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dal = make_desc(a_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dbh = make_desc(b_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            const uint64_t dbl = make_desc(b_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
            if (!(p.dbg_flags & 256)) umma_bf16_2cta(tacc, dah, dbh, idesc, acc);
            if (!(p.dbg_flags & (256 | 512))) {
              umma_bf16_2cta(tacc, dah, dbl, idesc, 1);
              umma_bf16_2cta(tacc, dal, dbh, idesc, 1);
            }
            acc = 1;
          }
          umma_commit_pair(&aop_empty[sa]);      // frees the stage in BOTH CTAs
          umma_commit_pair(&bop_empty[sb]);
        }
        umma_commit_pair(&acc_full[ab]);
      }
      if (prof && blockIdx.x < WC_DBG_BLOCKS) {     // cycles the issuer spent waiting: local stages | peer A | peer B | accumulators | total
        unsigned long long* d = p.dbg + blockIdx.x * WC_DBG_SLOTS;
        d[2] = w_loc; d[3] = w_pa; d[4] = w_pb; d[5] = w_acc; d[6] = clock64() - cstart; d[7] = ia;
      }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- relay (peer CTA): local "stage full" -> leader
      uint32_t ia = 0, ib = 0;
      for (long long pt = pair; pt < p.total_tiles; pt += npairs) {
        for (int it = 0; it < total_kb; ++it, ++ia, ++ib) {
          const int sa = ia % NAOP, sb = ib % NBOP;
          mbar_wait(&aop_full[sa], (ia / NAOP) & 1);
          mbar_arrive_remote_relaxed(&aop_pfull[sa], 0);
          mbar_wait(&bop_full[sb], (ib / NBOP) & 1);
          mbar_arrive_remote_relaxed(&bop_pfull[sb], 0);
        }
      }
    }
  } else if (warp >= W_EPI0 && warp < W_CONV0) {
    // ------------------------------------------------------------------ epilogue (4 warps = this CTA's 128 TMEM lanes)
    const int te = t - W_EPI0 * 32;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t lt = 0, ec = 0;
    const uint64_t pol_st = l2_policy(p.l2_store);
    for (long long pt = pair; pt < p.total_tiles; pt += npairs, ++lt) {
      const TileInfo ti = tile_info2(p, pt, rank);
      const uint32_t ab = lt & 1;
      const float* bias = p.bias ? p.bias + (long long)ti.slot * p.bias_slot_stride : nullptr;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < cfg::N; c0 += 32, ++ec) {
        const uint32_t eb = ec & 1;
        if (te == 0) bulk_wait_group_read<1>();
        asm volatile("bar.sync 2, %0;" ::"n"(NEPI));
        float acc[32];
        tmem_ld32(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * cfg::N + c0), acc);
        if (c0 + 32 == cfg::N) {           // last read of this accumulator: hand it back to the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (rank == 0) mbar_arrive(&acc_empty[ab]); else mbar_arrive_remote(&acc_pempty[ab], 0); }
        }
        const uint32_t ebase = smem_u32(smem + cfg::OFF_EPI + eb * STG_BYTES) + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
          const float s = p.bias_sign;
          float o0 = fmaf(s, bv.x, acc[4 * j]), o1 = fmaf(s, bv.y, acc[4 * j + 1]), o2 = fmaf(s, bv.z, acc[4 * j + 2]), o3 = fmaf(s, bv.w, acc[4 * j + 3]);
          if (p.relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f); }
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ebase + ((j ^ (row & 7)) << 4)), "f"(o0), "f"(o1), "f"(o2), "f"(o3) : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 2, %0;" ::"n"(NEPI));
        if (te == 0) {
          tma_store_3d_hint(&tm_out, c0, ti.r0, ti.sample, smem + cfg::OFF_EPI + eb * STG_BYTES, pol_st);
          bulk_commit_group();
        }
      }
    }
    if (te == 0) bulk_wait_group<0>();
  } else {
    // ------------------------------------------------------------------ A converters (256 threads, CTA-local)
    const int tcv = t - W_CONV0 * 32;
    const int crow0 = tcv >> 2, cch = tcv & 3;
    const uint32_t stg_u32 = smem_u32(smem), aop_u32 = smem_u32(smem + cfg::OFF_AOP);
    uint32_t ld_off[2][2], st_off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = crow0 + h * 64;
      ld_off[h][0] = r * 128 + (((2 * cch) ^ (r & 7)) << 4);
      ld_off[h][1] = r * 128 + (((2 * cch + 1) ^ (r & 7)) << 4);
      st_off[h] = kmajor_chunk_off<KB>(r, cch);
    }
    uint32_t ia = 0;
    for (long long pt = pair; pt < p.total_tiles; pt += npairs) {
      for (int it = 0; it < total_kb; ++it, ++ia) {
        const int sg = it / cfg::NKB, kb = it % cfg::NKB;
        const int s = ia % NSTG, sa = ia % NAOP;
        float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
        if (const float* mp = p.seg[sg].mean) {
          m0 = __ldg(reinterpret_cast<const float4*>(mp + kb * KB + cch * 8));
          m1 = __ldg(reinterpret_cast<const float4*>(mp + kb * KB + cch * 8 + 4));
        }
        mbar_wait(&stg_full[s], (ia / NSTG) & 1);
        uint32_t hi[2][4], lo[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 a = lds128(stg_u32 + s * STG_BYTES + ld_off[h][0]);
          float4 b = lds128(stg_u32 + s * STG_BYTES + ld_off[h][1]);
          a.x -= m0.x; a.y -= m0.y; a.z -= m0.z; a.w -= m0.w; b.x -= m1.x; b.y -= m1.y; b.z -= m1.z; b.w -= m1.w;
          split2(a.x, a.y, hi[h][0], lo[h][0]);
          split2(a.z, a.w, hi[h][1], lo[h][1]);
          split2(b.x, b.y, hi[h][2], lo[h][2]);
          split2(b.z, b.w, hi[h][3], lo[h][3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_empty[s]);
        mbar_wait(&aop_empty[sa], ((ia / NAOP) & 1) ^ 1);
        const uint32_t dst = aop_u32 + sa * 2 * cfg::A_IMG;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          sts128(dst + st_off[h], hi[h][0], hi[h][1], hi[h][2], hi[h][3]);
          sts128(dst + cfg::A_IMG + st_off[h], lo[h][0], lo[h][1], lo[h][2], lo[h][3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aop_full[sa]);
      }
    }
  }
  if (t == 0) wc_mark(p.dbg, 1);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the leader's last MMAs have read the peer's shared memory; both epilogues are done
  if (warp == W_MMA) tmem_dealloc2(td, cfg::TMEM_COLS);
}

// ---- operand packing: fp32 matrix -> bf16 hi/lo K-major SWIZZLE_64B images -------------------------------
// image(kb)[n][k in 32] = src(n, kb*32 + k);  one thread per 16-byte chunk (8 k's of one n).
// mode 0: src(n,k) = M[k*ld + n] (K3: T^T);  mode 1: src(n,k) = M[n*ld + k] (K6 term 1: T);
// mode 2: src(n,k) = blockdiag(A (G,g,g))[k][n]  (K6 term 2).
__global__ void __launch_bounds__(256) pack_b_kernel(const float* __restrict__ M, int C, int nslots, long long slot_stride_in,
                                                     int mode, int g, uint8_t* __restrict__ out, long long slot_stride_out) {
  wc_pdl_wait();
  const int nkb = C / KB;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long per_slot = (long long)C * nkb * 4;        // chunks per slot
  if (idx >= per_slot * nslots) return;
  int slot = (int)(idx / per_slot);
  int rem = (int)(idx % per_slot);
  int n = rem % C;
  int ch = (rem / C) % 4;
  int kb = rem / (C * 4);
  const float* src = M + (long long)slot * slot_stride_in;
  float vals[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int k = kb * KB + ch * 8 + q;
    float v;
    if (mode == 0) v = src[(long long)k * C + n];
    else if (mode == 1) v = src[(long long)n * C + k];
    else {
      int gk = k / g, gn = n / g;
      v = (gk == gn) ? src[(long long)gk * g * g + (long long)(k - gk * g) * g + (n - gn * g)] : 0.f;
    }
    vals[q] = v;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split2(vals[2 * q], vals[2 * q + 1], hi[q], lo[q]);
  const int b_img = C * KB * 2;
  uint8_t* dst = out + (long long)slot * slot_stride_out + (long long)kb * (2 * b_img) + kmajor_chunk_off<KB>(n, ch);
  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + b_img) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// fp32 activations viewed as [N][HW][C]; box = 128 rows x 32 channels of one sample, SWIZZLE_128B
int make_act_map(CUtensorMap* tm, const float* base, int N, int HW, int C, int rows = TILE_M) {
  return wc_make_f32_map3(tm, base, (uint64_t)C, (uint64_t)HW, (uint64_t)N, (uint64_t)C * 4, (uint64_t)HW * C * 4, KB, (uint32_t)rows);
}

template <int C, class R>
int launch_tc_apply_r(const TcApplyParams& p, const CUtensorMap& tm_a0, const CUtensorMap& tm_a1, const CUtensorMap& tm_out,
                      cudaStream_t st) {
  static bool attr_set = false;   // idempotent; a race only repeats the same call
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(tc_apply_kernel<C, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<C, R>::SMEM));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(p.total_tiles < WC_NUM_SMS ? p.total_tiles : WC_NUM_SMS);
  WC_CUDA(wc_launch(tc_apply_kernel<C, R>, grid, NTHREADS, Cfg<C, R>::SMEM, st, p, tm_a0, tm_a1, tm_out));
  WC_CHECK_LAUNCH();
  return 0;
}

template <int C>
int launch_tc_apply2(const TcApplyParams& p, const CUtensorMap& tm_a0, const CUtensorMap& tm_a1, const CUtensorMap& tm_out,
                     cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(tc_apply2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<C>::SMEM));
    attr_set = true;
  }
  const long long pairs = p.total_tiles < WC_NUM_SMS / 2 ? p.total_tiles : WC_NUM_SMS / 2;
  WC_CUDA(wc_launch_cluster(tc_apply2_kernel<C>, (unsigned)(2 * pairs), NTHREADS, Cfg2<C>::SMEM, st, 2, p, tm_a0, tm_a1, tm_out));
  WC_CHECK_LAUNCH();
  return 0;
}

template <int C>
int launch_tc_apply(TcApplyParams& p, const float* a0, const float* a1, float* out, int N, int HW, cudaStream_t st) {
  const bool cond = p.cls != nullptr && p.K > 1;
  if (!cond) { HW = N * HW; N = 1; p.cls = nullptr; }
  p.N = N; p.HW = HW;
  p.dbg = wc_debug_buf(); p.dbg_flags = wc_debug_flags();
  if (wc_debug_flags() & 2097152) {
    p.tile_rows = TILE_M;
    // CTA pairs: a pair tile is 256 rows of one sample
    p.tiles_per_sample = wc_cdiv(HW, 2 * TILE_M);
    p.total_tiles = (long long)N * p.tiles_per_sample;
    p.dbg = wc_debug_buf(); p.dbg_flags = wc_debug_flags();
    alignas(64) CUtensorMap tm_a0, tm_a1, tm_out;
    int rc = make_act_map(&tm_a0, a0, N, HW, C);
    if (!rc) rc = make_act_map(&tm_a1, a1 ? a1 : a0, N, HW, C);
    if (!rc) rc = make_act_map(&tm_out, out, N, HW, C);
    if (rc) return rc;
    return launch_tc_apply2<C>(p, tm_a0, tm_a1, tm_out, st);
  }
  // Wave quantisation: 128-row tiles of a 65536-row tensor are 512 tiles = 3.46 rounds on 148 CTAs.  Experiment (debug bit
  // 4194304, off by default): when all rows form one run (unconditional layer), lower the tile height so that the last
  // round is full as well (4 rounds of 112-row tiles; the MMA still runs M = 128, the rows beyond tile_rows hold stale
  // operands and are never stored).  Measured: K6 54.7 us instead of 52.4 us stand-alone, whole step +0.4 us -- the CTAs
  // share HBM bandwidth rather than run in lock-step rounds (while 68 CTAs work on their fourth tile they get the
  // bandwidth the 80 finished ones no longer use), so the idle tail is much smaller than the round count suggests, and
  // shorter tiles add per-tile overhead.
  p.tile_rows = TILE_M;
  if (!cond && (p.dbg_flags & 4194304)) {
    const long long rounds = (wc_cdiv(HW, TILE_M) + WC_NUM_SMS - 1) / WC_NUM_SMS;
    const long long r = ((long long)HW + rounds * WC_NUM_SMS - 1) / (rounds * WC_NUM_SMS);
    const int r8 = (int)((r + 7) / 8 * 8);
    if (rounds > 1 && r8 >= 64 && r8 < TILE_M) p.tile_rows = r8;
  }
  p.tiles_per_sample = wc_cdiv(HW, p.tile_rows);
  p.total_tiles = (long long)N * p.tiles_per_sample;
  alignas(64) CUtensorMap tm_a0, tm_a1, tm_out;
  int rc = make_act_map(&tm_a0, a0, N, HW, C, p.tile_rows);
  if (!rc) rc = make_act_map(&tm_a1, a1 ? a1 : a0, N, HW, C, p.tile_rows);
  if (!rc) rc = make_act_map(&tm_out, out, N, HW, C, p.tile_rows);
  if (rc) return rc;
  // default: 4 activation boxes + 3 matrix-image stages in flight (measured 1 us faster than 5 + 2 on K3 and K6)
  if (p.dbg_flags & 32768) return launch_tc_apply_r<C, Rings0>(p, tm_a0, tm_a1, tm_out, st);
  if (p.dbg_flags & 4096) return launch_tc_apply_r<C, Rings2>(p, tm_a0, tm_a1, tm_out, st);
  return launch_tc_apply_r<C, Rings1>(p, tm_a0, tm_a1, tm_out, st);
}

bool tc_shape_ok(int Cin, int Cout) { return Cin == Cout && (Cin == 128 || Cin == 256); }

}  // namespace

// 3-D fp32 tensor map [d2][d1][d0] (d0 contiguous), box = box1 x box0 elements of one d2 index, SWIZZLE_128B
// (box0 * 4 must be 128 bytes).  `tm` points at a 64-byte aligned CUtensorMap.
int wc_make_f32_map3(void* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return WC_E_UNSUPPORTED;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc((CUtensorMap*)tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : WC_E_UNSUPPORTED;
}

extern "C" size_t wc_pack_bytes(int C, int K, int nseg_shared) {
  if (!tc_shape_ok(C, C) || K <= 0) return 0;
  return (size_t)(K + nseg_shared) * C * C * 4;   // hi + lo bf16 per element
}

extern "C" int wc_pack_fwd(const float* T, int C, int K, void* packed, wc_stream_t stream) {
  if (!T || !packed || K <= 0) return WC_E_ARG;
  if (!tc_shape_ok(C, C)) return WC_E_UNSUPPORTED;
  long long chunks = (long long)K * C * (C / KB) * 4;
  WC_CUDA(wc_launch(pack_b_kernel, wc_cdiv(chunks, 256), 256, 0, (cudaStream_t)stream, T, C, K, (long long)C * C, 0, C, (uint8_t*)packed,
                                                                       (long long)C * C * 4));
  WC_CHECK_LAUNCH();
  return 0;
}

extern "C" int wc_pack_bwd(const float* T, const float* A, int C, int g, int K, void* packed, wc_stream_t stream) {
  if ((!T && !A) || !packed || K <= 0) return WC_E_ARG;
  if (!tc_shape_ok(C, C)) return WC_E_UNSUPPORTED;
  if (A && (g <= 0 || C % g)) return WC_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  long long chunks = (long long)K * C * (C / KB) * 4;
  if (T) {    // T == NULL: T's image is already in place (packed once in forward), only A's is refreshed
    WC_CUDA(wc_launch(pack_b_kernel, wc_cdiv(chunks, 256), 256, 0, st, T, C, K, (long long)C * C, 1, C, (uint8_t*)packed, (long long)C * C * 4));
    WC_CHECK_LAUNCH();
  }
  if (A) {
    chunks = (long long)C * (C / KB) * 4;
    WC_CUDA(wc_launch(pack_b_kernel, wc_cdiv(chunks, 256), 256, 0, st, A, C, 1, 0, 2, g, (uint8_t*)packed + (size_t)K * C * C * 4, 0));
    WC_CHECK_LAUNCH();
  }
  return 0;
}

// called from wc_apply.cu
int wc_tc_apply_fwd(const float* x, int N, int HW, int C, const float* mean, const void* packed, const float* b,
                    const int32_t* cls, int K, float* y, int32_t* info, int relu, int early, cudaStream_t st) {
  TcApplyParams p{};
  p.early = (wc_debug_flags() & 1073741824) ? 0 : early;
  p.seg[0] = TcSeg{mean, (const uint8_t*)packed, (long long)C * C * 4};
  p.nseg = 1;
  p.bias = b; p.bias_slot_stride = C; p.bias_sign = 1.f;
  p.cls = cls; p.K = K; p.info = info; p.relu = relu;
  const int l2 = wc_l2_mode();
  p.l2_load[0] = (l2 & 2) ? 2 : 0;    // x: last use of the forward pass
  p.l2_store = (l2 & 4) ? 2 : 0;      // y: written once, read by another layer much later
  return C == 256 ? launch_tc_apply<256>(p, x, nullptr, y, N, HW, st) : launch_tc_apply<128>(p, x, nullptr, y, N, HW, st);
}

int wc_tc_apply_bwd(const float* x, const float* dy, int N, int HW, int C, const float* mean, const void* packed,
                    bool has_A, const float* cvec, const int32_t* cls, int K, float* dx, int early, cudaStream_t st) {
  TcApplyParams p{};
  p.early = (wc_debug_flags() & 1073741824) ? 0 : early;
  p.seg[0] = TcSeg{nullptr, (const uint8_t*)packed, (long long)C * C * 4};
  p.nseg = 1;
  if (has_A) {
    p.seg[1] = TcSeg{mean, (const uint8_t*)packed + (size_t)K * C * C * 4, 0};
    p.nseg = 2;
  }
  p.bias = has_A ? cvec : nullptr; p.bias_slot_stride = 0; p.bias_sign = -1.f;
  p.cls = cls; p.K = K; p.info = nullptr;
  const int l2 = wc_l2_mode();
  p.l2_load[0] = p.l2_load[1] = (l2 & 16) ? 2 : 0;
  p.l2_store = (l2 & 32) ? 2 : 0;
  return C == 256 ? launch_tc_apply<256>(p, dy, has_A ? x : nullptr, dx, N, HW, st)
                  : launch_tc_apply<128>(p, dy, has_A ? x : nullptr, dx, N, HW, st);
}

bool wc_tc_apply_ok(int Cin, int Cout) { return tc_shape_ok(Cin, Cout); }

extern "C" int wc_path_for(int64_t M, int Cin, int Cout, int g) {
  int path = 0;
  if (tc_shape_ok(Cin, Cout)) path |= 3;   // bit0: tcgen05 apply kernels, bit1: tcgen05 gram kernels
  return path;
}
