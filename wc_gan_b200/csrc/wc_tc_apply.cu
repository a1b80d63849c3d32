// K3 / K6 on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM) for C in {128, 256}.
//
//   out[n][r, :] = sum_seg (A_seg[n][r, :] - mean_seg) . B_seg[slot(n)]  + sign * bias[slot(n)]
//
// fp32 accuracy on bf16 tensor cores: every fp32 operand is split x = hi + lo (two bf16, 16 mantissa
// bits) and the product is accumulated as hi.hi + hi.lo + lo.hi in the fp32 TMEM accumulator ("bf16x3";
// the dropped lo.lo term is 2^-18 relative).
//
// One CTA computes one 128-row x C output tile (rows of one sample), two CTAs are resident per SM so one
// tile's epilogue overlaps the other's main loop.  Warp roles (320 threads):
//   warps 0-7  A converters: coalesced 128-bit loads of fp32 activations -> centre -> bf16 hi/lo ->
//              st.shared into the SWIZZLE_64B K-major operand tiles; afterwards the epilogue
//              (tcgen05.ld -> smem transpose -> coalesced stores, + bias)
//   warp  8    MMA issuer (one elected thread), owns the TMEM allocation
//   warp  9    B loader: cp.async.bulk of pre-packed bf16 hi/lo operand images (written by wc_pack_*)
// Stages are handed over with mbarriers (full: 8 converter warps + 1 expect_tx arrival; empty: tcgen05.commit).
#include "wc_common.cuh"
#include "wc_tc.cuh"

using namespace wctc;

namespace {

constexpr int KB = 32;                 // K elements per stage (64-byte rows, SWIZZLE_64B)
constexpr int TILE_M = 128;
constexpr int NCONV_WARPS = 8;
constexpr int NTHREADS = (NCONV_WARPS + 2) * 32;

struct TcSeg {
  const float* A;        // activations [rows][C]
  const float* mean;     // per-channel shift (nullable)
  const uint8_t* Bp;     // packed operand images: [slot][kb][hi|lo][C x 64 B]
  long long slot_stride; // bytes between slots (0 = shared)
};

struct TcApplyParams {
  TcSeg seg[2];
  int nseg;
  const float* bias; long long bias_slot_stride; float bias_sign;
  int HW, tiles_per_sample;
  const int32_t* cls; int K; int32_t* info;
  float* out;
};

template <int C>
struct Cfg {
  static constexpr int N = C;
  static constexpr int NKB = C / KB;                       // K blocks per segment
  static constexpr int A_IMG = TILE_M * KB * 2;            // 8 KB
  static constexpr int B_IMG = N * KB * 2;                 // 16 KB (C=256) / 8 KB (C=128)
  static constexpr int STAGE = 2 * A_IMG + 2 * B_IMG;      // 48 KB / 32 KB
  static constexpr int STAGES = (C == 256) ? 2 : 3;
  static constexpr int SMEM = STAGES * STAGE + 1024;       // + alignment slack
};

template <int C>
__global__ void __launch_bounds__(NTHREADS, 2) tc_apply_kernel(TcApplyParams p) {
  using cfg = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[cfg::STAGES], empty_bar[cfg::STAGES], acc_bar;
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int sample = blockIdx.x / p.tiles_per_sample;
  const int rt = blockIdx.x % p.tiles_per_sample;
  const int r0 = rt * TILE_M;
  const int rows_valid = min(TILE_M, p.HW - r0);
  const long long row_base = (long long)sample * p.HW + r0;
  int slot = 0;
  if (p.cls) {
    int c = p.cls[sample];
    if (c < 0 || c >= p.K) {
      if (p.info && t == 0 && rt == 0) atomicMax(&p.info[WC_INFO_BAD_CLASS], sample + 1);
      c = c < 0 ? 0 : p.K - 1;
    }
    slot = c;
  }

  if (t == 0) {
    for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(&full_bar[s], NCONV_WARPS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == NCONV_WARPS) { tmem_alloc(&tmem_slot, cfg::N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_slot;
  const int total_kb = p.nseg * cfg::NKB;

  if (warp == NCONV_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TILE_M, cfg::N, 0, 0);
      uint32_t acc = 0;
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % cfg::STAGES;
        const uint32_t ph = (it / cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * cfg::STAGE);
        const uint32_t a_lo = a_hi + cfg::A_IMG;
        const uint32_t b_hi = a_hi + 2 * cfg::A_IMG;
        const uint32_t b_lo = b_hi + cfg::B_IMG;
#pragma unroll
        for (int kk = 0; kk < KB / 16; ++kk) {
          const uint64_t dah = make_desc(a_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
          const uint64_t dal = make_desc(a_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
          const uint64_t dbh = make_desc(b_hi + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
          const uint64_t dbl = make_desc(b_lo + kk * 32, 16, 8 * KB * 2, LAYOUT_SW64);
          umma_bf16(td, dah, dbh, idesc, acc);
          umma_bf16(td, dah, dbl, idesc, 1);
          umma_bf16(td, dal, dbh, idesc, 1);
          acc = 1;
        }
        umma_commit(&empty_bar[s]);      // stage reusable once these MMAs have read it
      }
      umma_commit(&acc_bar);             // accumulator complete
    }
  } else if (warp == NCONV_WARPS + 1) {
    // ------------------------------------------------------------------ B loader
    if (lane == 0) {
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % cfg::STAGES;
        const uint32_t ph = (it / cfg::STAGES) & 1;
        const int sg = it / cfg::NKB, kb = it % cfg::NKB;
        const uint8_t* src = p.seg[sg].Bp + (long long)slot * p.seg[sg].slot_stride + (long long)kb * (2 * cfg::B_IMG);
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* dst = smem + s * cfg::STAGE + 2 * cfg::A_IMG;
        mbar_arrive_expect_tx(&full_bar[s], 2 * cfg::B_IMG);
        bulk_g2s(dst, src, cfg::B_IMG, &full_bar[s]);
        bulk_g2s(dst + cfg::B_IMG, src + cfg::B_IMG, cfg::B_IMG, &full_bar[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ A converters
    // K block = 128 rows x 32 channels = 512 chunks of 8 channels; thread handles chunks t and t+256:
    // row = q/4, chunk-in-row = q%4 -> 4 consecutive threads read one row's 128 contiguous bytes.
    const int crow0 = t >> 2, cch = t & 3;
    // Global loads run TWO K blocks ahead of the conversion (two register buffers): ~32 KB in flight per CTA,
    // which is what it takes to cover HBM latency at full bandwidth with two CTAs per SM.
    float4 va[2][2], vb[2][2];
    auto gload = [&](int it, float4 (&v)[2][2]) {
      const int sg = it / cfg::NKB, kb = it % cfg::NKB;
      const float* A = p.seg[sg].A;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = crow0 + h * 64;
        if (row < rows_valid) {
          const float4* src = reinterpret_cast<const float4*>(A + (row_base + row) * C + kb * KB + cch * 8);
          v[h][0] = wc_ldg_stream(src);
          v[h][1] = wc_ldg_stream(src + 1);
        } else {
          v[h][0] = v[h][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto process = [&](int it, float4 (&v)[2][2]) {
      const int s = it % cfg::STAGES;
      const uint32_t ph = (it / cfg::STAGES) & 1;
      const int sg = it / cfg::NKB, kb = it % cfg::NKB;
      float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
      if (p.seg[sg].mean) {
        const float4* mp = reinterpret_cast<const float4*>(p.seg[sg].mean + kb * KB + cch * 8);
        m0 = __ldg(mp); m1 = __ldg(mp + 1);
      }
      uint32_t hi[2][4], lo[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = (crow0 + h * 64) < rows_valid;   // rows past the tile end stay exactly zero
        float4 a = v[h][0], b = v[h][1];
        if (ok) { a.x -= m0.x; a.y -= m0.y; a.z -= m0.z; a.w -= m0.w; b.x -= m1.x; b.y -= m1.y; b.z -= m1.z; b.w -= m1.w; }
        split2(a.x, a.y, hi[h][0], lo[h][0]);
        split2(a.z, a.w, hi[h][1], lo[h][1]);
        split2(b.x, b.y, hi[h][2], lo[h][2]);
        split2(b.z, b.w, hi[h][3], lo[h][3]);
      }
      if (it + 2 < total_kb) gload(it + 2, v);          // refill this buffer: its loads fly over the next block
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* a_hi = smem + s * cfg::STAGE;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = kmajor_chunk_off<KB>(crow0 + h * 64, cch);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[h][0], hi[h][1], hi[h][2], hi[h][3]);
        *reinterpret_cast<uint4*>(a_hi + cfg::A_IMG + off) = make_uint4(lo[h][0], lo[h][1], lo[h][2], lo[h][3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    };
    gload(0, va);
    if (total_kb > 1) gload(1, vb);
    for (int it = 0; it < total_kb; it += 2) {
      process(it, va);
      if (it + 1 < total_kb) process(it + 1, vb);
    }
    // ------------------------------------------------------------------ epilogue (same 8 warps)
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const float* bias = p.bias ? p.bias + (long long)slot * p.bias_slot_stride : nullptr;
    for (int c0 = half * (cfg::N / 2); c0 < (half + 1) * (cfg::N / 2); c0 += 32) {
      float acc[32];
      tmem_ld32(td + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = acc[j];
      __syncwarp();
      const float bv = bias ? p.bias_sign * __ldg(bias + c0 + lane) : 0.f;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = q * 32 + r;
        if (row < rows_valid) p.out[(row_base + row) * C + c0 + lane] = stg[r * 33 + lane] + bv;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCONV_WARPS) tmem_dealloc(td, cfg::N);
}

// ---- operand packing: fp32 matrix -> bf16 hi/lo K-major SWIZZLE_64B images -------------------------------
// image(kb)[n][k in 32] = src(n, kb*32 + k);  one thread per 16-byte chunk (8 k's of one n).
// mode 0: src(n,k) = M[k*ld + n] (K3: T^T);  mode 1: src(n,k) = M[n*ld + k] (K6 term 1: T);
// mode 2: src(n,k) = blockdiag(A (G,g,g))[k][n]  (K6 term 2).
__global__ void __launch_bounds__(256) pack_b_kernel(const float* __restrict__ M, int C, int nslots, long long slot_stride_in,
                                                     int mode, int g, uint8_t* __restrict__ out, long long slot_stride_out) {
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

template <int C>
int launch_tc_apply(TcApplyParams& p, int N, int HW, cudaStream_t st) {
  const bool cond = p.cls != nullptr && p.K > 1;
  if (!cond) { HW = N * HW; N = 1; p.cls = nullptr; }
  p.HW = HW;
  p.tiles_per_sample = wc_cdiv(HW, TILE_M);
  long long gx = (long long)N * p.tiles_per_sample;
  if (gx > INT32_MAX) return WC_E_UNSUPPORTED;
  static bool attr_set = false;   // idempotent; a race only repeats the same call
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(tc_apply_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<C>::SMEM));
    attr_set = true;
  }
  tc_apply_kernel<C><<<(unsigned)gx, NTHREADS, Cfg<C>::SMEM, st>>>(p);
  WC_CHECK_LAUNCH();
  return 0;
}

bool tc_shape_ok(int Cin, int Cout) { return Cin == Cout && (Cin == 128 || Cin == 256); }

}  // namespace

extern "C" size_t wc_pack_bytes(int C, int K, int nseg_shared) {
  if (!tc_shape_ok(C, C) || K <= 0) return 0;
  return (size_t)(K + nseg_shared) * C * C * 4;   // hi + lo bf16 per element
}

extern "C" int wc_pack_fwd(const float* T, int C, int K, void* packed, wc_stream_t stream) {
  if (!T || !packed || K <= 0) return WC_E_ARG;
  if (!tc_shape_ok(C, C)) return WC_E_UNSUPPORTED;
  long long chunks = (long long)K * C * (C / KB) * 4;
  pack_b_kernel<<<wc_cdiv(chunks, 256), 256, 0, (cudaStream_t)stream>>>(T, C, K, (long long)C * C, 0, C, (uint8_t*)packed,
                                                                       (long long)C * C * 4);
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
    pack_b_kernel<<<wc_cdiv(chunks, 256), 256, 0, st>>>(T, C, K, (long long)C * C, 1, C, (uint8_t*)packed, (long long)C * C * 4);
    WC_CHECK_LAUNCH();
  }
  if (A) {
    chunks = (long long)C * (C / KB) * 4;
    pack_b_kernel<<<wc_cdiv(chunks, 256), 256, 0, st>>>(A, C, 1, 0, 2, g, (uint8_t*)packed + (size_t)K * C * C * 4, 0);
    WC_CHECK_LAUNCH();
  }
  return 0;
}

// called from wc_apply.cu
int wc_tc_apply_fwd(const float* x, int N, int HW, int C, const float* mean, const void* packed, const float* b,
                    const int32_t* cls, int K, float* y, int32_t* info, cudaStream_t st) {
  TcApplyParams p{};
  p.seg[0] = TcSeg{x, mean, (const uint8_t*)packed, (long long)C * C * 4};
  p.nseg = 1;
  p.bias = b; p.bias_slot_stride = C; p.bias_sign = 1.f;
  p.cls = cls; p.K = K; p.info = info; p.out = y;
  return C == 256 ? launch_tc_apply<256>(p, N, HW, st) : launch_tc_apply<128>(p, N, HW, st);
}

int wc_tc_apply_bwd(const float* x, const float* dy, int N, int HW, int C, const float* mean, const void* packed,
                    bool has_A, const float* cvec, const int32_t* cls, int K, float* dx, cudaStream_t st) {
  TcApplyParams p{};
  p.seg[0] = TcSeg{dy, nullptr, (const uint8_t*)packed, (long long)C * C * 4};
  p.nseg = 1;
  if (has_A) {
    p.seg[1] = TcSeg{x, mean, (const uint8_t*)packed + (size_t)K * C * C * 4, 0};
    p.nseg = 2;
  }
  p.bias = has_A ? cvec : nullptr; p.bias_slot_stride = 0; p.bias_sign = -1.f;
  p.cls = cls; p.K = K; p.info = nullptr; p.out = dx;
  return C == 256 ? launch_tc_apply<256>(p, N, HW, st) : launch_tc_apply<128>(p, N, HW, st);
}

bool wc_tc_apply_ok(int Cin, int Cout) { return tc_shape_ok(Cin, Cout); }

extern "C" int wc_path_for(int64_t M, int Cin, int Cout, int g) {
  int path = 0;
  if (tc_shape_ok(Cin, Cout)) path |= 3;   // bit0: tcgen05 apply kernels, bit1: tcgen05 gram kernels
  return path;
}
