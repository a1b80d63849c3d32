// K1 (batch statistics) and K4 (backward reduction): "gram"-shaped reductions over rows,
//   Out[slot][i][j] += sum_r (A[r,i] - ma[i]) * (B[r,j] - mb[j]),
// FP32 FFMA path (any C, any g).  The tcgen05 path for C in {128,256} lives in wc_tc.cu.
//
// Reference segments replaced: gan/conditional_layers.py:652-660,:683 (K1) and TF autodiff of
// :707,:824,:1010 / generator.py:50 w.r.t. the matrices (K4).
#include "wc_common.cuh"
#include <algorithm>

namespace {

constexpr int TI = 64, TJ = 64, RK = 16, NT = 256;
constexpr int UNIT_ROWS = 256;  // a work unit never crosses a sample boundary

struct GramParams {
  const float* A; const float* B;   // [rows][lda], [rows][ldb]
  int lda, ldb;
  const float* ma; const float* mb; // per-column shifts (nullable)
  int nI, nJ;                       // columns of A / B taking part
  int grouped;                      // 1: only tiles on the block diagonal (A == B, nI == nJ)
  int g, tw, tpg;                   // group size, tile width along the diagonal (g < 64), tiles per group side (g >= 64)
  int ntiles;
  int N, HW;                        // samples, rows per sample
  int ups;                          // units per sample
  int units_per_chunk, nchunks;
  const int32_t* order;             // processing order of samples (nullable = identity)
  const int32_t* slot_of;           // slot (class) per sample, already clamped (nullable = 0)
  float* out; long long out_slot_stride; int ldo;  // Out[slot][i][j]
  float* asum; float* bsum; long long sum_slot_stride; // column sums per slot (nullable)
  int atomic;                       // 1: atomicAdd into out[slot]; 0: store into out[chunk] (slot ignored)
};

__device__ __forceinline__ void tile_bounds(const GramParams& p, int tile, int& i0, int& iend, int& j0, int& jend,
                                            bool& want_a, bool& want_b) {
  if (p.grouped) {
    if (p.g >= TI) {
      int per = p.tpg * p.tpg;
      int grp = tile / per, t = tile % per;
      int ti = t / p.tpg, tj = t % p.tpg;
      i0 = grp * p.g + ti * TI; iend = min(grp * p.g + p.g, i0 + TI);
      j0 = grp * p.g + tj * TJ; jend = min(grp * p.g + p.g, j0 + TJ);
      want_a = (tj == 0); want_b = false;
    } else {
      i0 = tile * p.tw; iend = min(p.nI, i0 + p.tw);
      j0 = i0; jend = iend;
      want_a = true; want_b = false;
    }
  } else {
    int tjn = (p.nJ + TJ - 1) / TJ;
    int ti = tile / tjn, tj = tile % tjn;
    i0 = ti * TI; iend = min(p.nI, i0 + TI);
    j0 = tj * TJ; jend = min(p.nJ, j0 + TJ);
    want_a = (tj == 0); want_b = (ti == 0);
  }
}

template <bool kVec>
__device__ __forceinline__ float4 load4(const float* base, long long row_off, int col, int cend, const float* shift, bool row_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  if (kVec) {
    if (col + 3 < cend) {
      v = wc_ldg_stream(reinterpret_cast<const float4*>(base + row_off + col));
      if (shift) { float4 s = *reinterpret_cast<const float4*>(shift + col); v.x -= s.x; v.y -= s.y; v.z -= s.z; v.w -= s.w; }
      return v;
    }
  }
  float t[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int c = col + q;
    t[q] = (c < cend) ? (__ldg(base + row_off + c) - (shift ? shift[c] : 0.f)) : 0.f;
  }
  return make_float4(t[0], t[1], t[2], t[3]);
}

template <bool kVec>
__global__ void __launch_bounds__(NT) gram_kernel(GramParams p) {
  wc_pdl_wait();
  __shared__ __align__(16) float As[2][RK][TI];
  __shared__ __align__(16) float Bs[2][RK][TJ];
  __shared__ float red[RK][TI];

  const int tile = blockIdx.x, chunk = blockIdx.y;
  int i0, iend, j0, jend; bool want_a, want_b;
  tile_bounds(p, tile, i0, iend, j0, jend, want_a, want_b);
  want_a = want_a && (p.asum != nullptr);
  want_b = want_b && (p.bsum != nullptr);

  const int t = threadIdx.x;
  const int lrow = t >> 4, lcol = (t & 15) * 4;   // loader mapping: row in step, first column
  const int tx = t & 15, ty = t >> 4;             // compute mapping: 4x4 micro tile

  const long long total_units = (long long)p.N * p.ups;
  long long u_begin = (long long)chunk * p.units_per_chunk;
  long long u_end = min(total_units, u_begin + p.units_per_chunk);

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float4 asum4 = make_float4(0, 0, 0, 0), bsum4 = make_float4(0, 0, 0, 0);

  // step iterator state
  long long unit = u_begin;
  int r_in_unit = 0;
  auto unit_info = [&](long long u, int& sample, int& row0, int& nrows) {
    int sidx = (int)(u / p.ups), sub = (int)(u % p.ups);
    sample = p.order ? p.order[sidx] : sidx;
    row0 = sub * UNIT_ROWS;
    nrows = min(UNIT_ROWS, p.HW - row0);
  };

  int cur_slot = -1;
  float4 ra, rb;
  int next_slot = 0;
  bool have = unit < u_end;
  auto fetch = [&]() {
    int sample, row0, nrows;
    unit_info(unit, sample, row0, nrows);
    next_slot = p.slot_of ? p.slot_of[sample] : 0;
    int r = r_in_unit + lrow;
    bool ok = r < nrows;
    long long grow = (long long)sample * p.HW + row0 + r;
    ra = load4<kVec>(p.A, grow * p.lda, i0 + lcol, iend, p.ma, ok);
    rb = load4<kVec>(p.B, grow * p.ldb, j0 + lcol, jend, p.mb, ok);
    r_in_unit += RK;
    if (r_in_unit >= nrows) { r_in_unit = 0; ++unit; }
  };
  auto flush = [&](int slot) {
    float* o = p.atomic ? p.out + (long long)slot * p.out_slot_stride
                        : p.out + (long long)chunk * p.out_slot_stride;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int i = i0 + ty * 4 + a;
      if (i >= iend) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int j = j0 + tx * 4 + b;
        if (j >= jend) continue;
        if (p.atomic) atomicAdd(o + (long long)i * p.ldo + j, acc[a][b]);
        else o[(long long)i * p.ldo + j] = acc[a][b];
        acc[a][b] = 0.f;
      }
    }
    if (want_a || want_b) {
      for (int pass = 0; pass < 2; ++pass) {
        bool want = pass == 0 ? want_a : want_b;
        if (!want) continue;
        float4 v = pass == 0 ? asum4 : bsum4;
        __syncthreads();
        red[lrow][lcol + 0] = v.x; red[lrow][lcol + 1] = v.y; red[lrow][lcol + 2] = v.z; red[lrow][lcol + 3] = v.w;
        __syncthreads();
        if (t < TI) {
          float sacc = 0.f;
#pragma unroll
          for (int q = 0; q < RK; ++q) sacc += red[q][t];
          int c0 = pass == 0 ? i0 : j0, cend = pass == 0 ? iend : jend;
          float* so = (pass == 0 ? p.asum : p.bsum) +
                      (p.atomic ? (long long)slot : (long long)chunk) * p.sum_slot_stride;
          if (c0 + t < cend) { if (p.atomic) atomicAdd(so + c0 + t, sacc); else so[c0 + t] = sacc; }
        }
      }
      asum4 = make_float4(0, 0, 0, 0); bsum4 = make_float4(0, 0, 0, 0);
      __syncthreads();
    }
  };

  if (have) {
    fetch();
    *reinterpret_cast<float4*>(&As[0][lrow][lcol]) = ra;
    *reinterpret_cast<float4*>(&Bs[0][lrow][lcol]) = rb;
    if (want_a) { asum4.x += ra.x; asum4.y += ra.y; asum4.z += ra.z; asum4.w += ra.w; }
    if (want_b) { bsum4.x += rb.x; bsum4.y += rb.y; bsum4.z += rb.z; bsum4.w += rb.w; }
    cur_slot = next_slot;
  }
  __syncthreads();
  int buf = 0;
  while (have) {
    bool more = unit < u_end;
    int slot_after = cur_slot;
    if (more) { fetch(); slot_after = next_slot; }
    // compute current step
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    if (!more) break;
    if (slot_after != cur_slot) {  // block-uniform: all threads read the same unit
      flush(cur_slot);
      cur_slot = slot_after;
    }
    *reinterpret_cast<float4*>(&As[buf ^ 1][lrow][lcol]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf ^ 1][lrow][lcol]) = rb;
    if (want_a) { asum4.x += ra.x; asum4.y += ra.y; asum4.z += ra.z; asum4.w += ra.w; }
    if (want_b) { bsum4.x += rb.x; bsum4.y += rb.y; bsum4.z += rb.z; bsum4.w += rb.w; }
    __syncthreads();
    buf ^= 1;
  }
  if (cur_slot >= 0) flush(cur_slot);
  else if (!p.atomic) {  // empty chunk in store mode: partials must still be defined
    cur_slot = 0; flush(0);
  }
}

// ---- K1 helpers ---------------------------------------------------------------------------
// pivot[c] = mean of the first `rows` rows: a cheap shift that removes the E[xx^T]-mu mu^T cancellation
__global__ void pivot_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ pivot) {
  wc_pdl_wait();
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int w = threadIdx.x >> 5;  // 8 warps stride over rows
  __shared__ float part[8][32];
  float s = 0.f;
  if (c < C) for (int r = w; r < rows; r += 8) s += x[(long long)r * C + c];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float a = 0.f;
    for (int q = 0; q < 8; ++q) a += part[q][threadIdx.x];
    pivot[c] = a / (float)rows;
  }
}

// reduce chunk partials (fixed order => deterministic), correct for the pivot:
//   d = sum(x-p)/M, mean = p + d, S = S_p - M d d^T   (diagonal blocks only)
__global__ void stats_finalize_kernel(const float* __restrict__ partS, const float* __restrict__ partSum,
                                      int nchunks, long long strideS, int C, int g, double M,
                                      const float* __restrict__ pivot, float* __restrict__ mean, float* __restrict__ S,
                                      int lower_only) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)C * g;
  if (idx >= total) return;
  int grp = (int)(idx / ((long long)g * g));
  int rem = (int)(idx % ((long long)g * g));
  int i = grp * g + rem / g, j = grp * g + rem % g;
  float s = 0.f, di = 0.f, dj = 0.f;
  for (int q = 0; q < nchunks; ++q) {
    // the tensor-core path only accumulates the blocks with j <= i (the scatter is symmetric)
    s += lower_only ? partS[(long long)q * strideS + (long long)max(i, j) * C + min(i, j)]
                    : partS[(long long)q * strideS + (long long)i * C + j];
    di += partSum[(long long)q * C + i];
    dj += partSum[(long long)q * C + j];
  }
  double ddi = (double)di / M, ddj = (double)dj / M;
  S[idx] = (float)((double)s - M * ddi * ddj);
  if (rem % g == rem / g) mean[i] = (float)((double)pivot[i] + ddi);
}

// Tensor-core path, single process: finalize and the K2 prep step (covariance scaling, moving averages, shrinkage;
// see factor_prep_kernel in wc_small.cu) in one launch.  The scatter is read from its lower triangle, so S and the
// covariance are exactly symmetric.
// Accumulator zeroing as a KERNEL instead of a memset node: a memset breaks the chain of programmatic dependent launches
// (the kernel after it cannot be scheduled, let alone run its prologue, before the memset has finished, and the memset
// cannot start before the kernel before it has drained).  The stores come after the dependency wait: the kernel before
// this one may still be reading the buffer (finalize of the previous call on the same workspace).
// (p2, n16_2), (p3, n16_3): optional further ranges (the adjoint accumulator that the split class sums of the backward chain
// add into; the scratch matrices its K-split products add into)
__global__ void __launch_bounds__(256) zero_kernel(uint4* __restrict__ p, size_t n16, uint32_t* __restrict__ tail, int ntail,
                                                   uint4* __restrict__ p2, size_t n16_2, uint4* __restrict__ p3, size_t n16_3) {
  wc_pdl_wait();
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16_2; i += (size_t)gridDim.x * blockDim.x) p2[i] = z;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16_3; i += (size_t)gridDim.x * blockDim.x) p3[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0u;
}
// *by_kernel (nullable) = 1 when the zeroing was done by zero_kernel, i.e. by a kernel that waits for everything before it on
// the stream before it lets its successor start (what tc_gram_kernel's late dependency wait relies on), 0 for a memset.
// (ptr2, bytes2): optional second range, a multiple of 16 bytes
int zero_launch(void* ptr, size_t bytes, cudaStream_t st, int* by_kernel = nullptr, void* ptr2 = nullptr, size_t bytes2 = 0,
                void* ptr3 = nullptr, size_t bytes3 = 0) {
  if (by_kernel) *by_kernel = 0;
  if (!ptr2) bytes2 = 0;
  if (!ptr3) bytes3 = 0;
  if (!wc_aligned16(ptr) || (bytes & 3) || !wc_aligned16(ptr2) || (bytes2 & 15) || !wc_aligned16(ptr3) || (bytes3 & 15) ||
      (wc_debug_flags() & 4096 * 65536 * 2)) {   // experiment bit 29: memset node
    WC_CUDA(cudaMemsetAsync(ptr, 0, bytes, st));
    if (bytes2) WC_CUDA(cudaMemsetAsync(ptr2, 0, bytes2, st));
    if (bytes3) WC_CUDA(cudaMemsetAsync(ptr3, 0, bytes3, st));
    return 0;
  }
  const size_t n16 = bytes / 16;
  const int ntail = (int)((bytes - n16 * 16) / 4);
  const size_t want = (n16 + bytes2 / 16 + bytes3 / 16 + 255) / 256;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > (size_t)WC_NUM_SMS * 4 ? (size_t)WC_NUM_SMS * 4 : want));
  WC_CUDA(wc_launch(zero_kernel, grid, 256, 0, st, reinterpret_cast<uint4*>(ptr), n16,
                    reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ptr) + n16 * 16), ntail,
                    reinterpret_cast<uint4*>(ptr2), bytes2 / 16, reinterpret_cast<uint4*>(ptr3), bytes3 / 16));
  WC_CHECK_LAUNCH();
  if (by_kernel) *by_kernel = 1;
  return 0;
}

__global__ void stats_finalize_prep_kernel(const float* __restrict__ partS, const float* __restrict__ partSum, int C, int g,
                                           double M, const float* __restrict__ pivot, float* __restrict__ mean,
                                           float* __restrict__ S, WcPrepFuse f) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)C * g;
  if (idx >= total) { wc_pdl_wait(); return; }
  int grp = (int)(idx / ((long long)g * g));
  int rem = (int)(idx % ((long long)g * g));
  int r = rem / g, c = rem % g;
  int i = grp * g + r, j = grp * g + c;
  // the moving averages do not depend on the statistics kernel: their (cold, DRAM) loads are issued BEFORE the dependency
  // wait and complete while that kernel drains
  const float mc_old = f.moving_cov[idx];
  const float mm_old = (r == c) ? f.moving_mean[i] : 0.f;
  wc_pdl_wait();
  const float s = partS[(long long)max(i, j) * C + min(i, j)];
  const float di = partSum[i], dj = partSum[j];
  const double ddi = (double)di / M, ddj = (double)dj / M;
  const float sv = (float)((double)s - M * ddi * ddj);
  S[idx] = sv;
  const float cov = sv / ((float)M - 1.f);
  f.moving_cov[idx] = mc_old * f.momentum + cov * (1.f - f.momentum);        // conditional_layers.py:687-689
  if (r == c) {
    const float mu = (float)((double)pivot[i] + ddi);
    mean[i] = mu;
    f.moving_mean[i] = mm_old * f.momentum + mu * (1.f - f.momentum);          // :684-686
  }
  f.Ashr[idx] = (1.f - f.eps) * cov + (r == c ? f.eps + f.extra_eps : 0.f);              // :691 (+ :671 for ZCA)
}

// Chan merge of R shard records {mean[C], S[G g g], count}
__global__ void stats_merge_kernel(const float* __restrict__ rec, int R, int C, int g,
                                   float* __restrict__ mean, float* __restrict__ S, float* __restrict__ count_total) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)C * g;
  if (idx >= total) return;
  long long stride = C + total + 1;
  int grp = (int)(idx / ((long long)g * g));
  int rem = (int)(idx % ((long long)g * g));
  int i = grp * g + rem / g, j = grp * g + rem % g;
  double n = 0, mi = 0, mj = 0;
  for (int r = 0; r < R; ++r) {
    double nr = rec[r * stride + C + total];
    n += nr; mi += nr * rec[r * stride + i]; mj += nr * rec[r * stride + j];
  }
  mi /= n; mj /= n;
  double s = 0;
  for (int r = 0; r < R; ++r) {
    double nr = rec[r * stride + C + total];
    double di = rec[r * stride + i] - mi, dj = rec[r * stride + j] - mj;
    s += rec[r * stride + C + idx] + nr * di * dj;
  }
  S[idx] = (float)s;
  if (rem % g == rem / g) mean[i] = (float)mi;
  if (idx == 0) count_total[0] = (float)n;
}

// ---- K4 helper: stable counting sort of samples by class (one CTA) -----------------------------
__global__ void class_order_kernel(const int32_t* __restrict__ cls, int N, int K,
                                   int32_t* __restrict__ order, int32_t* __restrict__ slot_of) {
  wc_pdl_wait();
  extern __shared__ int sh[];  // [K] starts
  int* start = sh;
  for (int k = threadIdx.x; k < K; k += blockDim.x) start[k] = 0;
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int c = cls[n]; c = c < 0 ? 0 : (c >= K ? K - 1 : c);
    slot_of[n] = c;
    atomicAdd(&start[c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < K; ++k) { int c = start[k]; start[k] = run; run += c; }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int c = slot_of[n], rank = 0;
    for (int m = 0; m < n; ++m) rank += (slot_of[m] == c);   // written above by this or another thread, after the barrier
    order[start[c] + rank] = n;
  }
}

struct StatsPlan { int ntiles, nchunks, units_per_chunk, ups, tw, tpg; size_t off_pivot, off_zero, off_sum, off_S, bytes; };

StatsPlan plan_stats(int64_t M, int C, int g) {
  StatsPlan pl{};
  if (g >= TI) { pl.tpg = wc_cdiv(g, TI); pl.ntiles = (C / g) * pl.tpg * pl.tpg; pl.tw = TI; }
  else { pl.tw = (TI / g) * g; pl.tpg = 1; pl.ntiles = wc_cdiv(C, pl.tw); }
  pl.ups = wc_cdiv(M, UNIT_ROWS);
  int want = std::max(1, (4 * WC_NUM_SMS) / pl.ntiles);
  int by_rows = (int)(M / 512 > 0 ? M / 512 : 1);
  pl.nchunks = std::min(std::min(want, by_rows), pl.ups);
  pl.units_per_chunk = wc_cdiv(pl.ups, pl.nchunks);
  pl.nchunks = wc_cdiv(pl.ups, pl.units_per_chunk);
  pl.off_pivot = 0;
  pl.off_zero = wc_align_up((size_t)C * 4, 256);      // 256 bytes cleared together with the accumulators (tensor-core path)
  pl.off_sum = pl.off_zero + 256;
  pl.off_S = pl.off_sum + wc_align_up((size_t)pl.nchunks * C * 4, 256);
  pl.bytes = pl.off_S + (size_t)pl.nchunks * C * C * 4;
  return pl;
}

}  // namespace

bool wc_tc_gram_ok(int Cin, int Cout);
int wc_tc_gram_self(const float* x, long long M, int C, int g, float* pivot_out, float* Sp, float* sums, int late_wait, cudaStream_t st);
int wc_tc_gram_cross(const float* x, const float* dy, int N, int HW, int C, const float* mean, const int32_t* order,
                     const int32_t* slot_of, int K, float* P, float* s, int late_wait, cudaStream_t st);

static bool g_use_tc_gram = true;
extern "C" void wc_set_tensor_core_gram(int enable) { g_use_tc_gram = enable != 0; }

extern "C" size_t wc_stats_workspace_bytes(int64_t M, int C, int g) {
  if (M <= 0 || C <= 0 || g <= 0 || C % g) return 0;
  return plan_stats(M, C, g).bytes;
}

extern "C" int wc_stats_fwd(const float* x, int64_t M, int C, int g, float* workspace, size_t workspace_bytes,
                            float* mean, float* S, wc_stream_t stream) {
  return wc_stats_fwd_impl(x, M, C, g, workspace, workspace_bytes, mean, S, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// fuse (nullable): also run the prep step of wc_factor_fwd when the tensor-core path is taken; *fused reports whether
// that happened (the caller then tells wc_factor_fwd_impl to skip its own prep launch)
int wc_stats_fwd_impl(const float* x, int64_t M, int C, int g, float* workspace, size_t workspace_bytes,
                      float* mean, float* S, const WcPrepFuse* fuse, int* fused, unsigned** zeroed_out, WcRawStats* raw_out,
                      cudaStream_t stream) {
  // raw_out (nullable): on the tensor-core path, do NOT launch the finalize kernel; report where the raw accumulators are
  // (raw_out->partS != NULL) and let the caller's next kernel finalize.  Otherwise raw_out->partS is NULL.
  if (fused) *fused = 0;
  if (zeroed_out) *zeroed_out = nullptr;
  if (raw_out) *raw_out = WcRawStats{nullptr, nullptr, nullptr};
  if (!x || !workspace || !mean || !S || M <= 0 || C <= 0 || g <= 0 || C % g) return WC_E_ARG;
  if (M > (int64_t)INT32_MAX) return WC_E_UNSUPPORTED;
  StatsPlan pl = plan_stats(M, C, g);
  if (workspace_bytes < pl.bytes) return WC_E_WORKSPACE;
  if (!wc_aligned16(workspace)) return WC_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  float* pivot = reinterpret_cast<float*>(ws + pl.off_pivot);
  float* psum = reinterpret_cast<float*>(ws + pl.off_sum);
  float* pS = reinterpret_cast<float*>(ws + pl.off_S);
  if (g_use_tc_gram && wc_tc_gram_ok(C, C) && wc_aligned16(x)) {
    // tensor-core path: the full C x C scatter is accumulated with vector reductions into one zeroed buffer
    // (partial slot 0; column sums and scatter are adjacent: one memset); the kernel computes its own pivot and
    // the finalize kernel keeps the diagonal blocks when g < C.
    int zk = 0;
    { const int rz = zero_launch(ws + pl.off_zero, (pl.off_S - pl.off_zero) + (size_t)C * C * 4, st, &zk); if (rz) return rz; }
    if (zeroed_out) *zeroed_out = reinterpret_cast<unsigned*>(ws + pl.off_zero);
    int rc = wc_tc_gram_self(x, M, C, g, pivot, pS, psum, zk, st);
    if (rc) return rc;
    long long tot = (long long)C * g;
    if (raw_out) { *raw_out = WcRawStats{pS, psum, pivot}; return 0; }
    if (fuse) {
      WC_CUDA(wc_launch(stats_finalize_prep_kernel, wc_cdiv(tot, 256), 256, 0, st, pS, psum, C, g, (double)M, pivot, mean, S, *fuse));
      WC_CHECK_LAUNCH();
      if (fused) *fused = 1;
      return 0;
    }
    WC_CUDA(wc_launch(stats_finalize_kernel, wc_cdiv(tot, 256), 256, 0, st, pS, psum, 1, (long long)C * C, C, g, (double)M, pivot, mean, S, 1));
    WC_CHECK_LAUNCH();
    return 0;
  }
  int prow = (int)(M < 1024 ? M : 1024);
  WC_CUDA(wc_launch(pivot_kernel, wc_cdiv(C, 32), 256, 0, st, x, prow, C, pivot));
  WC_CHECK_LAUNCH();
  GramParams p{};
  p.A = x; p.B = x; p.lda = C; p.ldb = C; p.ma = pivot; p.mb = pivot; p.nI = C; p.nJ = C;
  p.grouped = 1; p.g = g; p.tw = pl.tw; p.tpg = pl.tpg; p.ntiles = pl.ntiles;
  p.N = 1; p.HW = (int)M; p.ups = pl.ups; p.units_per_chunk = pl.units_per_chunk; p.nchunks = pl.nchunks;
  p.order = nullptr; p.slot_of = nullptr;
  p.out = pS; p.out_slot_stride = (long long)C * C; p.ldo = C;
  p.asum = psum; p.bsum = nullptr; p.sum_slot_stride = C; p.atomic = 0;
  dim3 grid(pl.ntiles, pl.nchunks);
  bool vec = (C % 4 == 0) && (g % 4 == 0) && wc_aligned16(x);
  if (vec) WC_CUDA(wc_launch(gram_kernel<true>, grid, NT, 0, st, p)); else WC_CUDA(wc_launch(gram_kernel<false>, grid, NT, 0, st, p));
  WC_CHECK_LAUNCH();
  long long total = (long long)C * g;
  WC_CUDA(wc_launch(stats_finalize_kernel, wc_cdiv(total, 256), 256, 0, st, pS, psum, pl.nchunks, (long long)C * C, C, g, (double)M, pivot, mean, S, 0));
  WC_CHECK_LAUNCH();
  return 0;
}

extern "C" int wc_stats_merge(const float* records, int R, int C, int g, float* mean, float* S, float* count_total,
                              wc_stream_t stream) {
  if (!records || !mean || !S || !count_total || R <= 0 || C <= 0 || g <= 0 || C % g) return WC_E_ARG;
  long long total = (long long)C * g;
  WC_CUDA(wc_launch(stats_merge_kernel, wc_cdiv(total, 256), 256, 0, (cudaStream_t)stream, records, R, C, g, mean, S, count_total));
  WC_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t wc_reduce_workspace_bytes(int N, int HW, int Cin, int Cout, int K) {
  if (N <= 0 || HW <= 0 || Cin <= 0 || Cout <= 0 || K <= 0) return 0;
  return wc_align_up((size_t)N * 4, 256) * 2;
}

extern "C" int wc_reduce_bwd(const float* x, const float* dy, int N, int HW, int Cin, int Cout,
                             const float* mean, const int32_t* cls, int K,
                             float* P, float* s, float* workspace, size_t workspace_bytes, wc_stream_t stream) {
  return wc_reduce_bwd_impl(x, dy, N, HW, Cin, Cout, mean, cls, K, P, s, workspace, workspace_bytes, 0, nullptr, 0, nullptr, 0, (cudaStream_t)stream);
}

// zero_tail_floats: floats after s (P | s | tail contiguous) that the accumulator memset also clears (the grid-barrier
// counter of the chain launch that consumes P and s); (zero2, zero2_floats): one more buffer cleared by the same launch
// (the adjoint accumulator of that chain; its K-split scratch matrices), nullable
int wc_reduce_bwd_impl(const float* x, const float* dy, int N, int HW, int Cin, int Cout, const float* mean, const int32_t* cls,
                       int K, float* P, float* s, float* workspace, size_t workspace_bytes, int zero_tail_floats,
                       float* zero2, size_t zero2_floats, float* zero3, size_t zero3_floats, cudaStream_t stream) {
  if (!x || !dy || !P || !s || N <= 0 || HW <= 0 || Cin <= 0 || Cout <= 0 || K <= 0) return WC_E_ARG;
  if ((int64_t)N * HW > (int64_t)INT32_MAX) return WC_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const bool cond = (cls != nullptr) && K > 1;
  int32_t *order = nullptr, *slot_of = nullptr;
  if (cond) {
    size_t need = wc_reduce_workspace_bytes(N, HW, Cin, Cout, K);
    if (!workspace || workspace_bytes < need) return WC_E_WORKSPACE;
    if (K > 8192) return WC_E_UNSUPPORTED;
    order = reinterpret_cast<int32_t*>(workspace);
    slot_of = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + need / 2);
    WC_CUDA(wc_launch(class_order_kernel, 1, 1024, K * sizeof(int), st, cls, N, K, order, slot_of));
    WC_CHECK_LAUNCH();
  }
  int zk = 0;
  if (s == P + (size_t)K * Cin * Cout) {   // P | s contiguous (wc_layer_bwd's workspace): one memset
    const int rz = zero_launch(P, ((size_t)K * Cin * Cout + (size_t)K * Cout + (size_t)zero_tail_floats) * 4, st, &zk, zero2,
                               zero2_floats * 4, zero3, zero3_floats * 4);
    if (rz) return rz;
  } else {
    if (zero_tail_floats || zero2 || zero3) return WC_E_ARG;
    WC_CUDA(cudaMemsetAsync(P, 0, (size_t)K * Cin * Cout * 4, st));
    WC_CUDA(cudaMemsetAsync(s, 0, (size_t)K * Cout * 4, st));
  }
  if (g_use_tc_gram && wc_tc_gram_ok(Cin, Cout) && wc_aligned16(x) && wc_aligned16(dy) && wc_aligned16(P) &&
      (!mean || wc_aligned16(mean)))
    return wc_tc_gram_cross(x, dy, N, HW, Cin, mean, order, slot_of, cond ? K : 1, P, s, zk, st);
  GramParams p{};
  p.A = x; p.B = dy; p.lda = Cin; p.ldb = Cout; p.ma = mean; p.mb = nullptr; p.nI = Cin; p.nJ = Cout;
  p.grouped = 0; p.g = Cin; p.tw = TI; p.tpg = 1;
  p.ntiles = wc_cdiv(Cin, TI) * wc_cdiv(Cout, TJ);
  // unconditional: one virtual sample of all rows so that units are full
  if (cond) { p.N = N; p.HW = HW; } else { p.N = 1; p.HW = N * HW; }
  p.ups = wc_cdiv(p.HW, UNIT_ROWS);
  long long units = (long long)p.N * p.ups;
  long long rows = (long long)N * HW;
  long long want = std::max(1, (4 * WC_NUM_SMS) / p.ntiles);
  long long by_rows = rows / 512 > 0 ? rows / 512 : 1;
  long long nch = std::min(std::min(want, by_rows), units);
  p.units_per_chunk = (int)((units + nch - 1) / nch);
  p.nchunks = (int)((units + p.units_per_chunk - 1) / p.units_per_chunk);
  p.order = order; p.slot_of = slot_of;
  p.out = P; p.out_slot_stride = (long long)Cin * Cout; p.ldo = Cout;
  p.asum = nullptr; p.bsum = s; p.sum_slot_stride = Cout; p.atomic = 1;
  dim3 grid(p.ntiles, p.nchunks);
  bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && wc_aligned16(x) && wc_aligned16(dy) && (!mean || wc_aligned16(mean));
  if (vec) WC_CUDA(wc_launch(gram_kernel<true>, grid, NT, 0, st, p)); else WC_CUDA(wc_launch(gram_kernel<false>, grid, NT, 0, st, p));
  WC_CHECK_LAUNCH();
  return 0;
}
