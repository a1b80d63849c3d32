// K2 / K2b / K5a / K5b: the small-matrix stages of the WC layer (everything that is C x C,
// never M x C).  These are latency-bound; see DESIGN.md for the launch list.
//
// Reference segments replaced: gan/conditional_layers.py:662-704 (factorisation, moving
// averages, shrinkage, renorm, test path), generator.py:41-78 + conditional_layers.py:811-819,
// :993-1005 (per-class kernels), and TF autodiff of all of them.
#include "wc_common.cuh"
#include <math.h>
#include <stdlib.h>

// =============================================================================================
// generic batched small matmul
// =============================================================================================
namespace {

struct SmmExtra { const float* count_dev; float count_host; int count_mode; float diag_scale; };

// ---- fast path: one CTA computes a 32 x 32 tile of C ------------------------------------------------------------
// Latency is what matters here (a 256^3 product is ~1 us of FP32 FFMA spread over 64 CTAs), so each K panel (up to 256)
// of both operands is fetched with ONE batch of 128-bit loads per operand into shared memory as As[k][i], Bs[k][j]
// (operands whose k index is the contiguous one are transposed on the way in), and the tile is then computed from
// there with 4 x 2 register blocking (128 threads).  Requires M, N multiples of 32, K a multiple of 4, unit fast
// strides and 16-byte aligned rows; anything else takes smm_generic_kernel.
constexpr int SMM_T = 32, SMM_KC = 256, SMM_P = 36, SMM_THREADS = 512, SMM_KG = 4;   // 4 k-groups of 128 threads
constexpr int SMM_SMEM = 2 * SMM_KC * SMM_P * 4;
constexpr int SMM_NV = SMM_KC * SMM_T / 4 / SMM_THREADS;   // float4 per thread per operand panel: 4

// panel(k, m) for k < kc, m < 32 -> dst[k * SMM_P + m].  kfast: element (m, k) at src[m * ld + k]; else at src[k * ld + m]
template <bool kKFast>
__device__ __forceinline__ void smm_load_panel(const float* __restrict__ src, long long ld, int kc, float4 (&v)[SMM_NV], int t) {
  if (kKFast) {
    const int m = t & 31, k4w = t >> 5;            // lanes walk m: conflict-free transposed stores
    const int nk4 = kc >> 2;
#pragma unroll
    for (int q = 0; q < SMM_NV; ++q) {
      const int k4 = k4w + q * (SMM_THREADS / 32);
      if (k4 < nk4) v[q] = __ldg(reinterpret_cast<const float4*>(src + (long long)m * ld) + k4);
    }
  } else {
    const int c4 = t & 7, kw = t >> 3;             // 8 lanes cover the 128 contiguous bytes of one k
#pragma unroll
    for (int q = 0; q < SMM_NV; ++q) {
      const int k = kw + q * (SMM_THREADS / 8);
      if (k < kc) v[q] = __ldg(reinterpret_cast<const float4*>(src + (long long)k * ld) + c4);
    }
  }
}
template <bool kKFast>
__device__ __forceinline__ void smm_store_panel(const float4 (&v)[SMM_NV], int kc, float* __restrict__ dst, int t) {
  if (kKFast) {
    const int m = t & 31, k4w = t >> 5;
    const int nk4 = kc >> 2;
#pragma unroll
    for (int q = 0; q < SMM_NV; ++q) {
      const int k4 = k4w + q * (SMM_THREADS / 32);
      if (k4 < nk4) {
        float* d = dst + (k4 * 4) * SMM_P + m;
        d[0] = v[q].x; d[SMM_P] = v[q].y; d[2 * SMM_P] = v[q].z; d[3 * SMM_P] = v[q].w;
      }
    }
  } else {
    const int c4 = t & 7, kw = t >> 3;
#pragma unroll
    for (int q = 0; q < SMM_NV; ++q) {
      const int k = kw + q * (SMM_THREADS / 8);
      if (k < kc) *reinterpret_cast<float4*>(dst + k * SMM_P + c4 * 4) = v[q];
    }
  }
}

// Optional epilogues of a fast-path tile (the fused small-matrix launches):
//   EPI_PACK_T  the tile also goes out as bf16 hi/lo operand images for the tcgen05 apply kernels: image 0 holds
//               src(n,k) = C[k][n] (K3's T^T), image 1 src(n,k) = C[n][k] (K6's T); either may be NULL
//   EPI_PSI     C is a lower-triangular product Z; what is stored is Psi = -(tril(Z) + tril(Z)^T - diag(Z))
//               (both triangles; CTAs of strictly upper tiles do nothing)
enum { EPI_NONE = 0, EPI_PACK = 1, EPI_PSI = 2 };
struct SmmEpi {
  int kind;
  uint8_t* pk0; uint8_t* pk1;      // packed images (wc_pack_* layout), NULL = skip
  long long pk_slot_stride;        // bytes between the images of consecutive b1 batches
  int pkC;                         // channels of the packed matrix
  int row_per_b2;                  // global row of the tile = b2 * row_per_b2 + local row (block-diagonal left factor)
};

__device__ __forceinline__ void smm_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// byte offset of the 16-byte chunk (8 k's) `chunk` of row n in a K-major SWIZZLE_64B image with 32 k's per row
__device__ __forceinline__ uint32_t smm_kmajor32_off(int n, int chunk) {
  const uint32_t off = (uint32_t)n * 64u + (uint32_t)chunk * 16u;
  return off ^ (((off >> 7) & 3u) << 4);
}

template <bool kAK, bool kBK>
__device__ __forceinline__ void smm_tile(const WcSmm& p, const SmmExtra& ex, const SmmEpi& ep, int bx, int by, int z, float* smm_sm) {
  float* As = smm_sm;                       // [k][SMM_P]
  float* Bs = smm_sm + SMM_KC * SMM_P;      // [k][SMM_P]
  if (ep.kind == EPI_PSI && by < bx) return;
  const int b1 = z / p.nb2, b2 = z % p.nb2;
  const float* A = p.A + b1 * p.bA1 + b2 * p.bA2;
  const float* B = p.B + b1 * p.bB1 + b2 * p.bB2;
  float* C = p.C + b1 * p.bC1 + b2 * p.bC2;
  const int i0 = by * SMM_T, j0 = bx * SMM_T;
  const int t = threadIdx.x, kg = t >> 7, tt = t & 127, tx = tt & 15, ty = tt >> 4;
  const long long lda = kAK ? p.rsA : p.csA, ldb = kBK ? p.csB : p.rsB;
  float acc[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = 0.f;
  const int nkc = (p.K + SMM_KC - 1) / SMM_KC;
  for (int it = 0; it < p.nsum * nkc; ++it) {
    const int q = it / nkc, k0 = (it % nkc) * SMM_KC;
    const int kc = min(SMM_KC, p.K - k0);
    const float* Aq = A + (long long)q * p.sA + (kAK ? (long long)i0 * lda + k0 : (long long)k0 * lda + i0);
    const float* Bq = B + (long long)q * p.sB + (kBK ? (long long)j0 * ldb + k0 : (long long)k0 * ldb + j0);
    float4 va[SMM_NV], vb[SMM_NV];
    smm_load_panel<kAK>(Aq, lda, kc, va, t);
    smm_load_panel<kBK>(Bq, ldb, kc, vb, t);
    if (it) __syncthreads();
    smm_store_panel<kAK>(va, kc, As, t);
    smm_store_panel<kBK>(vb, kc, Bs, t);
    __syncthreads();
    // k-group kg takes every 4th k (kc is a multiple of 4)
#pragma unroll 8
    for (int kk = kg; kk < kc; kk += SMM_KG) {
      const float4 a = *reinterpret_cast<const float4*>(As + kk * SMM_P + 4 * ty);
      const float2 b = *reinterpret_cast<const float2*>(Bs + kk * SMM_P + 2 * tx);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
    }
  }
  // reduce the k-groups through shared memory (As is free after the barrier)
  __syncthreads();
  float* red = smm_sm;                      // [SMM_KG - 1][8][128]
  if (kg > 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) red[((kg - 1) * 8 + a * 2 + b) * 128 + tt] = acc[a][b];
  }
  __syncthreads();
  if (kg > 0) return;
#pragma unroll
  for (int g2 = 0; g2 < SMM_KG - 1; ++g2)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) acc[a][b] += red[(g2 * 8 + a * 2 + b) * 128 + tt];
  float alpha = p.alpha;
  if (ex.count_mode) {
    float cnt = ex.count_dev ? ex.count_dev[0] : ex.count_host;
    alpha /= (ex.count_mode == 1 ? cnt : (cnt - 1.f));
  }
  if (ep.kind == EPI_PSI) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int i = i0 + ty * 4 + a, j = j0 + tx * 2 + b;
        if (j > i) continue;
        const float v = -alpha * acc[a][b];
        C[(long long)i * p.rsC + (long long)j * p.csC] = v;
        if (j < i) C[(long long)j * p.rsC + (long long)i * p.csC] = v;
      }
    return;
  }
  float* tile = smm_sm + 4096;              // [32][33], clear of the reduction scratch
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty * 4 + a, j = j0 + tx * 2 + b;
      float* cp = C + (long long)i * p.rsC + (long long)j * p.csC;
      float v = alpha * acc[a][b];
      if (i == j) v *= ex.diag_scale;
      if ((p.tri == 1 && j > i) || (p.tri == 2 && j < i)) v = 0.f;
      else if (p.beta != 0.f) v += p.beta * (*cp);
      *cp = v;
      if (ep.kind == EPI_PACK) tile[(ty * 4 + a) * 33 + tx * 2 + b] = v;
    }
  if (ep.kind != EPI_PACK) return;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int lane32 = tt & 31, ch = tt >> 5;            // one 16-byte chunk (8 k's) per thread and image
  const int gi0 = b2 * ep.row_per_b2 + i0;             // global row of the tile's first row
  const long long b_img = (long long)ep.pkC * 64;      // bytes of one bf16 image of a 32-wide k block
  uint32_t hi[4], lo[4];
  if (ep.pk0) {        // src(n,k) = C[k][n]: k block = the tile's rows, n = its columns
#pragma unroll
    for (int q = 0; q < 4; ++q) smm_split2(tile[(8 * ch + 2 * q) * 33 + lane32], tile[(8 * ch + 2 * q + 1) * 33 + lane32], hi[q], lo[q]);
    uint8_t* dst = ep.pk0 + (long long)b1 * ep.pk_slot_stride + (long long)(gi0 >> 5) * (2 * b_img) + smm_kmajor32_off(j0 + lane32, ch);
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + b_img) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (ep.pk1) {        // src(n,k) = C[n][k]: k block = the tile's columns, n = its rows
#pragma unroll
    for (int q = 0; q < 4; ++q) smm_split2(tile[lane32 * 33 + 8 * ch + 2 * q], tile[lane32 * 33 + 8 * ch + 2 * q + 1], hi[q], lo[q]);
    uint8_t* dst = ep.pk1 + (long long)b1 * ep.pk_slot_stride + (long long)(j0 >> 5) * (2 * b_img) + smm_kmajor32_off(gi0 + lane32, ch);
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + b_img) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <bool kAK, bool kBK>
__global__ void __launch_bounds__(SMM_THREADS) smm_kernel(WcSmm p, SmmExtra ex) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float smm_sm[];
  SmmEpi ep{};
  smm_tile<kAK, kBK>(p, ex, ep, blockIdx.x, blockIdx.y, blockIdx.z, smm_sm);
}

// ---- several independent products (and matrix-vector products) in ONE launch ---------------------------------------
// The small-matrix stages are chains of 256^3-sized products whose cost is the launch, not the math; products that do
// not depend on each other share a launch.
struct SmmJob { WcSmm p; SmmExtra ex; SmmEpi ep; int ak, bk; int gx, gy, gz; int cta0; };
// out[b][i] = scale * sum_q sum_k A_q[b](i,k) * x_q[b][k]   (x == NULL: ones)
struct VecJob {
  const float* A; const float* x; float* out;
  int M, K; long long rsA, csA;
  int nsum; long long sA, sx;
  int nb; long long bA, bx, bo;
  const float* count_dev; float count_host; int count_mode; float alpha;
  int cta0, nctas;
};
constexpr int SMM_MAXJOBS = 3, SMM_MAXVEC = 2;
// lead: CTAs [0, lead) exit at once.  The tile code is ~1500 straight-line instructions per operand-layout variant and a
// launch runs it once per SM: ~2.7 us slower when that SM's instruction cache last held another variant.  CTA i of
// consecutive launches lands on the same SM, so a launch can be shifted onto the SMs that just ran its variant.
struct SmmMulti { SmmJob job[SMM_MAXJOBS]; int njobs; VecJob vec[SMM_MAXVEC]; int nvec; int lead; };

// one CTA = 32 outputs of one batch; 16 k-slices of 32 threads
__device__ __forceinline__ void vec_job_run(const VecJob& v, int cta, float* sm) {
  const int per_b = (v.M + 31) / 32;
  const int b = cta / per_b, i0 = (cta % per_b) * 32;
  const int t = threadIdx.x, io = t & 31, ks = t >> 5;       // 16 slices
  const float* A = v.A + (long long)b * v.bA;
  const float* x = v.x ? v.x + (long long)b * v.bx : nullptr;
  float acc = 0.f;
  if (v.rsA == 1) {
    // consecutive outputs are contiguous: lanes walk i, slices walk k
    const int i = i0 + io;
    if (i < v.M)
      for (int q = 0; q < v.nsum; ++q)
        for (int k = ks; k < v.K; k += 16) acc = fmaf(A[(long long)q * v.sA + (long long)k * v.csA + i], x ? x[(long long)q * v.sx + k] : 1.f, acc);
    sm[ks * 32 + io] = acc;
    __syncthreads();
    if (ks == 0) {
#pragma unroll
      for (int s2 = 1; s2 < 16; ++s2) acc += sm[s2 * 32 + io];
    }
  } else {
    // rows of A are (near) contiguous in k: a warp walks k for two outputs
    for (int h = 0; h < 2; ++h) {
      const int i = i0 + ks * 2 + h;
      float a = 0.f;
      if (i < v.M)
        for (int q = 0; q < v.nsum; ++q)
          for (int k = io; k < v.K; k += 32) a = fmaf(A[(long long)q * v.sA + (long long)i * v.rsA + (long long)k * v.csA], x ? x[(long long)q * v.sx + k] : 1.f, a);
      a = wc_warp_sum(a);
      if (io == 0) sm[ks * 2 + h] = a;
    }
    __syncthreads();
    acc = (t < 32) ? sm[t] : 0.f;
  }
  if (t < 32 && i0 + t < v.M) {
    float scale = v.alpha;
    if (v.count_mode) {
      float cnt = v.count_dev ? v.count_dev[0] : v.count_host;
      scale /= (v.count_mode == 1 ? cnt : (cnt - 1.f));
    }
    v.out[(long long)b * v.bo + i0 + t] = scale * acc;
  }
}

__global__ void __launch_bounds__(SMM_THREADS) smm_multi_kernel(const __grid_constant__ SmmMulti mp) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float smm_sm[];
  const int cta = (int)blockIdx.x - mp.lead;
  if (cta < 0) return;
  for (int v = 0; v < mp.nvec; ++v)
    if (cta >= mp.vec[v].cta0 && cta < mp.vec[v].cta0 + mp.vec[v].nctas) { vec_job_run(mp.vec[v], cta - mp.vec[v].cta0, smm_sm); return; }
  int j = 0;
  for (int q = 1; q < mp.njobs; ++q) if (cta >= mp.job[q].cta0) j = q;
  const SmmJob& jb = mp.job[j];
  const int l = cta - jb.cta0;
  const int bx = l % jb.gx, by = (l / jb.gx) % jb.gy, z = l / (jb.gx * jb.gy);
  if (jb.ak) {
    if (jb.bk) smm_tile<true, true>(jb.p, jb.ex, jb.ep, bx, by, z, smm_sm); else smm_tile<true, false>(jb.p, jb.ex, jb.ep, bx, by, z, smm_sm);
  } else {
    if (jb.bk) smm_tile<false, true>(jb.p, jb.ex, jb.ep, bx, by, z, smm_sm); else smm_tile<false, false>(jb.p, jb.ex, jb.ep, bx, by, z, smm_sm);
  }
}

// ---- generic path: any shape, any strides ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smm_generic_kernel(WcSmm p, SmmExtra ex) {
  wc_pdl_wait();
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int z = blockIdx.z, b1 = z / p.nb2, b2 = z % p.nb2;
  const float* A = p.A + b1 * p.bA1 + b2 * p.bA2;
  const float* B = p.B + b1 * p.bB1 + b2 * p.bB2;
  float* C = p.C + b1 * p.bC1 + b2 * p.bC2;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_kfast = (p.csA == 1) || (p.rsA != 1);   // which index varies fastest across threads when loading
  const bool b_jfast = (p.csB == 1) || (p.rsB != 1);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // loader coordinates of this thread's 4 A and 4 B elements per k-chunk
  int aii[4], akk[4], bjj[4], bkb[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int idx = threadIdx.x + e * 256;
    int lo = idx & 31, hi = idx >> 5;
    aii[e] = a_kfast ? hi : lo; akk[e] = a_kfast ? lo : hi;
    bjj[e] = b_jfast ? lo : hi; bkb[e] = b_jfast ? hi : lo;
  }
  const int nk = (p.K + 31) / 32;
  const int total = p.nsum * nk;
  float ra[4], rb[4];
  auto gload = [&](int it) {
    const int q = it / nk, k0 = (it % nk) * 32;
    const float* Aq = A + (long long)q * p.sA;
    const float* Bq = B + (long long)q * p.sB;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ra[e] = (i0 + aii[e] < p.M && k0 + akk[e] < p.K) ? Aq[(long long)(i0 + aii[e]) * p.rsA + (long long)(k0 + akk[e]) * p.csA] : 0.f;
      rb[e] = (j0 + bjj[e] < p.N && k0 + bkb[e] < p.K) ? Bq[(long long)(k0 + bkb[e]) * p.rsB + (long long)(j0 + bjj[e]) * p.csB] : 0.f;
    }
  };
  if (total > 0) gload(0);
  for (int it = 0; it < total; ++it) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[aii[e]][akk[e]] = ra[e]; Bs[bkb[e]][bjj[e]] = rb[e]; }
    __syncthreads();
    if (it + 1 < total) gload(it + 1);     // next chunk's loads overlap this chunk's math
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      float a0 = As[ty * 2][kk], a1 = As[ty * 2 + 1][kk];
      float c0 = Bs[kk][tx * 2], c1 = Bs[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, c0, acc[0][0]); acc[0][1] = fmaf(a0, c1, acc[0][1]);
      acc[1][0] = fmaf(a1, c0, acc[1][0]); acc[1][1] = fmaf(a1, c1, acc[1][1]);
    }
  }
  float alpha = p.alpha;
  if (ex.count_mode) {
    float cnt = ex.count_dev ? ex.count_dev[0] : ex.count_host;
    alpha /= (ex.count_mode == 1 ? cnt : (cnt - 1.f));
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int i = i0 + ty * 2 + a, j = j0 + tx * 2 + b;
      if (i >= p.M || j >= p.N) continue;
      float* cp = C + (long long)i * p.rsC + (long long)j * p.csC;
      float v = alpha * acc[a][b];
      if (i == j) v *= ex.diag_scale;
      if ((p.tri == 1 && j > i) || (p.tri == 2 && j < i)) v = 0.f;
      else if (p.beta != 0.f) v += p.beta * (*cp);
      *cp = v;
    }
}


template <bool kAK, bool kBK>
int smm_fast_launch(const WcSmm& p, const SmmExtra& ex, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    WC_CUDA(cudaFuncSetAttribute(smm_kernel<kAK, kBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMM_SMEM));
    attr_set = true;
  }
  WC_CUDA(wc_launch(smm_kernel<kAK, kBK>, grid, SMM_THREADS, SMM_SMEM, st, p, ex));
  WC_CHECK_LAUNCH();
  return 0;
}

inline bool mult4(long long v) { return (v & 3) == 0; }

int smm(const WcSmm& p, cudaStream_t st, const float* count_dev = nullptr, float count_host = 0.f,
        int count_mode = 0, float diag_scale = 1.f) {
  if (p.M <= 0 || p.N <= 0 || p.nb1 <= 0 || p.nb2 <= 0) return 0;
  SmmExtra ex{count_dev, count_host, count_mode, diag_scale};
  const bool ak = p.csA == 1, bk = p.rsB == 1;       // k is the contiguous index of the operand
  const bool ai = p.rsA == 1, bj = p.csB == 1;
  const long long lda = ak ? p.rsA : p.csA, ldb = bk ? p.csB : p.rsB;
  const bool fast = (p.M % SMM_T == 0) && (p.N % SMM_T == 0) && (p.K % 4 == 0) && (ak || ai) && (bk || bj) &&
                    mult4(lda) && mult4(ldb) && wc_aligned16(p.A) && wc_aligned16(p.B) &&
                    mult4(p.bA1) && mult4(p.bA2) && mult4(p.bB1) && mult4(p.bB2) && mult4(p.sA) && mult4(p.sB);
  if (fast) {
    dim3 grid(p.N / SMM_T, p.M / SMM_T, p.nb1 * p.nb2);
    if (ak) return bk ? smm_fast_launch<true, true>(p, ex, grid, st) : smm_fast_launch<true, false>(p, ex, grid, st);
    return bk ? smm_fast_launch<false, true>(p, ex, grid, st) : smm_fast_launch<false, false>(p, ex, grid, st);
  }
  dim3 grid(wc_cdiv(p.N, 32), wc_cdiv(p.M, 32), p.nb1 * p.nb2);
  WC_CUDA(wc_launch(smm_generic_kernel, grid, 256, 0, st, p, ex));
  WC_CHECK_LAUNCH();
  return 0;
}

bool smm_is_fast(const WcSmm& p) {
  const bool ak = p.csA == 1, bk = p.rsB == 1, ai = p.rsA == 1, bj = p.csB == 1;
  const long long lda = ak ? p.rsA : p.csA, ldb = bk ? p.csB : p.rsB;
  return p.M > 0 && p.N > 0 && (p.M % SMM_T == 0) && (p.N % SMM_T == 0) && (p.K % 4 == 0) && (ak || ai) && (bk || bj) &&
         mult4(lda) && mult4(ldb) && wc_aligned16(p.A) && wc_aligned16(p.B) && mult4(p.bA1) && mult4(p.bA2) &&
         mult4(p.bB1) && mult4(p.bB2) && mult4(p.sA) && mult4(p.sB);
}

// Builder for one fused launch.  add() refuses (returns false) a product the fast tile cannot do; the caller then
// falls back to separate launches.
struct MultiLaunch {
  SmmMulti m{};
  int ctas = 0;
  bool add(const WcSmm& p, const SmmEpi& ep = SmmEpi{}, const float* count_dev = nullptr, float count_host = 0.f,
           int count_mode = 0, float diag_scale = 1.f) {
    if (m.njobs >= SMM_MAXJOBS || !smm_is_fast(p)) return false;
    if (ep.kind == EPI_PACK && (ep.pkC % 32 || (ep.pk0 && !wc_aligned16(ep.pk0)) || (ep.pk1 && !wc_aligned16(ep.pk1)))) return false;
    SmmJob& j = m.job[m.njobs++];
    j.p = p; j.ex = SmmExtra{count_dev, count_host, count_mode, diag_scale}; j.ep = ep;
    if (wc_debug_flags() & 16384) j.ep.kind = EPI_NONE;     // experiment: what does the packing epilogue cost (wrong results)
    j.ak = p.csA == 1; j.bk = p.rsB == 1;
    j.gx = p.N / SMM_T; j.gy = p.M / SMM_T; j.gz = p.nb1 * p.nb2;
    j.cta0 = ctas;
    ctas += j.gx * j.gy * j.gz;
    return true;
  }
  bool add_vec(const VecJob& v0) {
    if (m.nvec >= SMM_MAXVEC || v0.M <= 0 || v0.nb <= 0) return false;
    VecJob& v = m.vec[m.nvec++];
    v = v0;
    v.cta0 = ctas; v.nctas = ((v.M + 31) / 32) * v.nb;
    ctas += v.nctas;
    return true;
  }
  int launch(cudaStream_t st) {
    if (ctas == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
      WC_CUDA(cudaFuncSetAttribute(smm_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMM_SMEM));
      attr_set = true;
    }
    WC_CUDA(wc_launch(smm_multi_kernel, ctas + m.lead, SMM_THREADS, SMM_SMEM, st, m));
    WC_CHECK_LAUNCH();
    return 0;
  }
};

WcSmm smm_init() {
  WcSmm p{};
  p.nb1 = p.nb2 = 1; p.nsum = 1; p.alpha = 1.f; p.beta = 0.f; p.tri = 0;
  return p;
}

// C[grp] (g x n, row stride ldc) = alpha * op(W[grp]) (g x g) * B[grp rows] ; batched over (outer nb1, groups)
// W is (G,g,g).  transW: use W^T.
WcSmm smm_blockdiag_left(const float* W, bool transW, const float* B, float* C, int g, int G, int ncols, int ld,
                         int nouter, long long outer_stride) {
  WcSmm p = smm_init();
  p.A = W; p.B = B; p.C = C; p.M = g; p.N = ncols; p.K = g;
  p.rsA = transW ? 1 : g; p.csA = transW ? g : 1;
  p.rsB = ld; p.csB = 1; p.rsC = ld; p.csC = 1;
  p.nb1 = nouter; p.nb2 = G;
  p.bA1 = 0; p.bA2 = (long long)g * g;
  p.bB1 = outer_stride; p.bB2 = (long long)g * ld;
  p.bC1 = outer_stride; p.bC2 = (long long)g * ld;
  return p;
}

// (G,g,g) x (G,g,g) -> (G,g,g) with optional transposes
WcSmm smm_gg(const float* A, bool tA, const float* B, bool tB, float* C, int g, int G) {
  WcSmm p = smm_init();
  p.A = A; p.B = B; p.C = C; p.M = p.N = p.K = g;
  p.rsA = tA ? 1 : g; p.csA = tA ? g : 1;
  p.rsB = tB ? 1 : g; p.csB = tB ? g : 1;
  p.rsC = g; p.csC = 1;
  p.nb1 = 1; p.nb2 = G;
  p.bA2 = p.bB2 = p.bC2 = (long long)g * g;
  return p;
}

// =============================================================================================
// K2: prep (scale, EMA, shrink), Cholesky + triangular inverse, Jacobi eigensolver
// =============================================================================================
__global__ void factor_prep_kernel(const float* __restrict__ mean, const float* __restrict__ S,
                                   const float* count_dev, float count_host, int C, int g, float eps, float extra_eps,
                                   float momentum, int training, float* moving_mean, float* moving_cov,
                                   float* __restrict__ Ashr) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)C * g;
  if (idx >= total) return;
  int rem = (int)(idx % ((long long)g * g));
  int r = rem / g, c = rem % g;
  float cov;
  if (training) {
    float cnt = count_dev ? count_dev[0] : count_host;
    // symmetrise on read: the two triangles of S may differ in the last bit (different summation order)
    long long base = idx - rem;
    cov = 0.5f * (S[base + (long long)r * g + c] + S[base + (long long)c * g + r]) / (cnt - 1.f);
    moving_cov[idx] = moving_cov[idx] * momentum + cov * (1.f - momentum);        // conditional_layers.py:687-689
    if (r == c) {
      int ch = (int)(idx / ((long long)g * g)) * g + r;
      moving_mean[ch] = moving_mean[ch] * momentum + mean[ch] * (1.f - momentum);  // :684-686
    }
  } else {
    cov = moving_cov[idx];
  }
  Ashr[idx] = (1.f - eps) * cov + (r == c ? eps + extra_eps : 0.f);                // :691 (+ :671 for ZCA)
}

// One-sided (Hestenes) Jacobi SVD of the Cholesky factor L (A = L L^T = U diag(sigma^2) U^T).
// One CTA per group; Bm is a (G,g,g) COLUMN-major scratch (column j contiguous) in global memory (L1/L2
// resident, only this CTA touches its slice).  A warp owns one column pair per step.
__global__ void __launch_bounds__(1024) jacobi_kernel(const float* __restrict__ L, int g, float* __restrict__ Bm,
                                                      float* __restrict__ U, float* __restrict__ lam, int32_t* info) {
  wc_pdl_wait();
  __shared__ int pairs[2][256];      // round-robin tournament positions (g <= 512)
  __shared__ int rotated;
  const int grp = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
  const int warp = t >> 5, lane = t & 31, nw = nt >> 5;
  const float* Lg = L + (long long)grp * g * g;
  float* B = Bm + (long long)grp * g * g;
  for (int idx = t; idx < g * g; idx += nt) {
    int i = idx / g, j = idx % g;
    B[(long long)j * g + i] = Lg[idx];     // column-major copy
  }
  const int n = (g + 1) & ~1;              // even number of players (one dummy when g is odd)
  const int half = n / 2;
  int sweep = 0;
  for (; sweep < 20; ++sweep) {
    if (t == 0) rotated = 0;
    for (int i = t; i < half; i += nt) { pairs[0][i] = i; pairs[1][i] = n - 1 - i; }
    __syncthreads();
    for (int round = 0; round < n - 1; ++round) {
      for (int pi = warp; pi < half; pi += nw) {
        int p = pairs[0][pi], q = pairs[1][pi];
        if (p >= g || q >= g) continue;
        float* bp = B + (long long)p * g;
        float* bq = B + (long long)q * g;
        float al = 0.f, be = 0.f, ga = 0.f;
        for (int i = lane; i < g; i += 32) { float x = bp[i], y = bq[i]; al = fmaf(x, x, al); be = fmaf(y, y, be); ga = fmaf(x, y, ga); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        if (fabsf(ga) > 1e-6f * sqrtf(al * be) && ga != 0.f) {
          float zeta = (be - al) / (2.f * ga);
          float tt = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
          float c = rsqrtf(1.f + tt * tt), s = c * tt;
          for (int i = lane; i < g; i += 32) {
            float x = bp[i], y = bq[i];
            bp[i] = c * x - s * y;
            bq[i] = s * x + c * y;
          }
          if (lane == 0) rotated = 1;
        }
      }
      __syncthreads();
      // rotate the tournament: position 0 of row 0 is fixed
      int np0 = -1, np1 = -1;
      if (t < half) {
        np0 = (t == 0) ? pairs[0][0] : (t == 1 ? pairs[1][0] : pairs[0][t - 1]);
        np1 = (t == half - 1) ? pairs[0][half - 1] : pairs[1][t + 1];
      }
      __syncthreads();
      if (t < half) { pairs[0][t] = np0; pairs[1][t] = np1; }
      __syncthreads();
    }
    int again = rotated;
    __syncthreads();
    if (!again) break;
  }
  if (t == 0 && info) atomicMax(&info[WC_INFO_JACOBI], sweep + 1);
  // columns of B are U * sigma; lam = sigma^2 (eigenvalue of A).  U written row-major (G,g,g).
  float* Ug = U + (long long)grp * g * g;
  for (int j = warp; j < g; j += nw) {
    const float* bj = B + (long long)j * g;
    float nn = 0.f;
    for (int i = lane; i < g; i += 32) nn = fmaf(bj[i], bj[i], nn);
    nn = wc_warp_sum(nn);
    float inv = rsqrtf(fmaxf(nn, 1e-38f));
    for (int i = lane; i < g; i += 32) Ug[(long long)i * g + j] = bj[i] * inv;
    if (lane == 0) lam[grp * g + j] = nn;
  }
}

// Us = U * diag(lam^power)  (row-major (G,g,g))
__global__ void scale_cols_kernel(const float* __restrict__ U, const float* __restrict__ lam, int C, int g, float power,
                                  float* __restrict__ out) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)C * g) return;
  int grp = (int)(idx / ((long long)g * g));
  int j = (int)(idx % g);
  out[idx] = U[idx] * powf(lam[grp * g + j], power);
}

// Y[i,j] *= -1 / (r_i r_j (r_i + r_j)),  r = sqrt(lam)      (Daleckii-Krein for lam^-1/2)
__global__ void zca_mulF_kernel(float* __restrict__ Y, const float* __restrict__ lam, int C, int g) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)C * g) return;
  int grp = (int)(idx / ((long long)g * g));
  int rem = (int)(idx % ((long long)g * g));
  float ri = sqrtf(lam[grp * g + rem / g]), rj = sqrtf(lam[grp * g + rem % g]);
  Y[idx] *= -1.f / (ri * rj * (ri + rj));
}

// out = alpha * (X + X^T) per group; alpha optionally divided by (count - 1)
__global__ void sym_scale_kernel(const float* __restrict__ X, float alpha, const float* count_dev, float count_host,
                                 int use_count, int C, int g, float* __restrict__ out) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)C * g) return;
  long long base = idx - idx % ((long long)g * g);
  int rem = (int)(idx - base);
  int r = rem / g, c = rem % g;
  if (use_count) alpha /= ((count_dev ? count_dev[0] : count_host) - 1.f);
  out[idx] = alpha * (X[idx] + X[base + (long long)c * g + r]);
}

// =============================================================================================
// K2b: per-class matrices
// =============================================================================================
__global__ void build_wk_kernel(int Cin, int Cout, int K, const float* __restrict__ Wu, const float* __restrict__ bu,
                                const float* __restrict__ Wc, const float* __restrict__ bc, int triangular,
                                const float* __restrict__ cm, const float* __restrict__ D, int E,
                                const float* __restrict__ bf, float* __restrict__ ehat, float* __restrict__ Wk,
                                float* __restrict__ b, int identity) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long per = (long long)Cin * Cout;
  if (idx >= per * K) return;
  int k = (int)(idx / per);
  long long ij = idx % per;
  int i = (int)(ij / Cout), j = (int)(ij % Cout);
  float w = Wu ? Wu[ij] : ((identity && i == j) ? 1.f : 0.f);   // no coloring at all: W = I (generator.py:79-80)
  if (Wc && !(triangular && j < i)) w += Wc[idx];              // matrix_band_part(kernel, 0, -1): upper triangle
  if (cm) {
    float sq = 0.f;
    for (int e = 0; e < E; ++e) { float v = cm[k * E + e]; sq = fmaf(v, v, sq); }
    float rho = rsqrtf(fmaxf(sq, 1e-12f));                     // tf.nn.l2_normalize
    for (int e = 0; e < E; ++e) {
      float eh = cm[k * E + e] * rho;
      w = fmaf(eh, D[(long long)e * per + ij], w);
      if (ij == 0) ehat[k * E + e] = eh;
    }
  }
  Wk[idx] = w;
  if (i == 0 && b) b[(long long)k * Cout + j] = (bu ? bu[j] : 0.f) + (bc ? bc[(long long)k * Cout + j] : 0.f) +
                                                (bf ? bf[(long long)k * Cout + j] : 0.f);
}

// =============================================================================================
// K5a helpers
// =============================================================================================
__global__ void sum_over_k_kernel(const float* __restrict__ X, long long per, int K, float* __restrict__ out) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += X[(long long)k * per + idx];
  out[idx] = s;
}

__global__ void triu_copy_kernel(const float* __restrict__ X, int Cin, int Cout, long long total, float* __restrict__ out) {
  wc_pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long ij = idx % ((long long)Cin * Cout);
  int i = (int)(ij / Cout), j = (int)(ij % Cout);
  out[idx] = (j < i) ? 0.f : X[idx];
}

// de[k,e] = <Q_k, D_e>, one CTA per (k,e)
__global__ void __launch_bounds__(256) class_dot_kernel(const float* __restrict__ Q, const float* __restrict__ D, long long per, int E,
                                                        float* __restrict__ de) {
  wc_pdl_wait();
  __shared__ float part[8];
  int k = blockIdx.x / E, e = blockIdx.x % E;
  const float* q = Q + (long long)k * per;
  const float* d = D + (long long)e * per;
  float s = 0.f;
  for (long long i = threadIdx.x; i < per; i += 256) s = fmaf(q[i], d[i], s);
  s = wc_warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += part[w];
    de[blockIdx.x] = a;
  }
}

// backward of tf.nn.l2_normalize (conditional_layers.py:994): one thread per class
__global__ void l2norm_bwd_kernel(const float* __restrict__ cm, const float* __restrict__ ehat, const float* __restrict__ de,
                                  int K, int E, float* __restrict__ dcm) {
  wc_pdl_wait();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float sq = 0.f, dot = 0.f;
  for (int e = 0; e < E; ++e) { float v = cm[k * E + e]; sq = fmaf(v, v, sq); dot = fmaf(de[k * E + e], ehat[k * E + e], dot); }
  float rho = rsqrtf(fmaxf(sq, 1e-12f));
  bool clamped = !(sq > 1e-12f);
  for (int e = 0; e < E; ++e)
    dcm[k * E + e] = clamped ? rho * de[k * E + e] : rho * (de[k * E + e] - ehat[k * E + e] * dot);
}

size_t gg_bytes(int C, int g) { return wc_align_up((size_t)C * g * 4, 256); }

}  // namespace

int wc_smm_launch(const WcSmm& p, cudaStream_t st) { return smm(p, st); }

// =============================================================================================
// K2 entry points
// =============================================================================================
extern "C" size_t wc_factor_workspace_bytes(int C, int g, int decomposition, int renorm) {
  if (C <= 0 || g <= 0 || C % g) return 0;
  // Ashr, L, (zca: Bm, Us) ; renorm: Amov, Wmov, Lmov-scratch
  size_t n = 2;
  if (decomposition == WC_DECOMP_ZCA) n += 2;
  if (renorm) n += 4;
  return n * gg_bytes(C, g) + wc_align_up(wc_chol_pad_floats(C, g) * 4, 256);
}

namespace {
// factor one (G,g,g) SPD matrix `Ashr` -> Wm (inverse square root), optional sqrt factor, optional U/lam
int factor_matrix(const float* Ashr, int C, int g, int decomposition, float* Wm, float* sqrt_out, float* U, float* lam,
                  float* L /*scratch or output*/, float* Bm, float* Us, float* Xpad, int32_t* info, cudaStream_t st) {
  const int G = C / g;
  long long total = (long long)C * g;
  if (decomposition == WC_DECOMP_CHOLESKY) return wc_chol_inv_launch(Ashr, C, g, sqrt_out, Wm, Xpad, info, st);   // L only on request
  int rcc = wc_chol_inv_launch(Ashr, C, g, L, nullptr, Xpad, info, st);
  if (rcc) return rcc;
  WC_CUDA(wc_launch(jacobi_kernel, G, 1024, 0, st, L, g, Bm, U, lam, info));
  WC_CHECK_LAUNCH();
  WC_CUDA(wc_launch(scale_cols_kernel, wc_cdiv(total, 256), 256, 0, st, U, lam, C, g, -0.25f, Us));
  WC_CHECK_LAUNCH();
  int rc = smm(smm_gg(Us, false, Us, true, Wm, g, G), st);
  if (rc) return rc;
  if (sqrt_out) {
    WC_CUDA(wc_launch(scale_cols_kernel, wc_cdiv(total, 256), 256, 0, st, U, lam, C, g, 0.25f, Us));
    WC_CHECK_LAUNCH();
    rc = smm(smm_gg(Us, false, Us, true, sqrt_out, g, G), st);
    if (rc) return rc;
  }
  return 0;
}
}  // namespace

extern "C" int wc_factor_fwd(const float* mean, const float* S, const float* count_dev, double count_host, int C, int g,
                             float eps, float momentum, int decomposition, int renorm, int training,
                             float* moving_mean, float* moving_cov, float* Weff, float* Wm, float* R, float* U,
                             float* lam, int32_t* info, float* workspace, size_t workspace_bytes, wc_stream_t stream) {
  return wc_factor_fwd_impl(mean, S, count_dev, count_host, C, g, eps, momentum, decomposition, renorm, training, moving_mean,
                            moving_cov, Weff, Wm, R, U, lam, info, workspace, workspace_bytes, 0, (cudaStream_t)stream);
}

// prep_done: the shrunk covariance is already at the start of `workspace` and the moving averages are updated
// (stats_finalize_prep_kernel); only valid for training without renorm
int wc_factor_fwd_impl(const float* mean, const float* S, const float* count_dev, double count_host, int C, int g,
                       float eps, float momentum, int decomposition, int renorm, int training,
                       float* moving_mean, float* moving_cov, float* Weff, float* Wm, float* R, float* U,
                       float* lam, int32_t* info, float* workspace, size_t workspace_bytes, int prep_done, cudaStream_t stream) {
  if (prep_done && (!training || renorm)) return WC_E_ARG;
  if (C <= 0 || g <= 0 || C % g || !moving_cov || !Weff || !workspace) return WC_E_ARG;
  if (training && (!mean || !S || !moving_mean || !Wm)) return WC_E_ARG;
  if (decomposition != WC_DECOMP_CHOLESKY && decomposition != WC_DECOMP_ZCA) return WC_E_ARG;
  if (decomposition == WC_DECOMP_ZCA && training && (!U || !lam)) return WC_E_ARG;
  if (training && renorm && !R) return WC_E_ARG;
  if (g > 256) return WC_E_UNSUPPORTED;   // one CTA holds the whole group matrix in shared memory
  if (workspace_bytes < wc_factor_workspace_bytes(C, g, decomposition, renorm && training)) return WC_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = C / g;
  const size_t gb = gg_bytes(C, g);
  char* ws = reinterpret_cast<char*>(workspace);
  float* Ashr = reinterpret_cast<float*>(ws);
  float* L = reinterpret_cast<float*>(ws + gb);
  float *Bm = nullptr, *Us = nullptr;
  size_t off = 2 * gb;
  const bool zca = decomposition == WC_DECOMP_ZCA;
  if (zca) { Bm = reinterpret_cast<float*>(ws + off); Us = reinterpret_cast<float*>(ws + off + gb); off += 2 * gb; }
  float* Xpad = reinterpret_cast<float*>(ws + ((zca ? 4 : 2) + ((renorm && training) ? 4 : 0)) * gb);
  const float extra = zca ? eps : 0.f;   // conditional_layers.py:671 adds eps*I again on top of :691
  long long total = (long long)C * g;
  int rc;
  if (!training) {
    // test phase (:702-704): factor the moving covariance; U/lam scratch if the caller gave none
    WC_CUDA(wc_launch(factor_prep_kernel, wc_cdiv(total, 256), 256, 0, st, nullptr, nullptr, nullptr, 0.f, C, g, eps, extra, momentum, 0,
                                                            moving_mean, moving_cov, Ashr));
    WC_CHECK_LAUNCH();
    if (zca && (!U || !lam)) return WC_E_ARG;   // eigenvector / eigenvalue scratch is required in every phase
    return factor_matrix(Ashr, C, g, decomposition, Weff, nullptr, U, lam, L, Bm, Us, Xpad, info, st);
  }
  float *Amov = nullptr, *Wmov = nullptr, *Lb = nullptr, *Umov = nullptr;
  if (renorm) {
    Amov = reinterpret_cast<float*>(ws + off); Wmov = reinterpret_cast<float*>(ws + off + gb);
    Lb = reinterpret_cast<float*>(ws + off + 2 * gb); Umov = reinterpret_cast<float*>(ws + off + 3 * gb);
    // moving covariance as it is BEFORE this step's update (:695-696)
    WC_CUDA(wc_launch(factor_prep_kernel, wc_cdiv(total, 256), 256, 0, st, nullptr, nullptr, nullptr, 0.f, C, g, eps, extra, momentum, 0,
                                                            moving_mean, moving_cov, Amov));
    WC_CHECK_LAUNCH();
    // lam scratch for the moving factor: reuse R's first C floats (R is written afterwards)
    rc = factor_matrix(Amov, C, g, decomposition, Wmov, nullptr, Umov, R, L, Bm, Us, Xpad, info, st);
    if (rc) return rc;
  }
  if (!prep_done) {
    WC_CUDA(wc_launch(factor_prep_kernel, wc_cdiv(total, 256), 256, 0, st, mean, S, count_dev, (float)count_host, C, g, eps, extra, momentum, 1,
                                                            moving_mean, moving_cov, Ashr));
    WC_CHECK_LAUNCH();
  }
  rc = factor_matrix(Ashr, C, g, decomposition, Wm, renorm ? Lb : nullptr, U, lam, L, Bm, Us, Xpad, info, st);
  if (rc) return rc;
  if (renorm) {
    rc = smm(smm_gg(Wmov, false, Lb, false, R, g, G), st);           // R = l_mov_inverse . stop_gradient(l)   (:697-698)
    if (rc) return rc;
    rc = smm(smm_gg(R, false, Wm, false, Weff, g, G), st);           // Weff = R . l_inv
    if (rc) return rc;
  } else if (Weff != Wm) {
    WC_CUDA(cudaMemcpyAsync(Weff, Wm, (size_t)total * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// =============================================================================================
// K2b entry point
// =============================================================================================
int wc_compose_fwd_impl(const float* Weff, int Cin, int Cout, int g, int K, const float* Wu, const float* bu,
                        const float* Wc, const float* bc, int triangular, const float* class_matrix,
                        const float* D, int E, const float* bf, float* ehat, float* Wk, float* T, float* b,
                        void* Tp_fwd, void* Tp_bwd, const float* Wk_ready, cudaStream_t st) {
  // Wk_ready (nullable): the per-class matrix already exists (K == 1 with only the unconditional kernel: it IS Wu, and the
  // bias is bu); nothing is built and `Wk` / `b` are not written
  if (Cin <= 0 || Cout <= 0 || K <= 0 || (!Wk && !Wk_ready) || !T) return WC_E_ARG;
  if (Weff && (g <= 0 || Cin % g)) return WC_E_ARG;
  if (class_matrix && (!D || E <= 0 || !ehat)) return WC_E_ARG;
  long long per = (long long)Cin * Cout;
  if (!Wk_ready) {
    WC_CUDA(wc_launch(build_wk_kernel, wc_cdiv(per * K, 256), 256, 0, st, Cin, Cout, K, Wu, bu, Wc, bc, triangular, class_matrix, D, E, bf,
                                                           ehat, Wk, b, (!Wu && !Wc && !D) ? 1 : 0));
    WC_CHECK_LAUNCH();
  }
  const float* Wsrc = Wk_ready ? Wk_ready : Wk;
  int rc = 0;
  bool packed = false;
  if (!Weff) {
    if (T != Wsrc) WC_CUDA(cudaMemcpyAsync(T, Wsrc, (size_t)per * K * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    const WcSmm p = smm_blockdiag_left(Weff, true, Wsrc, T, g, Cin / g, Cout, Cout, K, per);
    MultiLaunch ml;
    SmmEpi ep{};
    if (Tp_fwd || Tp_bwd) {   // T leaves the product tile as fp32 and as the bf16 operand images of K3 / K6
      ep.kind = EPI_PACK; ep.pk0 = (uint8_t*)Tp_fwd; ep.pk1 = (uint8_t*)Tp_bwd;
      ep.pk_slot_stride = (long long)Cin * Cout * 4; ep.pkC = Cin; ep.row_per_b2 = g;
    }
    if (Cin == Cout && ml.add(p, ep)) {
      rc = ml.launch(st);
      packed = true;
    } else {
      rc = smm(p, st);
    }
    if (rc) return rc;
  }
  if (!packed) {
    if (Tp_fwd) { rc = wc_pack_fwd(T, Cin, K, Tp_fwd, (wc_stream_t)st); if (rc) return rc; }
    if (Tp_bwd) { rc = wc_pack_bwd(T, nullptr, Cin, Weff ? g : Cin, K, Tp_bwd, (wc_stream_t)st); if (rc) return rc; }
  }
  return 0;
}

extern "C" int wc_compose_fwd(const float* Weff, int Cin, int Cout, int g, int K, const float* Wu, const float* bu,
                              const float* Wc, const float* bc, int triangular, const float* class_matrix,
                              const float* D, int E, const float* bf, float* ehat, float* Wk, float* T, float* b,
                              wc_stream_t stream) {
  return wc_compose_fwd_impl(Weff, Cin, Cout, g, K, Wu, bu, Wc, bc, triangular, class_matrix, D, E, bf, ehat, Wk, T, b,
                             nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// =============================================================================================
// K5a entry points
// =============================================================================================
extern "C" size_t wc_color_bwd_workspace_bytes(int Cin, int Cout, int g, int K, int E) {
  if (Cin <= 0 || Cout <= 0 || K <= 0) return 0;
  return wc_align_up((size_t)K * Cin * Cout * 4, 256) + wc_align_up((size_t)K * (E > 0 ? E : 1) * 4, 256);
}

extern "C" int wc_color_bwd(const float* P, const float* s, const float* Weff, const float* Wk, int Cin, int Cout, int g,
                            int K, float* dWu, float* dbu, float* dWc, float* dbc, int triangular,
                            const float* class_matrix, const float* ehat, const float* D, int E, float* dD,
                            float* dclass_matrix, float* dbf, float* dWeff, float* gbar, float* workspace,
                            size_t workspace_bytes, wc_stream_t stream) {
  if (!P || !s || !Wk || Cin <= 0 || Cout <= 0 || K <= 0) return WC_E_ARG;
  if (Weff && (g <= 0 || Cin % g)) return WC_E_ARG;
  if ((dD || dclass_matrix) && (!class_matrix || !ehat || !D || E <= 0)) return WC_E_ARG;
  if (!workspace || workspace_bytes < wc_color_bwd_workspace_bytes(Cin, Cout, g, K, E)) return WC_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long per = (long long)Cin * Cout;
  float* Qws = workspace;
  float* de = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + wc_align_up((size_t)K * per * 4, 256));
  int rc;
  const bool need_q = dWu || dWc || dD || dclass_matrix;
  const float* Q = P;
  // The products that only need P, s and the saved matrices share one launch: Q_k = Weff P_k (= x_hat^T dy per class),
  // dWeff = sum_k Wk P_k^T, gbar = sum_k Wk s_k, dbu = sum_k s_k.  Whatever the fast tile cannot take runs separately.
  MultiLaunch ml;
  bool q_done = false, dweff_done = false, gbar_done = false, dbu_done = false;
  float* Qdst = Qws;
  if (dWc && !triangular) Qdst = dWc;
  else if (K == 1 && dWu) Qdst = dWu;
  WcSmm pq = smm_init(), pw = smm_init();
  if (need_q && Weff) {
    pq = smm_blockdiag_left(Weff, false, P, Qdst, g, Cin / g, Cout, Cout, K, per);
    q_done = ml.add(pq);
    Q = Qdst;
  }
  if (dWeff) {
    if (!Weff) return WC_E_ARG;
    // dWeff[grp] = sum_k Wk[k][grp rows,:] . P[k][grp rows,:]^T
    pw.A = Wk; pw.B = P; pw.C = dWeff; pw.M = g; pw.N = g; pw.K = Cout;
    pw.rsA = Cout; pw.csA = 1; pw.rsB = 1; pw.csB = Cout; pw.rsC = g; pw.csC = 1;
    pw.nb1 = 1; pw.nb2 = Cin / g;
    pw.bA2 = (long long)g * Cout; pw.bB2 = (long long)g * Cout; pw.bC2 = (long long)g * g;
    pw.nsum = K; pw.sA = per; pw.sB = per;
    dweff_done = ml.add(pw);
  }
  if (gbar) {
    VecJob v{};
    v.A = Wk; v.x = s; v.out = gbar; v.M = Cin; v.K = Cout; v.rsA = Cout; v.csA = 1;
    v.nsum = K; v.sA = per; v.sx = Cout; v.nb = 1; v.alpha = 1.f;
    gbar_done = ml.add_vec(v);
  }
  if (dbu) {
    VecJob v{};
    v.A = s; v.x = nullptr; v.out = dbu; v.M = Cout; v.K = K; v.rsA = 1; v.csA = Cout;
    v.nsum = 1; v.nb = 1; v.alpha = 1.f;
    dbu_done = ml.add_vec(v);
  }
  rc = ml.launch(st);
  if (rc) return rc;
  if (need_q && Weff && !q_done) { rc = smm(pq, st); if (rc) return rc; }
  if (dWu && Q != dWu) {
    if (K == 1) WC_CUDA(cudaMemcpyAsync(dWu, Q, (size_t)per * 4, cudaMemcpyDeviceToDevice, st));
    else { WC_CUDA(wc_launch(sum_over_k_kernel, wc_cdiv(per, 256), 256, 0, st, Q, per, K, dWu)); WC_CHECK_LAUNCH(); }
  }
  if (dbu && !dbu_done) { WC_CUDA(wc_launch(sum_over_k_kernel, wc_cdiv(Cout, 256), 256, 0, st, s, Cout, K, dbu)); WC_CHECK_LAUNCH(); }
  if (dWc && Q != dWc) {
    if (triangular) { WC_CUDA(wc_launch(triu_copy_kernel, wc_cdiv(per * K, 256), 256, 0, st, Q, Cin, Cout, per * K, dWc)); WC_CHECK_LAUNCH(); }
    else WC_CUDA(cudaMemcpyAsync(dWc, Q, (size_t)per * K * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (dbc) WC_CUDA(cudaMemcpyAsync(dbc, s, (size_t)K * Cout * 4, cudaMemcpyDeviceToDevice, st));
  if (dbf) WC_CUDA(cudaMemcpyAsync(dbf, s, (size_t)K * Cout * 4, cudaMemcpyDeviceToDevice, st));
  if (dD) {
    // dD[e] = sum_k ehat[k,e] Q_k  : (E x K) . (K x per)
    WcSmm p = smm_init();
    p.A = ehat; p.B = Q; p.C = dD; p.M = E; p.N = (int)per; p.K = K;
    p.rsA = 1; p.csA = E; p.rsB = per; p.csB = 1; p.rsC = per; p.csC = 1;
    rc = smm(p, st);
    if (rc) return rc;
  }
  if (dclass_matrix) {
    WC_CUDA(wc_launch(class_dot_kernel, K * E, 256, 0, st, Q, D, per, E, de));
    WC_CHECK_LAUNCH();
    WC_CUDA(wc_launch(l2norm_bwd_kernel, wc_cdiv(K, 128), 128, 0, st, class_matrix, ehat, de, K, E, dclass_matrix));
    WC_CHECK_LAUNCH();
  }
  if (dWeff && !dweff_done) { rc = smm(pw, st); if (rc) return rc; }
  if (gbar && !gbar_done) {
    // gbar = sum_k Wk[k] s[k]
    WcSmm p = smm_init();
    p.A = Wk; p.B = s; p.C = gbar; p.M = Cin; p.N = 1; p.K = Cout;
    p.rsA = Cout; p.csA = 1; p.rsB = 1; p.csB = 0; p.rsC = 1; p.csC = 1;
    p.nsum = K; p.sA = per; p.sB = Cout;
    rc = smm(p, st);
    if (rc) return rc;
  }
  return 0;
}

// =============================================================================================
// K5b entry points
// =============================================================================================
extern "C" size_t wc_factor_bwd_workspace_bytes(int C, int g, int decomposition, int renorm) {
  if (C <= 0 || g <= 0 || C % g) return 0;
  return 3 * gg_bytes(C, g);
}

int wc_factor_bwd_impl(const float* dWeff, const float* gbar, const float* Weff, const float* Wm, const float* R,
                       const float* U, const float* lam, const float* count_dev, double count_host, int C, int g,
                       float eps, int decomposition, int renorm, float* A, float* cvec, float* workspace,
                       size_t workspace_bytes, void* Tp_bwd, int pack_K, cudaStream_t st) {
  // Tp_bwd (nullable): K6's packed operand buffer [pack_K images of T | image of A]; A's image is refreshed here
  if (!dWeff || !gbar || !Weff || !Wm || !A || !cvec || C <= 0 || g <= 0 || C % g) return WC_E_ARG;
  if (renorm && !R) return WC_E_ARG;
  if (decomposition == WC_DECOMP_ZCA && (!U || !lam)) return WC_E_ARG;
  if (!workspace || workspace_bytes < wc_factor_bwd_workspace_bytes(C, g, decomposition, renorm)) return WC_E_WORKSPACE;
  const int G = C / g;
  const size_t gb = gg_bytes(C, g);
  char* ws = reinterpret_cast<char*>(workspace);
  float* W0 = reinterpret_cast<float*>(ws);
  float* W1 = reinterpret_cast<float*>(ws + gb);
  float* W2 = reinterpret_cast<float*>(ws + 2 * gb);
  const float ch = (float)count_host;
  long long total = (long long)C * g;
  int rc;
  const float* dWm = dWeff;
  if (renorm) {  // Weff = R Wm with R constant  =>  dWm = R^T dWeff
    rc = smm(smm_gg(R, true, dWeff, false, W2, g, G), st);
    if (rc) return rc;
    dWm = W2;
  }
  // cvec = Weff^T gbar / count
  VecJob vc{};
  vc.A = Weff; vc.x = gbar; vc.out = cvec; vc.M = g; vc.K = g; vc.rsA = 1; vc.csA = g;
  vc.nsum = 1; vc.nb = G; vc.bA = (long long)g * g; vc.bx = g; vc.bo = g;
  vc.count_dev = count_dev; vc.count_host = ch; vc.count_mode = 1; vc.alpha = 1.f;
  bool cvec_done = false, packed = false;
  if (decomposition == WC_DECOMP_CHOLESKY) {
    // Z = dWm Wm^T;  Phi = -(tril(Z) - diag(Z)/2);  dSigma' = sym(Wm^T Phi Wm)
    const WcSmm pz = smm_gg(dWm, false, Wm, true, W0, g, G);
    // Y = Psi Wm is stored TRANSPOSED (W1 = Y^T) and A = Wm^T Y computed as (Y^T Wm)^T with transposed output strides:
    // the same products in the same order, but both launches use the row-panel / row-panel operand variant of M1's
    // first job, whose code is still in the instruction cache of the SMs they run on (experiment bit 65536: read the
    // symmetric Psi transposed instead, the k-major variant)
    WcSmm py = smm_gg(W0, (wc_debug_flags() & 65536) != 0, Wm, false, W1, g, G);
    py.rsC = 1; py.csC = g;
    WcSmm pa = smm_gg(W1, false, Wm, false, A, g, G);
    pa.rsC = 1; pa.csC = g;
    if (smm_is_fast(pz) && smm_is_fast(py) && smm_is_fast(pa)) {
      // three launches: Psi = Phi + Phi^T straight from the lower tiles of Z (+ cvec);  Y = Psi Wm;
      // A = (1-eps)/(count-1) Wm^T Y, which also leaves as the bf16 operand image of K6
      MultiLaunch m1, m2, m3;
      SmmEpi e1{}; e1.kind = EPI_PSI;
      m1.add(pz, e1);
      cvec_done = m1.add_vec(vc);
      // (experiment bit 131072: shift this launch by one job's worth of CTAs, onto the SMs that ran M1's second job --
      // the same operand variant; measured: no gain, this launch is not bound by instruction fetch)
      if ((wc_debug_flags() & 131072) && m1.ctas + (g / SMM_T) * (g / SMM_T) * G <= WC_NUM_SMS) m1.m.lead = (g / SMM_T) * (g / SMM_T) * G;
      rc = m1.launch(st); if (rc) return rc;
      m2.add(py);
      rc = m2.launch(st); if (rc) return rc;
      pa.alpha = 1.f - eps;
      SmmEpi e3{};
      if (Tp_bwd && g == C) {   // the tile holds A^T: image src(n,k) = A[k][n] = tile[n][k], the second packing form
        e3.kind = EPI_PACK; e3.pk1 = (uint8_t*)Tp_bwd + (size_t)pack_K * C * C * 4; e3.pkC = C;
        packed = true;
      }
      if (!m3.add(pa, e3, count_dev, ch, 2, 1.f)) return WC_E_UNSUPPORTED;
      rc = m3.launch(st); if (rc) return rc;
    } else {
      WcSmm p = pz;
      p.alpha = -1.f; p.tri = 1;
      rc = smm(p, st, nullptr, 0.f, 0, 0.5f);
      if (rc) return rc;
      rc = smm(smm_gg(W0, false, Wm, false, W1, g, G), st);     // Y = Phi Wm
      if (rc) return rc;
      rc = smm(smm_gg(Wm, true, W1, false, W0, g, G), st);      // Qs = Wm^T Y
      if (rc) return rc;
      // A = (1-eps) (Qs + Qs^T) / (count-1)
      WC_CUDA(wc_launch(sym_scale_kernel, wc_cdiv(total, 256), 256, 0, st, W0, 1.f - eps, count_dev, ch, 1, C, g, A));
      WC_CHECK_LAUNCH();
    }
  } else {
    WC_CUDA(wc_launch(sym_scale_kernel, wc_cdiv(total, 256), 256, 0, st, dWm, 0.5f, nullptr, 0.f, 0, C, g, W0));   // sym(dWm)
    WC_CHECK_LAUNCH();
    rc = smm(smm_gg(U, true, W0, false, W1, g, G), st);       // U^T sym
    if (rc) return rc;
    rc = smm(smm_gg(W1, false, U, false, W0, g, G), st);      // (U^T sym) U
    if (rc) return rc;
    WC_CUDA(wc_launch(zca_mulF_kernel, wc_cdiv(total, 256), 256, 0, st, W0, lam, C, g));
    WC_CHECK_LAUNCH();
    rc = smm(smm_gg(U, false, W0, false, W1, g, G), st);      // U (.)
    if (rc) return rc;
    WcSmm p = smm_gg(W1, false, U, true, A, g, G);            // (.) U^T, scaled: A = 2(1-eps) dSigma / (count-1)
    p.alpha = 2.f * (1.f - eps);
    rc = smm(p, st, count_dev, ch, 2, 1.f);
    if (rc) return rc;
  }
  if (!cvec_done) {
    MultiLaunch mv;
    mv.add_vec(vc);
    rc = mv.launch(st);
    if (rc) return rc;
  }
  if (Tp_bwd && !packed) {
    rc = wc_pack_bwd(nullptr, A, C, g, pack_K, Tp_bwd, (wc_stream_t)st);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int wc_factor_bwd(const float* dWeff, const float* gbar, const float* Weff, const float* Wm, const float* R,
                             const float* U, const float* lam, const float* count_dev, double count_host, int C, int g,
                             float eps, int decomposition, int renorm, float* A, float* cvec, float* workspace,
                             size_t workspace_bytes, wc_stream_t stream) {
  return wc_factor_bwd_impl(dWeff, gbar, Weff, Wm, R, U, lam, count_dev, count_host, C, g, eps, decomposition, renorm, A, cvec,
                            workspace, workspace_bytes, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int wc_version(void) { return WC_B200_VERSION; }

#include <atomic>
static std::atomic<long long> g_launches{0};
void wc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long wc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- debug-only instrumentation ------------------------------------------------------------------------
static int g_dbg_flags = 0;
static unsigned long long* g_dbg_buf = nullptr;
int wc_debug_flags() { return g_dbg_flags; }
int wc_l2_mode() {
  static const int mode = [] { const char* e = getenv("WC_B200_L2"); return e ? atoi(e) : WC_L2_DEFAULT; }();
  return mode;
}
bool wc_pdl_enabled() {
  static const bool env_off = getenv("WC_B200_NO_PDL") != nullptr;    // A/B measurements
  return !env_off && !(g_dbg_flags & 8192);
}
unsigned long long* wc_debug_buf() { return (g_dbg_flags & 1) ? g_dbg_buf : nullptr; }
extern "C" int wc_debug_set(int flags) {
  g_dbg_flags = flags;
  if ((flags & 1) && !g_dbg_buf) {
    cudaError_t e = cudaMalloc(&g_dbg_buf, sizeof(unsigned long long) * WC_DBG_BLOCKS * WC_DBG_SLOTS);
    if (e != cudaSuccess) return (int)e;
  }
  if (g_dbg_buf) cudaMemset(g_dbg_buf, 0, sizeof(unsigned long long) * WC_DBG_BLOCKS * WC_DBG_SLOTS);
  return 0;
}
extern "C" int wc_debug_marks(unsigned long long* host_out) {
  if (!g_dbg_buf) return -1;
  return (int)cudaMemcpy(host_out, g_dbg_buf, sizeof(unsigned long long) * WC_DBG_BLOCKS * WC_DBG_SLOTS, cudaMemcpyDeviceToHost);
}
