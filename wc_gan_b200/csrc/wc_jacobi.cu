// K2 (ZCA): eigen-decomposition of the shrunk covariance by one-sided (Hestenes) Jacobi on its Cholesky factor,
//   Sigma'' = L L^T = U diag(sigma^2) U^T    (columns of B = L V converge to U sigma),
// replacing the reference's host SVD (tf.svd pinned to /cpu:0, gan/conditional_layers.py:668-676).
//
// Two kernels, both with the working columns in SHARED memory (the round-1 kernel kept the matrix in global memory and ran
// 4 column pairs per warp one after the other: 0.65 ms at g = 64, 15.7 ms at g = 256):
//
//   jacobi_smem_kernel   g <= 64: one CTA per group holds the whole g x g matrix; every round of the round-robin
//                        tournament rotates its g/2 disjoint column pairs at once, one warp per pair.
//   jacobi_q8_kernel     g in {16, 32, 64}: the same with eight lanes per pair (four pairs per warp), see below.
//   jacobi_block_kernel  g >= 128 (multiple of 32): the columns are cut into blocks of 16; a block-level round-robin gives
//                        nb/2 independent block pairs per block round, one CTA each (g = 256: 8 CTAs per group matrix).  A
//                        CTA (4 warps, eight lanes per column pair) loads its two blocks (2 x 16 columns x g) into shared
//                        memory, rotates every cross pair (in the
//                        first block round of a sweep also the pairs inside each block), writes the blocks back and meets
//                        the other CTAs of its group at a barrier in L2 (release-add / acquire-spin on a per-group counter).
//                        A sweep is nb-1 block rounds; the sweep loop ends when no CTA of the group rotated anything.
//
// A pair (p, q) is rotated when |b_p . b_q| > 1e-6 |b_p| |b_q|; the sweep loop ends after the first sweep in which no
// pair was further than 1e-4 from orthogonal (JSTOP2): every pair above 1e-6 was still rotated in that sweep, and what a
// sweep leaves behind is far below what it met (measured on the covariances of the parity tests and on log-uniform /
// rank-deficient spectra: a sweep that meets 2e-5 ... 4e-4 leaves 1e-6; the inverse square root is the same to all printed
// digits as with the former rule "stop after a sweep without rotations", which cost one to two more sweeps of ten).
// Sweeps are capped at 20.  info[WC_INFO_JACOBI] records the number of sweeps used.
#include "wc_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int JB = 16;                  // columns per block (block kernel)
constexpr int JBLK_THREADS = 128;       // 4 warps x 4 column pairs (eight lanes each): the 16 pairs of a sub-round
constexpr int JMAX_SWEEPS = 20;
constexpr float JROT2 = 1e-12f;         // rotate a pair when cos^2 > JROT2
constexpr float JSTOP2 = 1e-8f;         // another sweep is needed when some pair had cos^2 > JSTOP2

// round-robin tournament of n (even) players: partner of position `i` (0 <= i < n/2) in round r (0 <= r < n-1)
__device__ __forceinline__ void rr_pair(int n, int r, int i, int& p, int& q) {
  const int m = n - 1;              // r, i < m: one conditional correction replaces the modulo
  if (i == 0) { p = m; q = r; }
  else { p = r + i; if (p >= m) p -= m; q = r - i; if (q < 0) q += m; }
}

// rotate columns bp, bq (length len, lane-strided) of one pair; returns true if the pair was far enough from orthogonal
// to ask for another sweep
template <int EPL>     // elements per lane (len <= 32 * EPL)
__device__ __forceinline__ bool rotate_pair(float* bp, float* bq, int len, int lane) {
  float x[EPL], y[EPL];
  float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = lane + 32 * e;
    x[e] = i < len ? bp[i] : 0.f;
    y[e] = i < len ? bq[i] : 0.f;
    al = fmaf(x[e], x[e], al); be = fmaf(y[e], y[e], be); ga = fmaf(x[e], y[e], ga);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    al += __shfl_xor_sync(FULL, al, o);
    be += __shfl_xor_sync(FULL, be, o);
    ga += __shfl_xor_sync(FULL, ga, o);
  }
  // |ga| > 1e-6 sqrt(al be)  <=>  ga^2 > 1e-12 al be  (no square root on the decision path)
  if (!(ga * ga > JROT2 * al * be)) return false;
  const bool more = ga * ga > JSTOP2 * al * be;
  // tan of the rotation angle: the smaller root of t^2 + 2 zeta t - 1 = 0; fast reciprocal / rsqrt are accurate to ~1 ulp,
  // and an error in (c, s) only makes the rotation slightly less effective (c^2 + s^2 = 1 holds by construction of s = c t)
  const float zeta = __fdividef(be - al, 2.f * ga);
  const float az = fabsf(zeta);
  const float tt = copysignf(__frcp_rn(az + __fsqrt_rn(fmaf(zeta, zeta, 1.f))), zeta);
  const float c = rsqrtf(fmaf(tt, tt, 1.f)), s = c * tt;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = lane + 32 * e;
    if (i < len) { bp[i] = c * x[e] - s * y[e]; bq[i] = s * x[e] + c * y[e]; }
  }
  return more;
}

// the same for len == 32 * EPL with EPL a multiple of 4: a lane owns elements 4*lane .. +3 of every 128 (128-bit accesses)
template <int EPL>
__device__ __forceinline__ bool rotate_pair_v4(float* bp, float* bq, int lane) {
  static_assert(EPL % 4 == 0, "vector width");
  float4 x[EPL / 4], y[EPL / 4];
  float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
  for (int e = 0; e < EPL / 4; ++e) {
    x[e] = *reinterpret_cast<const float4*>(bp + 4 * lane + 128 * e);
    y[e] = *reinterpret_cast<const float4*>(bq + 4 * lane + 128 * e);
    al = fmaf(x[e].x, x[e].x, al); al = fmaf(x[e].y, x[e].y, al); al = fmaf(x[e].z, x[e].z, al); al = fmaf(x[e].w, x[e].w, al);
    be = fmaf(y[e].x, y[e].x, be); be = fmaf(y[e].y, y[e].y, be); be = fmaf(y[e].z, y[e].z, be); be = fmaf(y[e].w, y[e].w, be);
    ga = fmaf(x[e].x, y[e].x, ga); ga = fmaf(x[e].y, y[e].y, ga); ga = fmaf(x[e].z, y[e].z, ga); ga = fmaf(x[e].w, y[e].w, ga);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    al += __shfl_xor_sync(FULL, al, o);
    be += __shfl_xor_sync(FULL, be, o);
    ga += __shfl_xor_sync(FULL, ga, o);
  }
  if (!(ga * ga > JROT2 * al * be)) return false;
  const bool more = ga * ga > JSTOP2 * al * be;
  const float zeta = __fdividef(be - al, 2.f * ga);
  const float tt = copysignf(__frcp_rn(fabsf(zeta) + __fsqrt_rn(fmaf(zeta, zeta, 1.f))), zeta);
  const float c = rsqrtf(fmaf(tt, tt, 1.f)), s = c * tt;
#pragma unroll
  for (int e = 0; e < EPL / 4; ++e) {
    *reinterpret_cast<float4*>(bp + 4 * lane + 128 * e) = make_float4(c * x[e].x - s * y[e].x, c * x[e].y - s * y[e].y, c * x[e].z - s * y[e].z, c * x[e].w - s * y[e].w);
    *reinterpret_cast<float4*>(bq + 4 * lane + 128 * e) = make_float4(s * x[e].x + c * y[e].x, s * x[e].y + c * y[e].y, s * x[e].z + c * y[e].z, s * x[e].w + c * y[e].w);
  }
  return more;
}
template <int EPL>
__device__ __forceinline__ bool rotate_pair_blk(float* bp, float* bq, int len, int lane) {
  if constexpr (EPL % 4 == 0) return rotate_pair_v4<EPL>(bp, bq, lane);
  else return rotate_pair<EPL>(bp, bq, len, lane);
}
// EIGHT lanes per pair (columns of length 32 * NV, contiguous): lane l8 of the pair owns elements 4 l8 + 32 v + 0..3.  A
// quarter-warp reads 128 contiguous bytes of one column per access, a warp rotates four pairs at once, and the three dot
// products take a 3-step reduction among the 8 lanes (the whole-warp version: 5 steps, and 4x the warps per sub-round).
template <int NV>
__device__ __forceinline__ bool rotate_pair_l8(float* bp, float* bq, int l8) {
  float4 x[NV], y[NV];
  float al[2] = {0.f, 0.f}, be[2] = {0.f, 0.f}, ga[2] = {0.f, 0.f};       // two partial sums each: shorter FMA chains
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    x[v] = *reinterpret_cast<const float4*>(bp + 4 * l8 + 32 * v);
    y[v] = *reinterpret_cast<const float4*>(bq + 4 * l8 + 32 * v);
    float& a = al[v & 1]; float& b = be[v & 1]; float& c = ga[v & 1];
    a = fmaf(x[v].x, x[v].x, a); a = fmaf(x[v].y, x[v].y, a); a = fmaf(x[v].z, x[v].z, a); a = fmaf(x[v].w, x[v].w, a);
    b = fmaf(y[v].x, y[v].x, b); b = fmaf(y[v].y, y[v].y, b); b = fmaf(y[v].z, y[v].z, b); b = fmaf(y[v].w, y[v].w, b);
    c = fmaf(x[v].x, y[v].x, c); c = fmaf(x[v].y, y[v].y, c); c = fmaf(x[v].z, y[v].z, c); c = fmaf(x[v].w, y[v].w, c);
  }
  float sa = al[0] + al[1], sb = be[0] + be[1], sg = ga[0] + ga[1];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(FULL, sa, o);
    sb += __shfl_xor_sync(FULL, sb, o);
    sg += __shfl_xor_sync(FULL, sg, o);
  }
  if (!(sg * sg > JROT2 * sa * sb)) return false;
  const bool more = sg * sg > JSTOP2 * sa * sb;
  const float zeta = __fdividef(sb - sa, 2.f * sg);
  const float tt = copysignf(__frcp_rn(fabsf(zeta) + __fsqrt_rn(fmaf(zeta, zeta, 1.f))), zeta);
  const float c = rsqrtf(fmaf(tt, tt, 1.f)), s = c * tt;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    *reinterpret_cast<float4*>(bp + 4 * l8 + 32 * v) = make_float4(c * x[v].x - s * y[v].x, c * x[v].y - s * y[v].y, c * x[v].z - s * y[v].z, c * x[v].w - s * y[v].w);
    *reinterpret_cast<float4*>(bq + 4 * l8 + 32 * v) = make_float4(s * x[v].x + c * y[v].x, s * x[v].y + c * y[v].y, s * x[v].z + c * y[v].z, s * x[v].w + c * y[v].w);
  }
  return more;
}

// ---- g <= 64: whole matrix in one CTA ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) jacobi_smem_kernel(const float* __restrict__ L, int g, float* __restrict__ U,
                                                           float* __restrict__ lam, int32_t* info) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float jsm[];            // column-major: column j at jsm + j * ld
  const int grp = blockIdx.x, t = threadIdx.x, nt = blockDim.x, warp = t >> 5, lane = t & 31, nw = nt >> 5;
  const int ld = g | 1;                                    // odd pitch: the transposing load below is conflict-free
  const float* Lg = L + (long long)grp * g * g;
  for (int idx = t; idx < g * g; idx += nt) { const int i = idx / g, j = idx % g; jsm[j * ld + i] = Lg[idx]; }
  const int n = (g + 1) & ~1, half = n / 2;                // one dummy player when g is odd
  __syncthreads();
  int sweep = 0;
  for (; sweep < JMAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < n - 1; ++r) {
      for (int i = warp; i < half; i += nw) {
        int p, q; rr_pair(n, r, i, p, q);
        if (p < g && q < g) rotated |= (g <= 32 ? rotate_pair<1>(jsm + p * ld, jsm + q * ld, g, lane)
                                                : rotate_pair<2>(jsm + p * ld, jsm + q * ld, g, lane)) ? 1 : 0;
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) { ++sweep; break; }
  }
  if (t == 0 && info) atomicMax(&info[WC_INFO_JACOBI], sweep);
  float* Ug = U + (long long)grp * g * g;
  for (int j = warp; j < g; j += nw) {
    const float* bj = jsm + j * ld;
    float nn = 0.f;
    for (int i = lane; i < g; i += 32) nn = fmaf(bj[i], bj[i], nn);
    nn = wc_warp_sum(nn);
    const float inv = rsqrtf(fmaxf(nn, 1e-38f));
    for (int i = lane; i < g; i += 32) Ug[(long long)i * g + j] = bj[i] * inv;
    if (lane == 0) lam[grp * g + j] = nn;
  }
}

// ---- g in {16, 32, 64}: eight lanes per column pair ----------------------------------------------------------------------
// The one-warp-per-pair kernel above is issue bound at g = 64 (32 warps x ~70 instructions per round on one SM, a 5-step
// shuffle reduction for three 64-element dot products).  Here a pair is rotated by EIGHT lanes holding g / 8 elements of
// each column in registers (128-bit shared-memory accesses), so a warp rotates four pairs at once: g / 8 warps per group,
// a 3-step reduction, ~28 instructions per pair and a barrier among 256 instead of 1024 threads.
template <int EPL>      // elements per lane: g = 8 * EPL, EPL in {2, 4, 8}
__global__ void __launch_bounds__(256) jacobi_q8_kernel(const float* __restrict__ L, float* __restrict__ U, float* __restrict__ lam,
                                                        int32_t* info) {
  wc_pdl_wait();
  constexpr int g = 8 * EPL, ld = g + 4, half = g / 2;
  constexpr int VW = EPL >= 4 ? 4 : 2;                     // vector width of the accesses
  constexpr int NV = EPL / VW;                             // vectors per lane and column
  extern __shared__ __align__(16) float jsm[];            // column-major, pitch ld (16-byte aligned columns)
  const int grp = blockIdx.x, t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5;
  const int l8 = lane & 7, slot = warp * 4 + (lane >> 3);  // this lane's pair of the round
  const float* Lg = L + (long long)grp * g * g;
  for (int idx = t; idx < g * g; idx += nt) { const int i = idx / g, j = idx % g; jsm[j * ld + i] = Lg[idx]; }
  __syncthreads();
  int sweep = 0;
  for (; sweep < JMAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < g - 1; ++r) {
      if (slot < half) {
        int p, q; rr_pair(g, r, slot, p, q);
        float* bp = jsm + p * ld + VW * l8;                // lane's elements: VW * l8 + (8 * VW) * v + 0 .. VW-1
        float* bq = jsm + q * ld + VW * l8;
        float x[EPL], y[EPL];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if constexpr (VW == 4) {
            const float4 a = *reinterpret_cast<const float4*>(bp + 32 * v), b = *reinterpret_cast<const float4*>(bq + 32 * v);
            x[4 * v] = a.x; x[4 * v + 1] = a.y; x[4 * v + 2] = a.z; x[4 * v + 3] = a.w;
            y[4 * v] = b.x; y[4 * v + 1] = b.y; y[4 * v + 2] = b.z; y[4 * v + 3] = b.w;
          } else {
            const float2 a = *reinterpret_cast<const float2*>(bp + 16 * v), b = *reinterpret_cast<const float2*>(bq + 16 * v);
            x[2 * v] = a.x; x[2 * v + 1] = a.y; y[2 * v] = b.x; y[2 * v + 1] = b.y;
          }
        }
#pragma unroll
        for (int e = 0; e < EPL; ++e) { al = fmaf(x[e], x[e], al); be = fmaf(y[e], y[e], be); ga = fmaf(x[e], y[e], ga); }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {                  // the eight lanes of the pair
          al += __shfl_xor_sync(FULL, al, o);
          be += __shfl_xor_sync(FULL, be, o);
          ga += __shfl_xor_sync(FULL, ga, o);
        }
        if (ga * ga > JROT2 * al * be) {
          if (ga * ga > JSTOP2 * al * be) rotated = 1;
          const float zeta = __fdividef(be - al, 2.f * ga);
          const float tt = copysignf(__frcp_rn(fabsf(zeta) + __fsqrt_rn(fmaf(zeta, zeta, 1.f))), zeta);
          const float c = rsqrtf(fmaf(tt, tt, 1.f)), sn = c * tt;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if constexpr (VW == 4) {
              *reinterpret_cast<float4*>(bp + 32 * v) = make_float4(c * x[4 * v] - sn * y[4 * v], c * x[4 * v + 1] - sn * y[4 * v + 1],
                                                                    c * x[4 * v + 2] - sn * y[4 * v + 2], c * x[4 * v + 3] - sn * y[4 * v + 3]);
              *reinterpret_cast<float4*>(bq + 32 * v) = make_float4(sn * x[4 * v] + c * y[4 * v], sn * x[4 * v + 1] + c * y[4 * v + 1],
                                                                    sn * x[4 * v + 2] + c * y[4 * v + 2], sn * x[4 * v + 3] + c * y[4 * v + 3]);
            } else {
              *reinterpret_cast<float2*>(bp + 16 * v) = make_float2(c * x[2 * v] - sn * y[2 * v], c * x[2 * v + 1] - sn * y[2 * v + 1]);
              *reinterpret_cast<float2*>(bq + 16 * v) = make_float2(sn * x[2 * v] + c * y[2 * v], sn * x[2 * v + 1] + c * y[2 * v + 1]);
            }
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) { ++sweep; break; }
  }
  if (t == 0 && info) atomicMax(&info[WC_INFO_JACOBI], sweep);
  float* Ug = U + (long long)grp * g * g;
  for (int j = warp; j < g; j += nt >> 5) {
    const float* bj = jsm + j * ld;
    float nn = 0.f;
    for (int i = lane; i < g; i += 32) nn = fmaf(bj[i], bj[i], nn);
    nn = wc_warp_sum(nn);
    const float inv = rsqrtf(fmaxf(nn, 1e-38f));
    for (int i = lane; i < g; i += 32) Ug[(long long)i * g + j] = bj[i] * inv;
    if (lane == 0) lam[grp * g + j] = nn;
  }
}

// ---- g >= 128: blocks of 16 columns, one CTA per block pair, barriers in L2 ---------------------------------------------
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// sync: per group {barrier counter, rotated flag of sweep 0 .. JMAX_SWEEPS-1}, zero on entry
template <int EPL>
__global__ void __launch_bounds__(JBLK_THREADS, 1) jacobi_block_kernel(const float* __restrict__ L, int g, float* __restrict__ Bm,
                                                                       float* __restrict__ U, float* __restrict__ lam,
                                                                       unsigned* __restrict__ sync, int32_t* info) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float jsm[];            // 2 blocks x JB columns x g
  const int nb = g / JB, cpg = nb / 2;                     // blocks, CTAs per group
  const int grp = blockIdx.x / cpg, c = blockIdx.x % cpg;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* Lg = L + (long long)grp * g * g;
  float* B = Bm + (long long)grp * g * g;                  // column-major working copy in global memory (L2)
  unsigned* bar = sync + (size_t)grp * (1 + JMAX_SWEEPS);
  unsigned* rot = bar + 1;
  unsigned nbar = 0;
  // B = L, column-major (this CTA transposes the rows of its share)
  for (int idx = c * JBLK_THREADS + t; idx < g * g; idx += cpg * JBLK_THREADS) {
    const int i = idx / g, j = idx % g;
    B[(long long)j * g + i] = Lg[idx];
  }
  group_barrier(bar, ++nbar * cpg);
  float* blk[2] = {jsm, jsm + JB * g};
  int sweep = 0;
  for (; sweep < JMAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int R = 0; R < nb - 1; ++R) {
      int bp, bq; rr_pair(nb, R, c, bp, bq);
      // load the two blocks (each JB contiguous columns of B): 128-bit, L2 only (written by other CTAs in earlier rounds)
      // (JB * g / 4 = EPL * JBLK_THREADS vectors per block: all 2 * EPL loads of a thread are in flight together)
      {
        float4 tmp[2][EPL];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4* src = reinterpret_cast<const float4*>(B + (long long)(h ? bq : bp) * JB * g);
#pragma unroll
          for (int u = 0; u < EPL; ++u) tmp[h][u] = __ldcg(src + t + u * JBLK_THREADS);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int u = 0; u < EPL; ++u) reinterpret_cast<float4*>(blk[h])[t + u * JBLK_THREADS] = tmp[h][u];
      }
      __syncthreads();
      const int slot = warp * 4 + (lane >> 3), l8 = lane & 7;      // this lane's pair of the sub-round
      if (R == 0) {
        // pairs inside the 2 x JB columns, all of them: a full tournament of 32 players (31 sub-rounds of 16 pairs)
        for (int r = 0; r < 2 * JB - 1; ++r) {
          int p, q; rr_pair(2 * JB, r, slot, p, q);
          rotated |= rotate_pair_l8<EPL>(jsm + p * g, jsm + q * g, l8) ? 1 : 0;
          __syncthreads();
        }
      } else {
        // cross pairs only: column `slot` of the first block with column (slot + r) % JB of the second
        for (int r = 0; r < JB; ++r) {
          rotated |= rotate_pair_l8<EPL>(blk[0] + slot * g, blk[1] + ((slot + r) & (JB - 1)) * g, l8) ? 1 : 0;
          __syncthreads();
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4* dst = reinterpret_cast<float4*>(B + (long long)(h ? bq : bp) * JB * g);
        const float4* src = reinterpret_cast<const float4*>(blk[h]);
#pragma unroll
        for (int u = 0; u < EPL; ++u) dst[t + u * JBLK_THREADS] = src[t + u * JBLK_THREADS];
      }
      if (R == nb - 2 && __syncthreads_or(rotated) && t == 0) atomicOr(rot + sweep, 1u);
      group_barrier(bar, ++nbar * cpg);
    }
    unsigned any = 0;
    if (t == 0) any = __ldcg(rot + sweep);
    if (!__syncthreads_or((int)any)) { ++sweep; break; }
  }
  if (t == 0 && c == 0 && info) atomicMax(&info[WC_INFO_JACOBI], sweep);
  // columns of B are U sigma; lam = sigma^2.  This CTA normalises a contiguous share of the columns.
  float* Ug = U + (long long)grp * g * g;
  const int per = g / cpg;
  for (int j = c * per + warp; j < (c + 1) * per; j += JBLK_THREADS / 32) {
    const float* bj = B + (long long)j * g;
    float v[EPL], nn = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) { v[e] = __ldcg(bj + lane + 32 * e); nn = fmaf(v[e], v[e], nn); }
    nn = wc_warp_sum(nn);
    const float inv = rsqrtf(fmaxf(nn, 1e-38f));
#pragma unroll
    for (int e = 0; e < EPL; ++e) Ug[(long long)(lane + 32 * e) * g + j] = v[e] * inv;
    if (lane == 0) lam[grp * g + j] = nn;
  }
}

}  // namespace

// bytes of zero-initialised synchronisation state the block kernel needs (0: the shared-memory kernel is used)
size_t wc_jacobi_sync_bytes(int C, int g) {
  if (g <= 64 || (g % 32) || g > 256) return 0;
  return (size_t)(C / g) * (1 + JMAX_SWEEPS) * sizeof(unsigned);
}

// L (G,g,g) lower Cholesky factor -> U (G,g,g) eigenvectors (columns), lam (C) eigenvalues of L L^T.  Bm: (G,g,g) scratch;
// sync: wc_jacobi_sync_bytes bytes (zeroed here).  Returns WC_E_UNSUPPORTED for shapes neither kernel takes.
int wc_jacobi_launch(const float* L, int C, int g, float* Bm, float* U, float* lam, unsigned* sync, int32_t* info, cudaStream_t st) {
  const int G = C / g;
  if ((g == 16 || g == 32 || g == 64) && !(wc_debug_flags() & 33554432)) {
    const size_t smem = (size_t)g * (g + 4) * sizeof(float);
    if (g == 64) WC_CUDA(wc_launch(jacobi_q8_kernel<8>, G, 8 * 32, smem, st, L, U, lam, info));
    else if (g == 32) WC_CUDA(wc_launch(jacobi_q8_kernel<4>, G, 4 * 32, smem, st, L, U, lam, info));
    else WC_CUDA(wc_launch(jacobi_q8_kernel<2>, G, 2 * 32, smem, st, L, U, lam, info));
    WC_CHECK_LAUNCH();
    return 0;
  }
  if (g <= 64) {
    const int warps = g <= 2 ? 1 : (g / 2 < 32 ? (g + 1) / 2 : 32);
    const size_t smem = (size_t)g * (g | 1) * sizeof(float);
    WC_CUDA(wc_launch(jacobi_smem_kernel, G, warps * 32, smem, st, L, g, U, lam, info));
    WC_CHECK_LAUNCH();
    return 0;
  }
  if ((g % 32) || g > 256 || !Bm || !sync) return WC_E_UNSUPPORTED;
  WC_CUDA(cudaMemsetAsync(sync, 0, wc_jacobi_sync_bytes(C, g), st));
  const size_t smem = (size_t)2 * JB * g * sizeof(float);
  const int ctas = G * (g / JB / 2);
  if (ctas > WC_NUM_SMS) return WC_E_UNSUPPORTED;              // the CTAs of a group wait for each other: all must be resident
  auto launch = [&](auto kern) -> int {
    static bool attr = false;
    if (!attr) { WC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * JB * 256 * sizeof(float)))); attr = true; }
    WC_CUDA(wc_launch(kern, ctas, JBLK_THREADS, smem, st, L, g, Bm, U, lam, sync, info));
    WC_CHECK_LAUNCH();
    return 0;
  };
  switch (g / 32) {
    case 4: return launch(jacobi_block_kernel<4>);
    case 5: return launch(jacobi_block_kernel<5>);
    case 6: return launch(jacobi_block_kernel<6>);
    case 7: return launch(jacobi_block_kernel<7>);
    case 8: return launch(jacobi_block_kernel<8>);
    default: return WC_E_UNSUPPORTED;
  }
}
