// K2 core: Cholesky factor L of a (g x g) SPD matrix and its inverse L^-1, one CTA per group matrix,
// everything resident in shared memory as 32x32 tiles (g <= 256).
//
// Replaces tf.cholesky + tf.matrix_triangular_solve(L, I)  (gan/conditional_layers.py:665-666), which TF 1.5
// runs as cuSOLVER potrf + cuBLAS trsm launches.
//
// Layout: lower tiles (I >= J) of the matrix, each stored column-major with pitch 36 floats
// (element (r,c) at c*36 + r), so a tile is directly the "k-major" A operand of a warp-level 32x32x32 product
// and rows of a tile column are contiguous for 128-bit loads.
//
// Cholesky (right-looking, per block column J):  warp-level potrf of the diagonal tile in registers (shuffles),
// triangular solves of the panel tiles (lane = row, broadcast reads of L_JJ), rank-32 updates of the trailing tiles
// (one warp per tile, 4x8 register blocking).  The diagonal-tile inverses are computed on the side.
// Inverse (block forward substitution swept by diagonals): accumulators S_IJ = sum_K L_IK X_KJ stay in registers,
// X_IJ = -inv(L_II) S_IJ is formed with warp shuffles and written to global memory, from where it is re-read as the
// B operand of the following updates.  Dependency depth: 2 tile products per diagonal.
#include "wc_common.cuh"

namespace {

constexpr int TB = 32, TP = 36, TS = TP * TB;   // tile edge, pitch, floats per tile
constexpr int NW = 16;                          // warps per CTA
constexpr unsigned FULL = 0xffffffffu;

__device__ long long g_chol_ts[64];
#define TS_MARK(i) do { if (t == 0 && blockIdx.x == 0) g_chol_ts[i] = clock64(); } while (0)

__device__ __forceinline__ float* tile_at(float* base, int I, int J) { return base + (I * (I + 1) / 2 + J) * TS; }

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// in-place lower Cholesky of a 32x32 tile; lane owns one row.  rdiag[c] = 1 / L[c][c].
// The diagonal element of each row is tracked in its own register so that the pivot chain is
// shfl -> rsqrt -> mul -> fma per column (no dependency through the off-diagonal shuffles).
__device__ void warp_potrf32(float* T, float* rdiag, int lane, int& bad) {
  float a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = T[c * TP + lane];
  float dg = T[lane * TP + lane];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float piv = __shfl_sync(FULL, dg, k);
    if (!(piv > 0.f)) { if (!bad) bad = k + 1; piv = 1e-30f; }
    float rs = rsqrt_approx(piv);
    rs = rs * fmaf(-0.5f * piv * rs, rs, 1.5f);      // one Newton step: full fp32 accuracy
    const float lk = (lane == k) ? piv * rs : a[k] * rs;
    a[k] = lk;
    dg = fmaf(-lk, lk, dg);
    if (lane == k) rdiag[k] = rs;
#pragma unroll
    for (int j = k + 1; j < 32; ++j) {
      const float ljk = __shfl_sync(FULL, lk, j);
      a[j] = fmaf(-lk, ljk, a[j]);
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) T[c * TP + lane] = (c <= lane) ? a[c] : 0.f;
}

// solve X L^T = A for one 32x32 tile (lane = row of X), right-looking: once x[m] is final, every later
// column is updated with column m of L, which is contiguous in the column-major tile (128-bit broadcast loads).
__device__ __forceinline__ void warp_trsm_row(float (&x)[32], const float* Ljj, const float* rdiag) {
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    x[m] *= rdiag[m];
    const float xm = x[m];
#pragma unroll
    for (int q = (m + 1) / 4; q < 8; ++q) {
      const float4 l = *reinterpret_cast<const float4*>(Ljj + m * TP + 4 * q);   // L[4q..4q+3][m]
      if (4 * q + 0 > m) x[4 * q + 0] = fmaf(-xm, l.x, x[4 * q + 0]);
      if (4 * q + 1 > m) x[4 * q + 1] = fmaf(-xm, l.y, x[4 * q + 1]);
      if (4 * q + 2 > m) x[4 * q + 2] = fmaf(-xm, l.z, x[4 * q + 2]);
      if (4 * q + 3 > m) x[4 * q + 3] = fmaf(-xm, l.w, x[4 * q + 3]);
    }
  }
}

// acc(4x8 per lane) += A(32x32, k-major smem, pitch TP) * B(32x32, k-major, pitch pb; smem or global)
__device__ __forceinline__ void warp_tile_fma(float (&acc)[4][8], const float* At, const float* Bt, int pb, int ly, int lx) {
#pragma unroll 8
  for (int kk = 0; kk < 32; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(At + kk * TP + 4 * ly);
    const float4 b0 = *reinterpret_cast<const float4*>(Bt + (long long)kk * pb + 8 * lx);
    const float4 b1 = *reinterpret_cast<const float4*>(Bt + (long long)kk * pb + 8 * lx + 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
  }
}

// X = -Ad * S where S is held in the accumulator layout of this warp (B operand fetched with shuffles)
__device__ __forceinline__ void warp_neg_mul_from_regs(float (&out)[4][8], const float* Ad, const float (&S)[4][8], int ly, int lx) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) out[r][c] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 32; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(Ad + kk * TP + 4 * ly);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const int src = (kk >> 2) * 4 + lx;
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bv[c] = __shfl_sync(FULL, S[kk & 3][c], src);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) out[r][c] = fmaf(-av[r], bv[c], out[r][c]);
  }
}

// column-major smem tile (pitch TP) -> row-major global (leading dimension ld): 8 lanes write the 128 bytes of one row
__device__ __forceinline__ void store_tile_rowmajor(const float* T, float* dst, int ld, int lane) {
  const int rr = lane >> 3, q8 = lane & 7;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int row = 4 * e + rr;
    const float* s = T + (4 * q8) * TP + row;
    *reinterpret_cast<float4*>(dst + (long long)row * ld + 4 * q8) = make_float4(s[0], s[TP], s[2 * TP], s[3 * TP]);
  }
}

__device__ __forceinline__ void tri_index(int t, int& a, int& b) {   // t -> (a,b), 0 <= b <= a, t = a(a+1)/2 + b
  a = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((a + 1) * (a + 2) / 2 <= t) ++a;
  while (a * (a + 1) / 2 > t) --a;
  b = t - a * (a + 1) / 2;
}

__global__ void __launch_bounds__(NW * 32, 1) chol_inv_tiled_kernel(const float* __restrict__ Ain, int g, float* __restrict__ Lout,
                                                                    float* __restrict__ Xout, float* __restrict__ Xpad,
                                                                    int32_t* info, int split_inverse) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int NB = (g + TB - 1) / TB, gp = NB * TB;
  float* tiles = sm;                                  // NB(NB+1)/2 tiles
  float* xd = tiles + (NB * (NB + 1) / 2) * TS;       // NB diagonal inverses (column-major tiles)
  float* rdiag = xd + NB * TS;                        // gp reciprocals of diag(L)
  const int grp = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int ly = lane >> 2, lx = lane & 3;
  const float* Ag = Ain + (long long)grp * g * g;
  // X is produced row-major with leading dimension gp; when g is a multiple of 32 that is the output itself
  float* X = (Xout && gp == g) ? Xout + (long long)grp * g * g : Xpad + (long long)grp * gp * gp;

  TS_MARK(62);
  float* Lg = Lout ? Lout + (long long)grp * g * g : nullptr;
  if (gp == g) {
    // fast path (g multiple of 32).  The input is symmetric, so column c of the column-major tile (I,J) is a contiguous
    // piece of ROW J*32+c of A: 128-bit loads and 128-bit conflict-free stores, no transpose.  A warp takes whole tiles
    // (lane = column) and keeps all of its loads in flight at once.
    const int ntiles = NB * (NB + 1) / 2;
    // 8 lanes read the 128 contiguous bytes of one tile column (a row of A), a warp instruction covers 4 columns
    const int cq = lane >> 3, q8 = lane & 7;
    float4 v[3][8];
    int tI[3], tJ[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int tl = warp + q * NW;
      tI[q] = -1; tJ[q] = 0;
      if (tl < ntiles) {
        tri_index(tl, tI[q], tJ[q]);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[q][e] = __ldg(reinterpret_cast<const float4*>(Ag + (long long)(tJ[q] * TB + 4 * e + cq) * g + tI[q] * TB) + q8);
      }
    }
    TS_MARK(55);
    // strictly upper tiles of the outputs are zero; issue the stores now so they drain during the factorisation
    for (int ut = warp; ut < NB * NB; ut += NW) {
      const int I = ut / NB, J = ut % NB;
      if (J <= I) continue;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const long long off = (long long)(I * TB + 4 * e + cq) * g + J * TB + 4 * q8;
        if (Xout) *reinterpret_cast<float4*>(X + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Lg) *reinterpret_cast<float4*>(Lg + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    TS_MARK(56);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int tl = warp + q * NW;
      if (tl < ntiles) {
#pragma unroll
        for (int e = 0; e < 8; ++e) *reinterpret_cast<float4*>(tiles + tl * TS + (4 * e + cq) * TP + 4 * q8) = v[q][e];
      }
    }
  } else {
  // load the lower tiles (128-bit loads when rows are 16-byte aligned), pad with the identity; the strictly
  // upper tiles of the output X are zero-filled here so the stores drain while the factorisation runs
  const bool vec_ok = (g & 3) == 0;
#pragma unroll 4
  for (int i = warp; i < gp; i += NW) {
    for (int j4 = lane * 4; j4 < gp; j4 += 128) {
      if ((j4 >> 5) > (i >> 5)) {
        if (Xout) *reinterpret_cast<float4*>(X + (long long)i * gp + j4) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      float v[4];
      if (vec_ok && i < g && j4 + 3 < g) {
        const float4 q = *reinterpret_cast<const float4*>(Ag + (long long)i * g + j4);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int j = j4 + e; v[e] = (i < g && j < g) ? Ag[(long long)i * g + j] : (i == j ? 1.f : 0.f); }
      }
      float* T = tile_at(tiles, i >> 5, j4 >> 5) + (i & 31);
#pragma unroll
      for (int e = 0; e < 4; ++e) T[((j4 + e) & 31) * TP] = v[e];
    }
  }
  }
  __syncthreads();
  TS_MARK(0);

  // ------------------------------------------------------------------ Cholesky
  // (Overlapping the next diagonal factorisation with the trailing updates was measured and does not pay: the update
  // phase is bound by shared-memory operand bandwidth with all 16 warps busy, and the late steps are a serial
  // update -> potrf chain either way.)
  for (int J = 0; J < NB; ++J) {
    TS_MARK(1 + 3 * J);
    if (warp == 0) {
      int bad = 0;
      warp_potrf32(tile_at(tiles, J, J), rdiag + J * TB, lane, bad);
      if (bad && lane == 0 && info && J * TB + bad <= g) atomicCAS(&info[WC_INFO_NOT_PD], 0, grp * g + J * TB + bad);
    }
    __syncthreads();
    TS_MARK(2 + 3 * J);
    const float* Ljj = tile_at(tiles, J, J);
    const int npanel = NB - 1 - J;
    // panel solves (npanel tiles) + the inverse of the diagonal tile + the row-major copy of the diagonal tile
    for (int job = warp; job <= npanel + 1; job += NW) {
      float x[32];
      if (job < npanel) {
        float* T = tile_at(tiles, J + 1 + job, J);
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = T[c * TP + lane];
        warp_trsm_row(x, Ljj, rdiag + J * TB);
#pragma unroll
        for (int c = 0; c < 32; ++c) T[c * TP + lane] = x[c];
        if (Lg && gp == g) { __syncwarp(); store_tile_rowmajor(T, Lg + (long long)((J + 1 + job) * TB) * g + J * TB, g, lane); }
      } else if (job == npanel) {
        if (Xout) {
          // inv(L_JJ): row `lane` of L^-T = column `lane` of L^-1
#pragma unroll
          for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.f : 0.f;
          warp_trsm_row(x, Ljj, rdiag + J * TB);
          float* D = xd + J * TS;
#pragma unroll
          for (int c = 0; c < 32; ++c) D[lane * TP + c] = x[c];              // element (row c, col lane)
          __syncwarp();
          store_tile_rowmajor(D, X + (long long)(J * TB) * gp + J * TB, gp, lane);
        }
      } else if (Lg && gp == g) {
        store_tile_rowmajor(Ljj, Lg + (long long)(J * TB) * g + J * TB, g, lane);
      }
    }
    __syncthreads();
    TS_MARK(3 + 3 * J);
    // trailing updates: tile(I,K) -= L_IJ L_KJ^T for J < K <= I
    const int nupd = npanel * (npanel + 1) / 2;
    for (int job = warp; job < nupd; job += NW) {
      int a, b2;
      tri_index(job, a, b2);
      const int I = J + 1 + a, K = J + 1 + b2;
      float acc[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      warp_tile_fma(acc, tile_at(tiles, I, J), tile_at(tiles, K, J), TP, ly, lx);
      float* C = tile_at(tiles, I, K);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4* p = reinterpret_cast<float4*>(C + (8 * lx + c) * TP + 4 * ly);
        float4 v = *p;
        v.x -= acc[0][c]; v.y -= acc[1][c]; v.z -= acc[2][c]; v.w -= acc[3][c];
        *p = v;
      }
    }
    __syncthreads();
  }

  TS_MARK(30);
  if (Lg && gp != g) {
    for (int idx = t; idx < g * g; idx += NW * 32) {
      const int i = idx / g, j = idx % g;
      Lg[idx] = (j <= i) ? tile_at(tiles, i >> 5, j >> 5)[(j & 31) * TP + (i & 31)] : 0.f;
    }
  }
  TS_MARK(63);
  if (!Xout || split_inverse) return;   // split: trinv_kernel forms the off-diagonal blocks of X on many CTAs

  // ------------------------------------------------------------------ inverse
  // each warp owns up to two off-diagonal tiles (NB <= 8 -> 28 tiles); accumulators live in registers
  const int noff = NB * (NB - 1) / 2;
  float S[2][4][8];
  int tI[2], tJ[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int job = warp + s * NW;
    tI[s] = -1; tJ[s] = 0;
    if (job < noff) { int a, b; tri_index(job, a, b); tI[s] = a + 1; tJ[s] = b; }   // I = a+1 > J = b
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) S[s][r][c] = 0.f;
  }
  __syncthreads();   // diagonal tiles of X (global) and the zero fill are visible
  TS_MARK(31);
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (tI[s] >= 0)   // S_IJ = L_IJ X_JJ
      warp_tile_fma(S[s], tile_at(tiles, tI[s], tJ[s]), X + (long long)(tJ[s] * TB) * gp + tJ[s] * TB, gp, ly, lx);
  for (int d = 1; d < NB; ++d) {
    TS_MARK(32 + 2 * d);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (tI[s] >= 0 && tI[s] - tJ[s] == d) {
        float out[4][8];
        warp_neg_mul_from_regs(out, xd + tI[s] * TS, S[s], ly, lx);
        float* dst = X + (long long)(tI[s] * TB + 4 * ly) * gp + tJ[s] * TB + 8 * lx;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          *reinterpret_cast<float4*>(dst + (long long)r * gp) = make_float4(out[r][0], out[r][1], out[r][2], out[r][3]);
          *reinterpret_cast<float4*>(dst + (long long)r * gp + 4) = make_float4(out[r][4], out[r][5], out[r][6], out[r][7]);
        }
      }
    }
    __syncthreads();
    TS_MARK(33 + 2 * d);
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (tI[s] >= 0 && tI[s] - tJ[s] > d)   // S_IJ += L_{I,J+d} X_{J+d,J}
        warp_tile_fma(S[s], tile_at(tiles, tI[s], tJ[s] + d), X + (long long)((tJ[s] + d) * TB) * gp + tJ[s] * TB, gp, ly, lx);
  }
  TS_MARK(50);
  if (gp != g) {
    __syncthreads();
    float* Xg = Xout + (long long)grp * g * g;
    for (int idx = t; idx < g * g; idx += NW * 32) Xg[idx] = X[(long long)(idx / g) * gp + idx % g];
  }
}


// ---- L^-1 on many CTAs ------------------------------------------------------------------------------------
// Column j of X = L^-1 depends only on L, so the inverse parallelises over columns: one CTA per slab of 8 columns
// runs a blocked forward substitution  x_I = inv(L_II) (e_I - sum_{K<I} L_IK x_K)  with the 32 x 32 diagonal-block
// inverses that the Cholesky kernel already wrote into X.  g must be a multiple of 32.  L is row-major (G,g,g).
constexpr int TI_COLS = 8, TI_THREADS = 32 * TI_COLS, TI_LP = 260, TI_DP = 33, TI_NPRE = 7;

__global__ void __launch_bounds__(TI_THREADS) trinv_kernel(const float* __restrict__ L, int g, float* __restrict__ X, unsigned long long* dbg, int dflags) {
  wc_pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int NB = g / TB;
  const int slabs = g / TI_COLS;
  const int grp = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const int j0 = slab * TI_COLS, J = j0 / TB;
  if (J == NB - 1) return;                             // last block column: only the diagonal block, already in X
  const float* Lg = L + (long long)grp * g * g;
  float* Xg = X + (long long)grp * g * g;
  float* Ls = sm;                                      // [32][TI_LP] current row panel of L (rows read back as float4)
  float* Ds = Ls + TB * TI_LP;                         // [32][TI_DP] inverse of the current diagonal block
  float* xs = Ds + TB * TI_DP;                         // [TI_COLS][g] solution, k contiguous (16-byte aligned)
  float* rs = xs + TI_COLS * g;                        // [32][TI_COLS + 1] right-hand side of the current block row
  const int t = threadIdx.x, r = t & 31, c = t >> 5;
  // x_J = inv(L_JJ)[:, j0..j0+7]: columns of the diagonal-block inverse the Cholesky kernel left in X
  xs[c * g + J * TB + r] = Xg[(long long)(J * TB + r) * g + j0 + c];
  // register prefetch of the next block row: panel L[I-block rows, 32J .. 32I) (up to 32 x 224 floats) and inv(L_II)
  float4 pre[TI_NPRE], dpre;
  auto prefetch = [&](int I) {
    // 8 threads per panel row, each takes every 8th float4 of the row (32 (I - J) floats per row)
    const float4* src = reinterpret_cast<const float4*>(Lg + (long long)(I * TB + (t >> 3)) * g + J * TB) + (t & 7);
#pragma unroll
    for (int q = 0; q < TI_NPRE; ++q)
      if (q < I - J && !(dflags & 16)) pre[q] = __ldg(src + 8 * q);
    if (!(dflags & 16)) dpre = __ldg(reinterpret_cast<const float4*>(Xg + (long long)(I * TB + (t >> 3)) * g + I * TB) + (t & 7));
  };
  auto commit = [&](int I) {
    float* dstp = Ls + (t >> 3) * TI_LP + (t & 7) * 4;
#pragma unroll
    for (int q = 0; q < TI_NPRE; ++q)
      if (q < I - J) *reinterpret_cast<float4*>(dstp + 32 * q) = pre[q];
    float* d = Ds + (t >> 3) * TI_DP + (t & 7) * 4;
    d[0] = dpre.x; d[1] = dpre.y; d[2] = dpre.z; d[3] = dpre.w;
  };
  if (t == 0) wc_mark(dbg, 0);
  prefetch(J + 1);
  for (int I = J + 1; I < NB; ++I) {
    __syncthreads();                                   // previous step's readers of Ls / Ds / rs are done; xs of block I-1 visible
    if (t == 0 && I - J < WC_DBG_SLOTS) wc_mark(dbg, I - J);
    commit(I);
    if (I + 1 < NB) prefetch(I + 1);
    __syncthreads();
    // rhs[r][c] = - sum_k L[I*32 + r][J*32 + k] x[k][c]
    const int kn = (I - J) * TB;
    const float* lrow = Ls + r * TI_LP;
    const float* xc = xs + c * g + J * TB;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < kn; k += 4) {
      const float4 lv = *reinterpret_cast<const float4*>(lrow + k);
      const float4 xv = *reinterpret_cast<const float4*>(xc + k);
      a0 = fmaf(lv.x, xv.x, a0); a1 = fmaf(lv.y, xv.y, a1);
      a2 = fmaf(lv.z, xv.z, a2); a3 = fmaf(lv.w, xv.w, a3);
    }
    rs[r * (TI_COLS + 1) + c] = -((a0 + a1) + (a2 + a3));
    __syncthreads();
    // x_I = inv(L_II) rhs   (lower triangular)
    const float* drow = Ds + r * TI_DP;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int k = 0; k < TB; k += 2) {
      v0 = fmaf(drow[k], rs[k * (TI_COLS + 1) + c], v0);          // the strict upper triangle of inv(L_II) is stored as zeros
      v1 = fmaf(drow[k + 1], rs[(k + 1) * (TI_COLS + 1) + c], v1);
    }
    const float v = v0 + v1;
    xs[c * g + I * TB + r] = v;
    if (!(dflags & 32)) Xg[(long long)(I * TB + r) * g + j0 + c] = v;
  }
}


// ======================================================================================================
// Cluster Cholesky + inverse: one thread-block cluster of 8 CTAs per group matrix (g multiple of 32, 128 <= g <= 256)
// ======================================================================================================
// The one-CTA kernel above is bound by the serial chain potrf -> panel solves -> trailing updates with every tile
// product on one SM.  Here CTA r of the cluster owns block-row r of L (left-looking):
//     L_rJ = (A_rJ - sum_{K<J} L_rK L_JK^T) L_JJ^-T   for J < r,      L_rr = chol(A_rr - sum_{K<r} L_rK L_rK^T)
// and, once its row is done, block-column r of X = L^-1:
//     X_rr = inv(L_rr),     X_Ir = -inv(L_II) sum_{K=r..I-1} L_IK X_Kr   for I > r.
// Finished tiles are published through a global scratch (L2) as column-major 32x32 tiles and announced with remote
// mbarrier arrivals in the consumers' shared memory (DSMEM): A[J] = off-diagonal tiles of row J, B[J] = L_JJ and
// 1/diag, D[J] = inv(L_JJ).  The critical path is  potrf(J) -> publish -> solve of tile (J+1,J) -> diagonal update ->
// potrf(J+1); everything else runs behind that wavefront on the other CTAs.
constexpr int CLN = 8;                 // CTAs per cluster = maximum number of block rows
constexpr int CLW = 8;                 // warps per CTA
constexpr int GT = TB * TB;            // floats per published tile (column-major, pitch 32)
__device__ unsigned long long g_cl_ts[CLN][32];
#define CL_MARKW(i) do { if (dbg) { if (lane == 0 && grp == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); g_cl_ts[r][i] = t__; } __syncwarp(); } } while (0)
#define CL_MARK(i) do { if (dbg) { if (lane == 0 && warp == 0 && grp == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); g_cl_ts[r][i] = t__; } __syncwarp(); } } while (0)

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void cl_arrive_remote(uint64_t* local_bar, uint32_t rank) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_bar);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cl_wait(uint64_t* bar) {   // each flag is used once per launch: phase 0
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "CL_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], 0;\n\t"
      "@P1 bra CL_DONE;\n\t"
      "bra CL_WAIT;\n\t"
      "CL_DONE:\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ uint32_t cl_map(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ float4 cl_ld4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// acc(4x8 per lane) += A(32x32, k-major, pitch PA) * B(32x32, k-major, pitch PB); GA / GB: operand lives in the global
// scratch (read through L2 only: it was written by another SM during this launch)
template <int PA, int PB, bool GA, bool GB>
__device__ __forceinline__ void cl_tile_fma(float (&acc)[4][8], const float* At, const float* Bt, int ly, int lx) {
#pragma unroll 8
  for (int kk = 0; kk < 32; ++kk) {
    const float* ap = At + kk * PA + 4 * ly;
    const float* bp = Bt + kk * PB + 8 * lx;
    const float4 a = GA ? ldcg4(ap) : *reinterpret_cast<const float4*>(ap);
    const float4 b0 = GB ? ldcg4(bp) : *reinterpret_cast<const float4*>(bp);
    const float4 b1 = GB ? ldcg4(bp + 4) : *reinterpret_cast<const float4*>(bp + 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
  }
}

// in-place lower Cholesky of a 32x32 tile by one warp (lane = row).  Column k of the factor is broadcast through
// shared memory (one store + 128-bit broadcast loads) instead of 31-k shuffles; the pivot chain per column is
// shfl -> rsqrt -> mul -> fma.  rsqrt.approx is accurate to 2^-22.4 relative (PTX ISA), i.e. a few ulp of the fp32 result.
// bc: 64 floats of shared memory.
__device__ __forceinline__ void cta_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cta_mbar_wait0(uint64_t* bar) {     // one-shot barriers: phase 0
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "CB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
      "@P1 bra CB_DONE;\n\t"
      "bra CB_WAIT;\n\t"
      "CB_DONE:\n\t}" ::"r"(a) : "memory");
}

// Columns are announced to the follower in groups of four, singly for the last four: a wait per column made the follower
// slower per step than the factorisation it follows (it fell 0.65 us behind by column 31); with a wait per group it runs
// ahead of the announcements and ends about one column behind.
__device__ __forceinline__ constexpr bool colbar_signalled(int k) { return (k & 3) == 3 || k >= 28; }

// in-place lower Cholesky of a 32x32 tile by one warp (lane = row).  Column k of the factor is written to the tile as
// soon as it is final -- it is at once the broadcast buffer for the trailing update of this warp (128-bit broadcast loads)
// and, announced on colbar[k], the input of the warp that builds the inverse one column behind (warp_inverse_follow).
// The serial chain per column is  pivot -> rsqrt -> l_k -> (diagonal of lane k+1, element (.,k+1)) -> next pivot; both
// hand-overs to column k+1 go through shuffles.  rsqrt.approx is accurate to 2^-22.4 relative (PTX ISA).
__device__ void warp_potrf32_bc(float* T, float* rdiag, uint64_t* colbar, int lane, int& bad) {
  float a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = T[c * TP + lane];
  float dg = T[lane * TP + lane];
  __syncwarp();                                   // every lane holds its row before columns are overwritten
  float piv = __shfl_sync(FULL, dg, 0);
  if (!(piv > 0.f)) { if (!bad) bad = 1; piv = 1e-30f; }
  float rs = rsqrt_approx(piv);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float lk = (lane == k) ? piv * rs : a[k] * rs;
    if (lane == k) rdiag[k] = rs;
    T[k * TP + lane] = (lane >= k) ? lk : 0.f;    // final column k (zero above the diagonal)
    if (k < 31) {
      dg = fmaf(-lk, lk, dg);
      piv = __shfl_sync(FULL, dg, k + 1);                  // lane k+1's diagonal is final after column k
      const float lnext = __shfl_sync(FULL, lk, k + 1);    // L[k+1][k]
      a[k + 1] = fmaf(-lk, lnext, a[k + 1]);
      // the reciprocal square root of the NEXT pivot is started here, so that its latency (the longest link of the
      // per-column chain) runs under the trailing update below instead of at the top of the next column
      if (!(piv > 0.f)) { if (!bad) bad = k + 2; piv = 1e-30f; }
      rs = rsqrt_approx(piv);
    }
    __syncwarp();
    if (colbar && lane == 0 && colbar_signalled(k)) cta_mbar_arrive(&colbar[k]);
    if (k < 30) {
      const float* b = T + k * TP;
#pragma unroll
      for (int q = (k + 2) / 4; q < 8; ++q) {
        const float4 l = *reinterpret_cast<const float4*>(b + 4 * q);
        if (4 * q + 0 > k + 1) a[4 * q + 0] = fmaf(-lk, l.x, a[4 * q + 0]);
        if (4 * q + 1 > k + 1) a[4 * q + 1] = fmaf(-lk, l.y, a[4 * q + 1]);
        if (4 * q + 2 > k + 1) a[4 * q + 2] = fmaf(-lk, l.z, a[4 * q + 2]);
        if (4 * q + 3 > k + 1) a[4 * q + 3] = fmaf(-lk, l.w, a[4 * q + 3]);
      }
    }
  }
}

// one warp, one column behind warp_potrf32_bc: x (row `lane` of L^-T = column `lane` of L^-1) by right-looking
// substitution; step m needs column m of L and 1/L[m][m], which colbar[m] announces
__device__ void warp_inverse_follow(float (&x)[32], const float* T, const float* rdiag, uint64_t* colbar, int lane) {
#pragma unroll
  for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.f : 0.f;
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    // a group is consumed once its LAST column is announced
    if (colbar && m < 28 && (m & 3) == 0) cta_mbar_wait0(&colbar[m + 3]);
    if (colbar && m >= 28) cta_mbar_wait0(&colbar[m]);
    x[m] *= rdiag[m];
    const float xm = x[m];
#pragma unroll
    for (int q = (m + 1) / 4; q < 8; ++q) {
      const float4 l = *reinterpret_cast<const float4*>(T + m * TP + 4 * q);   // L[4q..4q+3][m]
      if (4 * q + 0 > m) x[4 * q + 0] = fmaf(-xm, l.x, x[4 * q + 0]);
      if (4 * q + 1 > m) x[4 * q + 1] = fmaf(-xm, l.y, x[4 * q + 1]);
      if (4 * q + 2 > m) x[4 * q + 2] = fmaf(-xm, l.z, x[4 * q + 2]);
      if (4 * q + 3 > m) x[4 * q + 3] = fmaf(-xm, l.w, x[4 * q + 3]);
    }
  }
}

// all 256 threads: acc(2x2 per thread: rows 2tr+i, cols 2tc+j) += A(32x32, k-major, pitch TP) * B(32x32, k-major, pitch TP)
__device__ __forceinline__ void cta_tile_mac(float (&acc)[2][2], const float* At, const float* Bt, int tr, int tc) {
#pragma unroll
  for (int kk = 0; kk < 32; ++kk) {
    const float2 a = *reinterpret_cast<const float2*>(At + kk * TP + 2 * tr);
    const float2 b = *reinterpret_cast<const float2*>(Bt + kk * TP + 2 * tc);
    acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
    acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
  }
}
// all 256 threads: n published tiles (pitch 32, stride GT) -> column-major smem tiles (pitch TP, stride TS)
__device__ __forceinline__ void cta_fetch_tiles(const float* gt, int n, float* dst, int t) {
  const int col = t >> 3, q8 = t & 7;
  float4 v[CLN];
#pragma unroll
  for (int i = 0; i < CLN; ++i)
    if (i < n) v[i] = ldcg4(gt + i * GT + col * TB + 4 * q8);
#pragma unroll
  for (int i = 0; i < CLN; ++i)
    if (i < n) *reinterpret_cast<float4*>(dst + i * TS + col * TP + 4 * q8) = v[i];
}

__global__ void __cluster_dims__(CLN, 1, 1) __launch_bounds__(CLW * 32, 1)
    chol_cluster_kernel(const float* __restrict__ Ain, int g, float* __restrict__ scratch, long long scratch_stride,
                        float* __restrict__ Lout, float* __restrict__ Xout, int32_t* info, int dbg) {
  // (the dependency wait comes after the barriers are initialised, the strictly upper output tiles are zeroed and the cluster
  // has synchronised: none of that reads what the preceding kernel writes, so it runs under that kernel's tail)
  extern __shared__ __align__(16) float sm[];
  __shared__ uint64_t flagA[CLN], flagA2[CLN], flagB[CLN], colbar[TB], flagN, flagT;
  __shared__ __align__(16) float rd_own[TB], rd_warm[2 * TB];
  const int NB = g / TB;
  const int grp = blockIdx.x / CLN;
  const int r = (int)cl_rank();
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int tr = t & 15, tc = t >> 4;
  float* own = sm;                       // CLN tiles: block row r of L, column-major
  float* xcol = own + CLN * TS;          // CLN tiles: block column r of X, row-major, indexed by block row
  float* ljj = xcol + CLN * TS;          // staged L_JJ / inv(L_II), column-major
  float* stile = ljj + TS;               // S tile, row-major; also inv(L_rr) column-major
  float* stg = stile + TS;               // CLN staged tiles of another block row
  float* dinv = stg + CLN * TS;          // inv(L_rr), column-major: what the rows below multiply their tiles with
  float* dnext = dinv + TS;              // inv(L_{r-1,r-1}), pushed here by CTA r-1 (the hand-over on the critical path)
  float* lnext = dnext + TS;             // L_{r-1,r-2}, pushed here by CTA r-1: the last tile our row's products wait for
  const float* Ag = Ain + (long long)grp * g * g;
  float* gs = scratch + (long long)grp * scratch_stride;
  float* Lg = Lout ? Lout + (long long)grp * g * g : nullptr;
  float* Xg = Xout + (long long)grp * g * g;
  auto gtile = [&](int I, int J) { return gs + (I * (I + 1) / 2 + J) * GT; };

  CL_MARK(0);
  if (t == 0) {
    for (int j = 0; j < CLN; ++j) {
      const uint32_t a = (uint32_t)__cvta_generic_to_shared(&flagA[j]), b = (uint32_t)__cvta_generic_to_shared(&flagB[j]),
                     a2 = (uint32_t)__cvta_generic_to_shared(&flagA2[j]);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a2) : "memory");
    }
    for (int j = 0; j < TB; ++j)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&colbar[j])) : "memory");
    // flagN: the inverse of the previous CTA's diagonal tile arrives as asynchronous stores that count their bytes on this
    // barrier (32 x 32 floats); armed here, before the cluster barrier that precedes any push
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&flagN)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&flagN)), "r"(TB * TB * 4) : "memory");
    // flagT: the same for tile (r-1, r-2) of L
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&flagT)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&flagT)), "r"(TB * TB * 4) : "memory");
  }
  if (r < NB) {
    // strictly upper tiles of the outputs are zero
    const int cq = lane >> 3, q8 = lane & 7;
    for (int J = r + 1 + warp; J < NB; J += CLW) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const long long off = (long long)(r * TB + 4 * e + cq) * g + J * TB + 4 * q8;
        *reinterpret_cast<float4*>(Xg + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Lg) *reinterpret_cast<float4*>(Lg + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();
  cl_sync();      // every CTA's flags are initialised before anyone can arrive on them
  wc_pdl_wait();
  // load block row r: column c of the column-major tile (r,J) is a contiguous piece of ROW J*32+c of the symmetric input
  if (r < NB) {
    const int cq = lane >> 3, q8 = lane & 7;
    for (int J = warp; J <= r; J += CLW) {
      float4 v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldcg(reinterpret_cast<const float4*>(Ag + (long long)(J * TB + 4 * e + cq) * g + r * TB) + q8);
#pragma unroll
      for (int e = 0; e < 8; ++e) *reinterpret_cast<float4*>(own + J * TS + (4 * e + cq) * TP + 4 * q8) = v[e];
    }
  }
  __syncthreads();
  CL_MARK(1);

  if (r < NB) {
    // The factorisation of a diagonal tile and the inverse that follows it are ~75 KB of unrolled straight-line code
    // that a CTA runs exactly once, on the critical path; cold, they are bound by instruction fetch.  Rows below the first
    // wait ~4 us for CTA 0 anyway: in "phase 0" two otherwise idle warps run THE SAME code (one call site, this loop is
    // not unrolled) on scratch tiles, so that phase 1 -- the real thing -- finds it in the SM's instruction cache.
#pragma unroll 1
    for (int phase = (r > 0) ? 0 : 1; phase < 2; ++phase) {
    if (phase == 1) {
    // ------------------------------------------------------------------ block row r of L
    for (int J = 0; J < r; ++J) {
      float* T = own + J * TS;
      // T_rJ -= sum_{K<J} L_rK L_JK^T, row J's tiles coming from the scratch.  Row J is announced in two parts: its tiles
      // 0..J-2 one block step early (flagA2), the last one (J,J-1) with flagA.  The early part of the NEXT column's sum is
      // taken here too, before this column's wait for inv(L_JJ): the CTA would idle there, and afterwards the same products
      // would sit between CTA J's factorisation and ours (for the CTA next to the diagonal they were the critical path).
      auto subtract = [&](float* Tt, const float (&acc)[2][2]) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float2* p = reinterpret_cast<float2*>(Tt + (2 * tc + j) * TP + 2 * tr);
          float2 v = *p;
          v.x -= acc[0][j]; v.y -= acc[1][j];
          *p = v;
        }
        __syncthreads();
      };
      if (J > 0) {
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        if (J == r - 1) {
          // our predecessor pushed this tile into our shared memory the moment it had it: no scratch round trip, no fence
          cl_wait(&flagT);
          cta_tile_mac(acc, own + (J - 1) * TS, lnext, tr, tc);
        } else {
          cl_wait(&flagA[J]);
          cta_fetch_tiles(gtile(J, J - 1), 1, stg + (J - 1) * TS, t);
          __syncthreads();
          cta_tile_mac(acc, own + (J - 1) * TS, stg + (J - 1) * TS, tr, tc);
        }
        subtract(T, acc);
      }
      if (J >= 1 && J + 1 < r) {
        const int Jn = J + 1;               // tiles (Jn, 0..Jn-2) x our tiles (r, 0..Jn-2), all final since column J-1
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        cl_wait(&flagA2[Jn]);
        cta_fetch_tiles(gtile(Jn, 0), Jn - 1, stg, t);
        __syncthreads();
        for (int K = 0; K < Jn - 1; ++K) cta_tile_mac(acc, own + K * TS, stg + K * TS, tr, tc);
        subtract(own + Jn * TS, acc);
      }
      {
        // L_rJ = T inv(L_JJ)^T as a tile product by the whole CTA; inv(L_JJ) comes straight out of CTA J's shared memory
        // (it was built one column behind J's factorisation, so it is there as soon as L_JJ is)
        if (J == r - 1) CL_MARK(2);
        // the inverse of the predecessor's diagonal tile is the hand-over on the critical path: it comes as st.async stores
        // that complete on flagN -- data and signal in one operation; plain remote stores followed by a release-arrive
        // took 0.45 us on the sender (the release waits for the stores) and 0.3 us more to wake us
        if (J == r - 1) cl_wait(&flagN); else cl_wait(&flagB[J]);
        if (J == r - 1) CL_MARK(3);
        const float* Dj = dnext;                 // J == r-1: CTA J pushed its inverse into our shared memory
        if (J < r - 1) {
          const int col = t >> 3, q8 = t & 7;
          const uint32_t rt = cl_map((uint32_t)__cvta_generic_to_shared(dinv), (uint32_t)J);
          const float4 v = cl_ld4(rt + (uint32_t)((col * TP + 4 * q8) * 4));
          *reinterpret_cast<float4*>(ljj + col * TP + 4 * q8) = v;
          Dj = ljj;
          __syncthreads();
        }
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        cta_tile_mac(acc, T, Dj, tr, tc);        // (row 2tr+i, col 2tc+j) = sum_k T(row,k) inv(L_JJ)(col,k)
        __syncthreads();                         // everyone has read T
        float* gt = gtile(r, J);
        if (J == r - 1 && r + 1 < NB) {
          const uint32_t rn = cl_map((uint32_t)__cvta_generic_to_shared(lnext), (uint32_t)(r + 1));
          const uint32_t rb = cl_map((uint32_t)__cvta_generic_to_shared(&flagT), (uint32_t)(r + 1));
#pragma unroll
          for (int j = 0; j < 2; ++j)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];" ::"r"(
                             rn + (uint32_t)(((2 * tc + j) * TP + 2 * tr) * 4)), "f"(acc[0][j]), "f"(acc[1][j]), "r"(rb) : "memory");
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float2 v = make_float2(acc[0][j], acc[1][j]);
          *reinterpret_cast<float2*>(T + (2 * tc + j) * TP + 2 * tr) = v;
          *reinterpret_cast<float2*>(gt + (2 * tc + j) * TB + 2 * tr) = v;
        }
        __syncthreads();
        if (J == r - 1) CL_MARK(4);
        // (row r is announced to the other CTAs from the diagonal phase below: they have a whole factorisation of slack)
        if (J == r - 2 && warp == 3) {
          // tiles (r, 0..r-2) are in the scratch (stored before the barrier above): the early part of row r
          cl_fence();
          __syncwarp();
          if (lane < NB && lane > r) cl_arrive_remote(&flagA2[r], lane);
        }
      }
      {   // right-looking update of the diagonal tile: T_rr -= L_rJ L_rJ^T
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        cta_tile_mac(acc, T, T, tr, tc);
        float* Td = own + r * TS;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float2* p = reinterpret_cast<float2*>(Td + (2 * tc + j) * TP + 2 * tr);
          float2 v = *p;
          v.x -= acc[0][j]; v.y -= acc[1][j];
          *p = v;
        }
      }
      if (Lg && warp == 1) store_tile_rowmajor(T, Lg + (long long)(r * TB) * g + J * TB, g, lane);
    }
    __syncthreads();
    CL_MARK(5);
    }   // phase 1
    // ------------------------------------------------------------------ diagonal tile
    // one warp factors it column by column; another follows one column behind with the inverse, so inv(L_rr) exists
    // (almost) as soon as L_rr does and the rows below never run a triangular solve
    const int w_fac = phase ? 0 : CLW - 1, w_inv = phase ? 1 : CLW - 2;
    float* Tt = phase ? own + r * TS : stg;                    // phase 0: scratch tiles, values irrelevant
    float* rdv = phase ? rd_own : rd_warm;
    uint64_t* cb = phase ? colbar : nullptr;
    if (warp == w_fac) {
      CL_MARK(6);
      int bad = 0;
      warp_potrf32_bc(Tt, rdv, cb, lane, bad);
      CL_MARK(7);
      if (phase) {
        if (bad && lane == 0 && info) atomicCAS(&info[WC_INFO_NOT_PD], 0, grp * g + r * TB + bad);
        __syncwarp();
        if (Lg) store_tile_rowmajor(Tt, Lg + (long long)(r * TB) * g + r * TB, g, lane);
      }
    } else if (warp == w_inv) {
      float x[32];
      warp_inverse_follow(x, phase ? Tt : stg + TS, phase ? rdv : rd_warm + TB, cb, lane);
      if (phase) {
      CL_MARKW(16);
      // dinv: element (row c, col lane).  The next CTA gets its copy pushed (it needs it first, on the critical path).
      if (r + 1 < NB) {
        const uint32_t rn = cl_map((uint32_t)__cvta_generic_to_shared(dnext), (uint32_t)(r + 1)) + (uint32_t)(lane * TP * 4);
        const uint32_t rb = cl_map((uint32_t)__cvta_generic_to_shared(&flagN), (uint32_t)(r + 1));
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(rn + c * 4),
                       "f"(x[c]), "f"(x[c + 1]), "f"(x[c + 2]), "f"(x[c + 3]), "r"(rb) : "memory");
      }
#pragma unroll
      for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dinv + lane * TP + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
      // the arrivals are release operations at cluster scope and cumulative over what __syncwarp ordered before them:
      // no separate fence on this (critical) hand-over
      __syncwarp();
      if (lane < NB && lane != r && lane != r + 1) cl_arrive_remote(&flagB[r], lane);
      CL_MARKW(8);
      float* Xr = xcol + r * TS;              // row-major
#pragma unroll
      for (int c = 0; c < 32; ++c) Xr[c * TP + lane] = x[c];
      __syncwarp();
      CL_MARKW(9);
      store_tile_rowmajor(dinv, Xg + (long long)(r * TB) * g + r * TB, g, lane);
      } else {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc += x[c];
        rd_warm[TB + lane] = acc;              // keeps the warm-up alive
      }
    } else if (phase && warp == 2 && r > 0) {
      // all off-diagonal tiles of row r are in the scratch (written before the barrier above): announce them
      cl_fence();
      __syncwarp();
      if (lane < NB && lane != r) cl_arrive_remote(&flagA[r], lane);
    }
    if (phase) {
      __syncthreads();
      CL_MARK(10);
    }
    }   // phase loop
    // ------------------------------------------------------------------ block column r of X
    for (int I = r + 1; I < NB; ++I) {
      cl_wait(&flagA[I]);
      if (I == NB - 1) CL_MARK(11);
      const int np = I - r;
      cta_fetch_tiles(gtile(I, r), np, stg, t);      // L_IK, K = r .. I-1
      __syncthreads();
      {   // S = sum_K L_IK X_Kr  (row-major)
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        for (int w = 0; w < np; ++w) cta_tile_mac(acc, stg + w * TS, xcol + (r + w) * TS, tr, tc);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          *reinterpret_cast<float2*>(stile + (2 * tr + i) * TP + 2 * tc) = make_float2(acc[i][0], acc[i][1]);
      }
      if (I == NB - 1) CL_MARK(12);
      cl_wait(&flagB[I]);                            // inv(L_II) is in CTA I's shared memory (column-major)
      if (I == NB - 1) CL_MARK(13);
      {
        const int col = t >> 3, q8 = t & 7;
        const uint32_t rt = cl_map((uint32_t)__cvta_generic_to_shared(dinv), (uint32_t)I);
        const float4 v = cl_ld4(rt + (uint32_t)((col * TP + 4 * q8) * 4));
        *reinterpret_cast<float4*>(ljj + col * TP + 4 * q8) = v;
      }
      __syncthreads();
      {   // X_Ir = -inv(L_II) S
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        cta_tile_mac(acc, ljj, stile, tr, tc);
        float* Xi = xcol + I * TS;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          *reinterpret_cast<float2*>(Xi + (2 * tr + i) * TP + 2 * tc) = make_float2(-acc[i][0], -acc[i][1]);
      }
      __syncthreads();
      {   // row-major tile -> output rows, 128-bit coalesced
        const int row = t >> 3, q8 = t & 7;
        *reinterpret_cast<float4*>(Xg + (long long)(I * TB + row) * g + r * TB + 4 * q8) =
            *reinterpret_cast<const float4*>(xcol + I * TS + row * TP + 4 * q8);
      }
    }
  }
  CL_MARK(14);
  // nobody leaves while a peer may still arrive on its flags or read its shared memory; no data is handed over here, so
  // the arrival does not have to wait for this CTA's global stores (relaxed)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  CL_MARK(15);
}

size_t chol_cluster_scratch_floats(int g) {
  const int NB = g / TB;
  return (size_t)(NB * (NB + 1) / 2 + NB) * GT + (size_t)NB * TB;
}
inline bool chol_cluster_ok(int g, const float* Xout) { return Xout && (g % TB) == 0 && g >= 128 && g <= TB * CLN; }

}  // namespace

// scratch floats the factorisation needs besides its outputs: the padded inverse when g is not a multiple of 32, the
// published tiles of the cluster kernel when g >= 128
size_t wc_chol_pad_floats(int C, int g) {
  int NB = (g + TB - 1) / TB, gp = NB * TB;
  if (gp != g) return (size_t)(C / g) * gp * gp;
  return (g >= 128 && g <= TB * CLN) ? (size_t)(C / g) * chol_cluster_scratch_floats(g) : 0;
}

// Ain (G,g,g) SPD -> Lout (G,g,g) lower factor (nullable), Xout (G,g,g) = L^-1 (nullable). Xpad: scratch of
// wc_chol_pad_floats floats (only used when g is not a multiple of 32).
int wc_chol_inv_launch(const float* Ain, int C, int g, float* Lout, float* Xout, float* Xpad, int32_t* info, cudaStream_t st) {
  if (g > 256) return WC_E_UNSUPPORTED;
  const int NB = (g + TB - 1) / TB;
  if (chol_cluster_ok(g, Xout) && !(wc_debug_flags() & 2048)) {
    if (!Xpad) return WC_E_ARG;
    const size_t csm = (size_t)(2 * CLN + 5 + CLW) * TS * sizeof(float);
    static bool cattr = false;
    if (!cattr) {
      WC_CUDA(cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
      cattr = true;
    }
    WC_CUDA(wc_launch(chol_cluster_kernel, (C / g) * CLN, CLW * 32, csm, st, Ain, g, Xpad, (long long)chol_cluster_scratch_floats(g), Lout, Xout, info, wc_debug_flags() & 1));
    WC_CHECK_LAUNCH();
    return 0;
  }
  size_t smem = ((size_t)(NB * (NB + 1) / 2 + NB) * TS + (size_t)NB * TB) * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    WC_CUDA(cudaFuncSetAttribute(chol_inv_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  // g >= 128 (multiple of 32): the off-diagonal blocks of the inverse are formed by trinv_kernel on g/8 CTAs per group
  const bool split = Xout && Lout && g >= 128 && (g % TB) == 0;
  WC_CUDA(wc_launch(chol_inv_tiled_kernel, C / g, NW * 32, smem, st, Ain, g, Lout, Xout, Xpad, info, split ? 1 : 0));
  WC_CHECK_LAUNCH();
  if (split) {
    size_t tsm = ((size_t)TB * TI_LP + (size_t)TB * TI_DP + (size_t)TI_COLS * g + (size_t)TB * (TI_COLS + 1)) * sizeof(float);
    static size_t tattr = 0;
    if (tsm > 48 * 1024 && tsm > tattr) {
      WC_CUDA(cudaFuncSetAttribute(trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      tattr = tsm;
    }
    WC_CUDA(wc_launch(trinv_kernel, (C / g) * (g / TI_COLS), TI_THREADS, tsm, st, Lout, g, Xout, wc_debug_buf(), wc_debug_flags()));
    WC_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int wc_debug_chol_cluster_timing(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_cl_ts, sizeof(unsigned long long) * CLN * 32);
}
extern "C" int wc_debug_chol_timing(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_chol_ts, sizeof(long long) * 64);
}
