// K3 (fused centre + whiten + colour) and K6 (backward apply), FP32 FFMA path (any shape).
//   out[n][r, j] = sum_k (A1[n][r,k] - m1[k]) B1[slot(n)][k, j]
//                + sum_k (A2[n][r,k] - m2[k]) B2[k, j]            (optional second term)
//                + sign * bias[slot(n)][j]
// K3: A1 = x, m1 = mean, B1 = T_k, bias = b_k.
// K6: A1 = dy, B1 = T_k^T, A2 = x, m2 = mean, B2 = A (block diagonal), bias = -cvec.
// The tcgen05 path for C in {128,256} lives in wc_tc.cu.
//
// Reference segments replaced: gan/conditional_layers.py:660,:706-712,:798-846,:979-1029,
// generator.py:50,:58,:76 (K3) and TF autodiff of the same w.r.t. the activations (K6).
#include "wc_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, KC = 16, NT = 256, PAD = 4;

struct Term {
  const float* A; int lda; const float* m;       // activations [rows][lda], per-column shift (nullable)
  const float* B; long long slot_stride;         // matrix per slot (0 = shared)
  long long rsB, csB;                            // element (k,j) at B[k*rsB + j*csB]
  int Kdim; int band;                            // inner dimension; >0: block diagonal with square blocks of `band`
  int compact;                                   // 1: B holds only the diagonal blocks, (G,band,band) row-major
};

struct ApplyParams {
  Term t1, t2; int nterms;
  const float* bias; long long bias_slot_stride; float bias_sign;
  int N, HW, Nout;
  const int32_t* cls; int K; int32_t* info;
  float* out; int ldo;
  int tiles_per_sample;
  int relu;                                      // 1: out = max(0, .)  (the activation that follows the WC layer, fused)
};

template <bool kVec>
__device__ __forceinline__ void run_term(const Term& T, int slot, long long row_base, int rows_valid, int j0, int jend,
                                         float (&acc)[4][4], float (*As)[TM + PAD], float (*Bs)[TN + PAD]) {
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const float* Bm = T.B + (long long)slot * T.slot_stride;
  int klo = 0, khi = T.Kdim;
  if (T.band > 0) {
    klo = (j0 / T.band) * T.band;
    khi = min(T.Kdim, ((jend - 1) / T.band + 1) * T.band);
  }
  // loader mappings
  const int arow = t >> 2, ak = (t & 3) * 4;      // A: 64 rows x 16 k, float4 along k
  const int bk = t >> 4, bj = (t & 15) * 4;       // B row-major (csB == 1): 16 k x 64 j, float4 along j
  const int bjt = t >> 2, bkt = (t & 3) * 4;      // B k-contiguous (rsB == 1): float4 along k
  const bool b_rowmajor = (T.csB == 1), b_kmajor = (T.rsB == 1) && !b_rowmajor;

  for (int k0 = klo; k0 < khi; k0 += KC) {
    // ---- global -> registers
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (arow < rows_valid) {
      const float* ap = T.A + (row_base + arow) * T.lda + k0 + ak;
      if (kVec && k0 + ak + 3 < khi) {
        float4 v = wc_ldg_stream(reinterpret_cast<const float4*>(ap));
        av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        if (T.m) { float4 s = *reinterpret_cast<const float4*>(T.m + k0 + ak); av[0] -= s.x; av[1] -= s.y; av[2] -= s.z; av[3] -= s.w; }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (k0 + ak + q < khi) av[q] = __ldg(ap + q) - (T.m ? T.m[k0 + ak + q] : 0.f);
      }
    }
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (T.compact) {
      // block-diagonal matrix stored as (G,band,band): element (k,j) is non-zero only inside one block
      int k = k0 + bk, j = j0 + bj;
      if (k < khi) {
        int gk = k / T.band;
        const float* bp = Bm + (long long)gk * T.band * T.band + (long long)(k - gk * T.band) * T.band;
        if (kVec && j + 3 < jend) {
          int gj = j / T.band;
          if (gj == gk) { float4 v = *reinterpret_cast<const float4*>(bp + (j - gj * T.band)); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int gj = (j + q) / T.band;
            if (j + q < jend && gj == gk) bv[q] = bp[j + q - gj * T.band];
          }
        }
      }
    } else if (b_rowmajor) {
      int k = k0 + bk, j = j0 + bj;
      if (k < khi) {
        const float* bp = Bm + (long long)k * T.rsB + j;
        if (kVec && j + 3 < jend) { float4 v = *reinterpret_cast<const float4*>(bp); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (j + q < jend) bv[q] = bp[q];
        }
      }
    } else {
      int j = j0 + bjt, k = k0 + bkt;
      if (j < jend) {
        const float* bp = Bm + (long long)j * T.csB + (long long)k * T.rsB;
        if (kVec && b_kmajor && k + 3 < khi) { float4 v = *reinterpret_cast<const float4*>(bp); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k + q < khi) bv[q] = bp[(long long)q * T.rsB];
        }
      }
    }
    __syncthreads();   // previous chunk fully consumed
#pragma unroll
    for (int q = 0; q < 4; ++q) As[ak + q][arow] = av[q];
    if (b_rowmajor || T.compact) {
      *reinterpret_cast<float4*>(&Bs[bk][bj]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[bkt + q][bjt] = bv[q];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a4[4] = {a.x, a.y, a.z, a.w}, b4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a4[x], b4[y], acc[x][y]);
    }
  }
}

template <bool kVec>
__global__ void __launch_bounds__(NT) apply_kernel(ApplyParams p) {
  wc_pdl_wait();
  __shared__ __align__(16) float As[KC][TM + PAD];
  __shared__ __align__(16) float Bs[KC][TN + PAD];
  const int sample = blockIdx.x / p.tiles_per_sample;
  const int rt = blockIdx.x % p.tiles_per_sample;
  const int r0 = rt * TM;
  const int rows_valid = min(TM, p.HW - r0);
  const long long row_base = (long long)sample * p.HW + r0;
  const int j0 = blockIdx.y * TN, jend = min(p.Nout, j0 + TN);
  int slot = 0;
  if (p.cls) {
    int c = p.cls[sample];
    if (c < 0 || c >= p.K) {
      if (p.info && threadIdx.x == 0 && rt == 0 && blockIdx.y == 0) atomicMax(&p.info[WC_INFO_BAD_CLASS], sample + 1);
      c = c < 0 ? 0 : p.K - 1;
    }
    slot = c;
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  run_term<kVec>(p.t1, slot, row_base, rows_valid, j0, jend, acc, As, Bs);
  if (p.nterms > 1) run_term<kVec>(p.t2, 0, row_base, rows_valid, j0, jend, acc, As, Bs);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j = j0 + tx * 4;
  float bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
    const float* bp = p.bias + (long long)slot * p.bias_slot_stride;
#pragma unroll
    for (int q = 0; q < 4; ++q) if (j + q < jend) bb[q] = p.bias_sign * bp[j + q];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int r = ty * 4 + a;
    if (r >= rows_valid) continue;
    float* op = p.out + (row_base + r) * p.ldo + j;
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { o[q] = acc[a][q] + bb[q]; if (p.relu) o[q] = fmaxf(o[q], 0.f); }
    if (kVec && j + 3 < jend) {
      *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (j + q < jend) op[q] = o[q];
    }
  }
}

int launch_apply(ApplyParams& p, bool vec, cudaStream_t st) {
  const bool cond = p.cls != nullptr && p.K > 1;
  if (!cond) { p.HW = p.N * p.HW; p.N = 1; p.cls = nullptr; }   // one virtual sample => full row tiles
  p.tiles_per_sample = wc_cdiv(p.HW, TM);
  long long gx = (long long)p.N * p.tiles_per_sample;
  if (gx > INT32_MAX) return WC_E_UNSUPPORTED;
  dim3 grid((unsigned)gx, wc_cdiv(p.Nout, TN));
  if (vec) WC_CUDA(wc_launch(apply_kernel<true>, grid, NT, 0, st, p)); else WC_CUDA(wc_launch(apply_kernel<false>, grid, NT, 0, st, p));
  WC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int wc_tc_apply_fwd(const float* x, int N, int HW, int C, const float* mean, const void* packed, const float* b,
                    const int32_t* cls, int K, float* y, int32_t* info, int relu, int early, cudaStream_t st);
int wc_tc_apply_bwd(const float* x, const float* dy, int N, int HW, int C, const float* mean, const void* packed,
                    bool has_A, const float* cvec, const int32_t* cls, int K, float* dx, int early, cudaStream_t st);
bool wc_tc_apply_ok(int Cin, int Cout);

extern "C" int wc_apply_fwd(const float* x, int N, int HW, int Cin, int Cout, const float* mean, const float* T,
                            const void* T_packed, const float* b, const int32_t* cls, int K, int band, float* y,
                            int32_t* info, wc_stream_t stream) {
  return wc_apply_fwd_impl(x, N, HW, Cin, Cout, mean, T, T_packed, b, cls, K, band, y, info, 0, (cudaStream_t)stream, 0);
}

// relu != 0: y = max(0, .), the Activation('relu') that follows every WC layer (gan/layer_utils.py:170,178; generator.py:155)
int wc_apply_fwd_impl(const float* x, int N, int HW, int Cin, int Cout, const float* mean, const float* T, const void* T_packed,
                      const float* b, const int32_t* cls, int K, int band, float* y, int32_t* info, int relu, cudaStream_t stream,
                      int early) {
  if (!x || !T || !y || N <= 0 || HW <= 0 || Cin <= 0 || Cout <= 0 || K <= 0) return WC_E_ARG;
  if ((int64_t)N * HW > (int64_t)INT32_MAX) return WC_E_UNSUPPORTED;
  if (T_packed && wc_tc_apply_ok(Cin, Cout) && wc_aligned16(x) && wc_aligned16(y) && wc_aligned16(T_packed) &&
      (!mean || wc_aligned16(mean)))
    return wc_tc_apply_fwd(x, N, HW, Cin, mean, T_packed, b, cls, K, y, info, relu, early, (cudaStream_t)stream);
  ApplyParams p{};
  p.t1 = Term{x, Cin, mean, T, (long long)Cin * Cout, Cout, 1, Cin, band, 0};
  p.nterms = 1;
  p.bias = b; p.bias_slot_stride = Cout; p.bias_sign = 1.f;
  p.N = N; p.HW = HW; p.Nout = Cout; p.cls = cls; p.K = K; p.info = info; p.out = y; p.ldo = Cout; p.relu = relu;
  bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && (band <= 0 || band % 4 == 0) && wc_aligned16(x) && wc_aligned16(T) &&
             wc_aligned16(y) && (!mean || wc_aligned16(mean));
  return launch_apply(p, vec, (cudaStream_t)stream);
}

extern "C" int wc_apply_bwd(const float* x, const float* dy, int N, int HW, int Cin, int Cout, const float* mean,
                            const float* T, const float* A, const void* TA_packed, const float* cvec, int g,
                            const int32_t* cls, int K, float* dx, wc_stream_t stream) {
  return wc_apply_bwd_impl(x, dy, N, HW, Cin, Cout, mean, T, A, TA_packed, cvec, g, cls, K, dx, (cudaStream_t)stream, 0);
}
// early: TcApplyParams::early (whole-layer calls)
int wc_apply_bwd_impl(const float* x, const float* dy, int N, int HW, int Cin, int Cout, const float* mean,
                      const float* T, const float* A, const void* TA_packed, const float* cvec, int g,
                      const int32_t* cls, int K, float* dx, cudaStream_t stream, int early) {
  if (!dy || !T || !dx || N <= 0 || HW <= 0 || Cin <= 0 || Cout <= 0 || K <= 0) return WC_E_ARG;
  if (A && (!x || g <= 0 || Cin % g)) return WC_E_ARG;
  if ((int64_t)N * HW > (int64_t)INT32_MAX) return WC_E_UNSUPPORTED;
  if (TA_packed && wc_tc_apply_ok(Cin, Cout) && wc_aligned16(dy) && wc_aligned16(dx) && wc_aligned16(TA_packed) &&
      (!A || (wc_aligned16(x) && (!mean || wc_aligned16(mean)))))
    return wc_tc_apply_bwd(x, dy, N, HW, Cin, mean, TA_packed, A != nullptr, cvec, cls, K, dx, early, (cudaStream_t)stream);
  ApplyParams p{};
  // term 1: dy (rows x Cout) times T^T (Cout x Cin): element (k=jout, j=iin) = T[iin*Cout + jout]
  p.t1 = Term{dy, Cout, nullptr, T, (long long)Cin * Cout, 1, Cout, Cout, 0, 0};
  p.nterms = 1;
  if (A) {
    // term 2: (x - mean) (rows x Cin) times blockdiag(A), A stored compactly as (G,g,g)
    p.t2 = Term{x, Cin, mean, A, 0, g, 1, Cin, g, 1};
    p.nterms = 2;
  }
  p.bias = A ? cvec : nullptr; p.bias_slot_stride = 0; p.bias_sign = -1.f;
  p.N = N; p.HW = HW; p.Nout = Cin; p.cls = cls; p.K = K; p.info = nullptr; p.out = dx; p.ldo = Cin;
  bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && (g % 4 == 0) && wc_aligned16(dy) && wc_aligned16(T) && wc_aligned16(dx) &&
             (!A || (wc_aligned16(x) && wc_aligned16(A) && (!mean || wc_aligned16(mean))));
  return launch_apply(p, vec, (cudaStream_t)stream);
}
