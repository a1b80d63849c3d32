// Standalone probe for the tcgen05 operand layouts / descriptors used by wc_tc_*.cu.
// Each case runs one CTA: fill smem operand tiles with the SAME address functions the product kernels use,
// issue the MMAs, read the accumulator back from TMEM and compare with a CPU product (inputs are small
// integers, so fp32 results are exact).   Build: nvcc -gencode arch=compute_100a,code=sm_100a umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../wc_gan_b200/csrc/wc_tc.cuh"
using namespace wctc;

struct Case {
  int mode;      // 0: K-major KB=64, 1: K-major KB=32, 2: MN-major (KROWS=32)
  int M, N, K;   // M = 128
  int a_blk_off; // MN-major: A operand starts at this 64-block of the shared X tile (gram-like); -1: separate A tile
  int products;  // 1 or 3 accumulating products (same operands) to test accumulation
};

template <int MODE>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ D,
                                            int N, int K, int a_blk_off, int products, int lda, int ldb) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  constexpr int KB = MODE == 0 ? 64 : 32;
  uint8_t* As = smem;
  uint8_t* Bs = smem + 64 * 1024;
  if (MODE <= 1) {
    // A [128][K], B [N][K], K-major tiles of KB
    for (int idx = t; idx < 128 * (K / 8); idx += 128) {
      int row = idx / (K / 8), ch = idx % (K / 8);
      int kb = (ch * 8) / KB, c = ch % (KB / 8);
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack_bf16x2(Ag[row * lda + ch * 8 + 2 * q], Ag[row * lda + ch * 8 + 2 * q + 1]);
      *(uint4*)(As + kb * (128 * KB * 2) + kmajor_chunk_off<KB>(row, c)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int idx = t; idx < N * (K / 8); idx += 128) {
      int row = idx / (K / 8), ch = idx % (K / 8);
      int kb = (ch * 8) / KB, c = ch % (KB / 8);
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack_bf16x2(Bg[row * ldb + ch * 8 + 2 * q], Bg[row * ldb + ch * 8 + 2 * q + 1]);
      *(uint4*)(Bs + kb * (N * KB * 2) + kmajor_chunk_off<KB>(row, c)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    // X tiles [k rows][cols], MN-major, KROWS = 32 per K block
    int acols = a_blk_off >= 0 ? 0 : 128;
    for (int idx = t; idx < K * (acols / 8); idx += 128) {
      int k = idx / (acols / 8), ch = idx % (acols / 8);
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack_bf16x2(Ag[k * lda + ch * 8 + 2 * q], Ag[k * lda + ch * 8 + 2 * q + 1]);
      *(uint4*)(As + (k / 32) * (2 * 32 * 128) + mnmajor_chunk_off<32>(ch / 8, k % 32, ch % 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int idx = t; idx < K * (N / 8); idx += 128) {
      int k = idx / (N / 8), ch = idx % (N / 8);
      uint32_t w[4];
      for (int q = 0; q < 4; ++q) w[q] = pack_bf16x2(Bg[k * ldb + ch * 8 + 2 * q], Bg[k * ldb + ch * 8 + 2 * q + 1]);
      *(uint4*)(Bs + (k / 32) * ((N / 64) * 32 * 128) + mnmajor_chunk_off<32>(ch / 8, k % 32, ch % 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_base;
  if (t == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, MODE == 2, MODE == 2);
    uint32_t acc = 0;
    for (int p = 0; p < products; ++p) {
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint64_t ad, bd;
        if (MODE <= 1) {
          int kb = k0 / KB, kin = k0 % KB;
          ad = make_desc(smem_u32(As + kb * (128 * KB * 2)) + kin * 2, 16, 8 * KB * 2, MODE == 0 ? LAYOUT_SW128 : LAYOUT_SW64);
          bd = make_desc(smem_u32(Bs + kb * (N * KB * 2)) + kin * 2, 16, 8 * KB * 2, MODE == 0 ? LAYOUT_SW128 : LAYOUT_SW64);
        } else {
          int kb = k0 / 32, kin = k0 % 32;
          uint32_t bbase = smem_u32(Bs + kb * ((N / 64) * 32 * 128)) + (kin / 8) * 1024;
          uint32_t abase = a_blk_off >= 0 ? bbase + a_blk_off * (32 * 128) : smem_u32(As + kb * (2 * 32 * 128)) + (kin / 8) * 1024;
          ad = make_desc(abase, 32 * 128, 1024, LAYOUT_SW128);
          bd = make_desc(bbase, 32 * 128, 1024, LAYOUT_SW128);
        }
        umma_bf16(td, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(td + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j) D[(warp * 32 + (t & 31)) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(td, 256);
}

int main() {
  Case cases[] = {{0, 128, 256, 64, -1, 1}, {0, 128, 128, 128, -1, 3}, {1, 128, 128, 32, -1, 1}, {1, 128, 256, 64, -1, 1},
                  {2, 128, 256, 32, -1, 1}, {2, 128, 128, 64, -1, 1}, {2, 128, 256, 64, 2, 1}, {2, 128, 256, 32, 0, 3}};
  int nfail = 0;
  for (auto& c : cases) {
    int M = 128, N = c.N, K = c.K;
    int lda, ldb;
    std::vector<float> A, B, ref((size_t)M * N, 0.f);
    srand(1234 + c.mode * 7 + N + K);
    auto rnd = []() { return (float)((rand() % 9) - 4); };
    if (c.mode <= 1) {
      lda = K; ldb = K; A.resize((size_t)M * K); B.resize((size_t)N * K);
      for (auto& v : A) v = rnd();
      for (auto& v : B) v = rnd();
      for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { float s = 0; for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k]; ref[(size_t)i * N + j] = s * c.products; }
    } else {
      lda = 128; ldb = N; A.resize((size_t)K * 128); B.resize((size_t)K * N);
      for (auto& v : A) v = rnd();
      for (auto& v : B) v = rnd();
      for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        float s = 0;
        for (int k = 0; k < K; ++k) { float a = c.a_blk_off >= 0 ? B[(size_t)k * N + c.a_blk_off * 64 + i] : A[(size_t)k * 128 + i]; s += a * B[(size_t)k * N + j]; }
        ref[(size_t)i * N + j] = s * c.products;
      }
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, (size_t)M * N * 4);
    size_t smem = 64 * 1024 + 128 * 1024 + 1024;
    if (c.mode == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<0><<<1, 128, smem>>>(dA, dB, dD, N, K, c.a_blk_off, c.products, lda, ldb); }
    if (c.mode == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<1><<<1, 128, smem>>>(dA, dB, dD, N, K, c.a_blk_off, c.products, lda, ldb); }
    if (c.mode == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<1, 128, smem>>>(dA, dB, dD, N, K, c.a_blk_off, c.products, lda, ldb); }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D((size_t)M * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - ref[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr || d != d) maxerr = d; }
    printf("case mode=%d N=%d K=%d a_blk=%d prod=%d : %s  cuda=%s maxerr=%g bad=%d/%zu  D[0..3]=%g %g %g %g ref=%g %g %g %g\n", c.mode, N, K, c.a_blk_off,
           c.products, bad == 0 && e == cudaSuccess ? "OK" : "FAIL", cudaGetErrorString(e), maxerr, bad, D.size(), D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    if (bad || e != cudaSuccess) ++nfail;
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    if (e != cudaSuccess) { printf("aborting after CUDA error\n"); break; }
  }
  printf("%s\n", nfail ? "PROBE FAILED" : "PROBE PASSED");
  return nfail != 0;
}
