// Shared helpers for the libwc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/wc_b200.h"

// statistics only (bench.py's gpu_launches): kernels launched by this library since load
extern "C" long long wc_launch_count(void);
void wc_count_launch();

#define WC_CHECK_LAUNCH()                                   \
  do {                                                      \
    wc_count_launch();                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return (int)e__;                \
  } while (0)

#define WC_CUDA(call)                                       \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return (int)e__;                \
  } while (0)

// ---- launches with programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of this library starts with wc_pdl_wait() before it touches global memory and is launched with the
// programmatic-stream-serialization attribute: the grid may be scheduled (and run its prologue) while its predecessor
// on the stream is still draining, and blocks in wc_pdl_wait() until that predecessor has completed and its writes are
// visible.  A layer step is ~20 dependent launches; this hides their launch latency.  wc_debug_set bit 8192 turns it off.
__device__ __forceinline__ void wc_pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool wc_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t wc_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = wc_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same with a thread-block cluster of `cluster` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t wc_launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = wc_pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- debug-only instrumentation (not part of include/wc_b200.h) ---------------------------------------
// wc_debug_set(flags): bit0 = kernels record globaltimer marks into a device buffer [block][slot] (allocated by
// wc_debug_set, the only place this library allocates); other bits are kernel-specific experiment switches.
constexpr int WC_DBG_BLOCKS = 160, WC_DBG_SLOTS = 32;   // slots 0-7: streaming kernels, 8-15: chain kernel stages, 16-31: inside its last tile
int wc_debug_flags();
unsigned long long* wc_debug_buf();     // nullptr unless bit0 is set
__device__ __forceinline__ void wc_mark(unsigned long long* buf, int slot) {
  if (buf && blockIdx.x < WC_DBG_BLOCKS) {
    unsigned long long tns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
    buf[blockIdx.x * WC_DBG_SLOTS + slot] = tns;
  }
}

// L2 residency policy of the four streaming kernels (bit set = hint used): 1 K1 loads evict_last, 2 K3 loads evict_first,
// 4 K3 stores evict_first, 8 K4 loads evict_last, 16 K6 loads evict_first, 32 K6 stores evict_first.
// Default WC_L2_DEFAULT; the environment variable WC_B200_L2 (read once) overrides it for A/B measurements.
constexpr int WC_L2_DEFAULT = 0;
int wc_l2_mode();

// cheap SM-local cycle stamp (globaltimer reads cost hundreds of ns each and perturb what they measure)
__device__ __forceinline__ void wc_mark_clk(unsigned long long* buf, int slot) {
  if (buf && blockIdx.x < WC_DBG_BLOCKS) buf[blockIdx.x * WC_DBG_SLOTS + slot] = (unsigned long long)clock64();
}

static inline bool wc_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline size_t wc_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int wc_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int WC_NUM_SMS = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float wc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not allocate in L1 (activations are read once per kernel)
__device__ __forceinline__ float4 wc_ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---- internal launchers shared between translation units ---------------------------------
// generic batched small matmul:  Cm[b1,b2] = alpha * sum_{q<nsum} opA opB + beta * Cm
struct WcSmm {
  const float* A; const float* B; float* C;
  int M, N, K;                 // C is M x N, inner dimension K
  long long rsA, csA;          // element (i,k) of op(A) at A[i*rsA + k*csA]
  long long rsB, csB;          // element (k,j) of op(B) at B[k*rsB + j*csB]
  long long rsC, csC;
  int nb1, nb2;                // two batch dimensions (grid.z = nb1*nb2)
  long long bA1, bA2, bB1, bB2, bC1, bC2;
  int nsum;                    // extra summed dimension
  long long sA, sB;
  float alpha, beta;
  int tri;                     // 0 none, 1 keep lower (incl. diag) of C, 2 keep upper (incl. diag); others zeroed
};
int wc_smm_launch(const WcSmm& p, cudaStream_t st);

// K1's finalize kernel can run K2's prep step in the same launch (single-process training, tensor-core path)
struct WcPrepFuse { float eps, extra_eps, momentum; float* moving_mean; float* moving_cov; float* Ashr; };
// raw accumulators of the tensor-core K1 (lower blocks of the pivot-shifted scatter, column sums, pivot): set by
// wc_stats_fwd_impl when asked to leave the finalize step to the consumer (the peer-memory exchange kernel)
struct WcRawStats { const float* partS; const float* psum; const float* pivot; };
// zeroed_out (nullable): receives the address of 64 zeroed bytes inside the workspace (cleared by the accumulator memset of
// the tensor-core path; NULL when that path is not taken) -- the grid-barrier counter of the forward chain launch
int wc_stats_fwd_impl(const float* x, int64_t M, int C, int g, float* workspace, size_t workspace_bytes,
                      float* mean, float* S, const WcPrepFuse* fuse, int* fused, unsigned** zeroed_out, WcRawStats* raw_out,
                      cudaStream_t stream);
int wc_factor_fwd_impl(const float* mean, const float* S, const float* count_dev, double count_host, int C, int g,
                       float eps, float momentum, int decomposition, int renorm, int training,
                       float* moving_mean, float* moving_cov, float* Weff, float* Wm, float* R, float* U,
                       float* lam, int32_t* info, float* workspace, size_t workspace_bytes, int prep_done, cudaStream_t stream);

// stage functions with the operand packing of the tcgen05 apply kernels folded in (wc_small.cu; used by wc_layer.cu)
int wc_compose_fwd_impl(const float* Weff, int Cin, int Cout, int g, int K, const float* Wu, const float* bu,
                        const float* Wc, const float* bc, int triangular, const float* class_matrix,
                        const float* D, int E, const float* bf, float* ehat, float* Wk, float* T, float* b,
                        void* Tp_fwd, void* Tp_bwd, const float* Wk_ready, unsigned* barrier, float* fact_ws, cudaStream_t st);
size_t wc_compose_fact_workspace_bytes(int C, int K, int E, bool has_class_matrix, bool has_wc);
int wc_factor_bwd_impl(const float* dWeff, const float* gbar, const float* Weff, const float* Wm, const float* R,
                       const float* U, const float* lam, const float* count_dev, double count_host, int C, int g,
                       float eps, int decomposition, int renorm, float* A, float* cvec, float* workspace,
                       size_t workspace_bytes, void* Tp_bwd, int pack_K, cudaStream_t st);

int wc_reduce_bwd_impl(const float* x, const float* dy, int N, int HW, int Cin, int Cout, const float* mean, const int32_t* cls,
                       int K, float* P, float* s, float* workspace, size_t workspace_bytes, int zero_tail_floats,
                       float* zero2, size_t zero2_floats, float* zero3, size_t zero3_floats, cudaStream_t stream);

// K5a + K5b as one persistent multi-stage launch (wc_small.cu); WC_E_UNSUPPORTED: use the per-product launches
bool wc_bwd_factorised_ok(int C, int K, int E, bool has_class_matrix, bool has_wc);
int wc_bwd_chain(const float* P, const float* s, const float* Weff, const float* Wk, const float* Wu, int Cin, int Cout, int g, int K,
                 float* dWu, float* dbu, float* dWc, float* dbc, int triangular, const float* class_matrix, const float* ehat,
                 const float* D, int E, float* dD, float* dclass_matrix, float* dbf, float* red, float* A, float* cvec,
                 const float* count_dev, double count_host, float eps, void* Tp_bwd, int pack_K, float* ws_color,
                 float* ws_factor, unsigned* barrier, int stage_lo, int stage_hi, void* const* xchg_peers, void* xchg_state,
                 int xchg_rank, int xchg_world, int32_t* info, int wk_missing, int dweff_zeroed, int ws_zeroed, cudaStream_t st);
int wc_rebuild_wk(int Cin, int Cout, int K, const float* Wu, const float* class_matrix, const float* D, int E, float* Wk, cudaStream_t st);

int wc_apply_fwd_impl(const float* x, int N, int HW, int Cin, int Cout, const float* mean, const float* T, const void* T_packed,
                      const float* b, const int32_t* cls, int K, int band, float* y, int32_t* info, int relu, cudaStream_t stream, int early = 0);
int wc_apply_bwd_impl(const float* x, const float* dy, int N, int HW, int Cin, int Cout, const float* mean,
                      const float* T, const float* A, const void* TA_packed, const float* cvec, int g,
                      const int32_t* cls, int K, float* dx, cudaStream_t stream, int early);

// peer-memory exchanges (wc_exchange.cu)
int wc_stats_exchange(float* rec, float count_local, int C, int g, void* const* peers, void* state, int rank, int world, float eps,
                      float extra_eps, float momentum, float* moving_mean, float* moving_cov, float* Ashr, int32_t* info,
                      const float* partS, const float* psum, const float* pivot, double M, cudaStream_t st);
int wc_adjoint_exchange(float* red, int C, int g, void* const* peers, void* state, int rank, int world, int32_t* info,
                        cudaStream_t st);

// tensor-map encoder for the TMA kernels (wc_tc_apply.cu)
int wc_make_f32_map3(void* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1);

// Jacobi eigensolvers of the ZCA path (wc_jacobi.cu)
size_t wc_jacobi_sync_bytes(int C, int g);
int wc_jacobi_launch(const float* L, int C, int g, float* Bm, float* U, float* lam, unsigned* sync, int32_t* info, cudaStream_t st);

// tiled Cholesky + inverse (wc_chol.cu)
size_t wc_chol_pad_floats(int C, int g);
int wc_chol_inv_launch(const float* Ain, int C, int g, float* Lout, float* Xout, float* Xpad, int32_t* info, cudaStream_t st);
