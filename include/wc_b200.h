/*
 * wc_b200.h — C ABI of libwc_b200.so: the whitening-coloring (WC) layer hot path of
 * AliaksandrSiarohin/wc-gan as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI (it is pure Python on TensorFlow 1.5 ops), so each entry point
 * below names the reference graph segment it replaces (file:line under /root/reference).
 * A maintainer binds these with ctypes exactly as wc_gan_b200/_lib.py does; see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - activations are fp32, NHWC, contiguous:  x[(n*HW + p)*C + c];  M = N*HW rows;
 *   - "group whitening": channels are split into G = C/g contiguous groups, every
 *     (G,g,g) array holds one row-major g x g matrix per group (g == C is the reference);
 *   - per-class arrays are indexed by the int32 label cls[n] in [0,K); K == 1 (cls == NULL)
 *     is the unconditional layer;
 *   - the caller owns every buffer including workspaces; the compute path keeps no mutable
 *     global state, never allocates, never synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).  The only process-wide switches are diagnostics that
 *     default to off/automatic: wc_set_tensor_core_gram below and include/wc_b200_debug.h;
 *   - return value: 0 ok; <0 invalid argument / unsupported shape (WC_E_*); >0 a
 *     cudaError_t from the launch.
 */
#ifndef WC_B200_H_
#define WC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WC_B200_VERSION 102

#define WC_E_ARG        (-1)  /* null pointer / non-positive size / C % g != 0 */
#define WC_E_WORKSPACE  (-2)  /* workspace too small */
#define WC_E_UNSUPPORTED (-3) /* shape outside what the kernels implement */
#define WC_E_ALIGN      (-4)  /* pointer not 16-byte aligned */

#define WC_DECOMP_CHOLESKY 0
#define WC_DECOMP_ZCA      1

/* info[] slots written by the device (int32[WC_INFO_SLOTS], caller zero-initialises) */
#define WC_INFO_SLOTS      4
#define WC_INFO_NOT_PD     0  /* 1 + index of the first non-positive Cholesky pivot (TF raises InvalidArgumentError; gan/train.py:109-112 swallows it) */
#define WC_INFO_BAD_CLASS  1  /* 1 + first sample index whose label is outside [0,K) (label is clamped, never read out of bounds) */
#define WC_INFO_JACOBI     2  /* sweeps the ZCA eigensolver used */
#define WC_INFO_XCHG_TIMEOUT 3 /* synchronised batch over peer memory: 1 + rank of a peer whose record never arrived within
                                  WC_B200_XCHG_TIMEOUT_MS (default 10000); the step's results are then undefined */

typedef void* wc_stream_t;

int wc_version(void);
/* sizeof(WcLayerCall) / sizeof(WcSavedLayout) as compiled: a binding checks its own struct mirror against these */
size_t wc_sizeof_layer_call(void);
size_t wc_sizeof_saved_layout(void);

/* Statistics only: number of kernels this library has launched since it was loaded
 * (a relaxed atomic counter; bench.py reports the per-step delta as gpu_launches).        */
long long wc_launch_count(void);

/* Which implementation the dispatcher would choose for a shape: bit0 = tcgen05 tensor-core
 * apply kernels, bit1 = tcgen05 gram kernels; 0 = FP32 FFMA kernels only.                 */
int wc_path_for(int64_t M, int Cin, int Cout, int g);

/* Process-wide switch (default on): 0 makes K1/K4 use the FP32 FFMA gram kernels even for shapes the
 * tcgen05 gram kernels support.  For A/B accuracy tests; not needed in normal use.               */
void wc_set_tensor_core_gram(int enable);

/* ---- K1  batch statistics ----------------------------------------------------------------
 * Replaces gan/conditional_layers.py:652-660 (transpose, reshape, reduce_mean, subtract) and
 * :683 (f f^T) for one shard.  Output: mean[C] and the CENTRED scatter S (G,g,g) =
 * sum_rows (x-mean)(x-mean)^T (not yet divided by M-1, so shards can be merged).           */
size_t wc_stats_workspace_bytes(int64_t M, int C, int g);
int wc_stats_fwd(const float* x, int64_t M, int C, int g,
                 float* workspace, size_t workspace_bytes,
                 float* mean, float* S, wc_stream_t stream);

/* Merge R shard records {mean[C], S[G*g*g], count} (packed, stride C+G*g*g+1 floats), e.g.
 * the output of an all-gather, with Chan's parallel update.  New in this build (the
 * reference is single-GPU); equals :658,:683 on the concatenated batch.                   */
int wc_stats_merge(const float* records, int R, int C, int g,
                   float* mean, float* S, float* count_total, wc_stream_t stream);

/* ---- K2  factorisation -------------------------------------------------------------------
 * Replaces gan/conditional_layers.py:683 (/(M-1)), :684-690 (moving averages), :691 (shrink),
 * :662-678 (cholesky + triangular solve | svd on /cpu:0), :693-698 (renorm), :702-704 (test).
 * count = rows in the (global) batch, read from count_dev[0] when non-NULL else count_host.
 * training != 0: factor S/(count-1), update moving_mean/moving_cov in place (EMA of the
 *   un-shrunk covariance).  training == 0: factor moving_cov (mean, S ignored; the caller
 *   centres with moving_mean).
 * Outputs (each (G,g,g) unless noted): Weff = matrix applied to centred rows; Wm = batch
 *   inverse square root (== Weff unless renorm); R (renorm only) = stop-gradient factor with
 *   Weff = R Wm; U, lam[C] (ZCA only) = eigenvectors / eigenvalues of the shrunk matrix.   */
size_t wc_factor_workspace_bytes(int C, int g, int decomposition, int renorm);
int wc_factor_fwd(const float* mean, const float* S, const float* count_dev, double count_host,
                  int C, int g, float eps, float momentum,
                  int decomposition, int renorm, int training,
                  float* moving_mean, float* moving_cov,
                  float* Weff, float* Wm, float* R, float* U, float* lam,
                  int32_t* info, float* workspace, size_t workspace_bytes, wc_stream_t stream);

/* ---- K2b  compose the per-class affine maps ------------------------------------------------
 * Replaces generator.py:41-78 (the uconv / cconv / ucconv / fconv / ufconv stacks and their
 * Add), gan/conditional_layers.py:811-819 (triangular mask + gather), :993-1005 (embedding
 * gather, l2_normalize, dictionary contraction).  For every class k:
 *   W_k = Wu + triu?(Wc_k) + sum_e ehat[k,e] D_e,   b_k = bu + bc_k + bf_k,
 *   T_k = Weff^T W_k  (Weff == NULL: T_k = W_k).
 * Any of Wu/bu/Wc/bc/class_matrix+D/bf may be NULL.  Wk, T: (K,Cin,Cout); b: (K,Cout);
 * ehat: (K,E).  Weff is block diagonal over Cin with block g.                              */
int wc_compose_fwd(const float* Weff, int Cin, int Cout, int g, int K,
                   const float* Wu, const float* bu,
                   const float* Wc, const float* bc, int triangular,
                   const float* class_matrix, const float* D, int E, const float* bf,
                   float* ehat, float* Wk, float* T, float* b, wc_stream_t stream);

/* ---- K3  fused centre + whiten + colour ----------------------------------------------------
 * Replaces gan/conditional_layers.py:660,:706-712 (centre, whiten, relayout), generator.py:50
 * (Conv2D 1x1), gan/conditional_layers.py:798-846 and :979-1029 (batched matmul + bias +
 * relayouts):   y[n] = (x[n] - mean) T[cls[n]] + b[cls[n]].
 * mean, b, cls may be NULL (no centring / no bias / class 0).  band > 0 promises T is block
 * diagonal with square blocks of `band` (whitening only).                                  */
int wc_apply_fwd(const float* x, int N, int HW, int Cin, int Cout,
                 const float* mean, const float* T, const void* T_packed, const float* b,
                 const int32_t* cls, int K, int band,
                 float* y, int32_t* info, wc_stream_t stream);

/* ---- operand packing for the tcgen05 path (wc_path_for(...) & 1) --------------------------------
 * The tensor-core K3/K6 kernels read the C x C matrices as pre-split bf16 hi/lo operand images
 * (K-major, SWIZZLE_64B, one image per 32-channel K block) that cp.async.bulk drops straight into
 * shared memory.  wc_pack_fwd packs T (K,C,C) for K3; wc_pack_bwd packs T for K6's dy.T^T term and,
 * when A != NULL, the block-diagonal A (G,g,g) for its (x-mean).A term (nseg_shared = 1).
 * Passing the packed buffer to wc_apply_fwd / wc_apply_bwd selects the tensor-core kernels; NULL
 * selects the FP32 FFMA kernels (any shape).                                                      */
size_t wc_pack_bytes(int C, int K, int nseg_shared);
int wc_pack_fwd(const float* T, int C, int K, void* packed, wc_stream_t stream);
int wc_pack_bwd(const float* T, const float* A, int C, int g, int K, void* packed, wc_stream_t stream);

/* ---- K4  backward reduction ----------------------------------------------------------------
 * First half of TF autodiff through the segment K3 replaces:  per class
 *   P[k] = sum_{n: cls[n]=k} (x[n]-mean)^T dy[n]   (K,Cin,Cout),   s[k] = sum rows dy[n]  (K,Cout). */
size_t wc_reduce_workspace_bytes(int N, int HW, int Cin, int Cout, int K);
int wc_reduce_bwd(const float* x, const float* dy, int N, int HW, int Cin, int Cout,
                  const float* mean, const int32_t* cls, int K,
                  float* P, float* s, float* workspace, size_t workspace_bytes,
                  wc_stream_t stream);

/* ---- K5a  coloring-parameter gradients and the whitening-matrix adjoint ---------------------
 * Autodiff of K2b.  With Q_k = Weff P_k:
 *   dWu = sum_k Q_k, dbu = sum_k s_k, dWc[k] = triu?(Q_k), dbc = dbf = s,
 *   dD[e] = sum_k ehat[k,e] Q_k, dclass_matrix = l2_normalize' (<Q_k, D_e>),
 *   dWeff (G,g,g) = diag blocks of sum_k W_k P_k^T,  gbar[Cin] = sum_k W_k s_k.
 * Output pointers may be NULL when the gradient is not wanted.  dWeff and gbar are what a
 * synchronised-batch layer all-reduces before K5b.                                         */
size_t wc_color_bwd_workspace_bytes(int Cin, int Cout, int g, int K, int E);
int wc_color_bwd(const float* P, const float* s, const float* Weff, const float* Wk,
                 int Cin, int Cout, int g, int K,
                 float* dWu, float* dbu, float* dWc, float* dbc, int triangular,
                 const float* class_matrix, const float* ehat, const float* D, int E,
                 float* dD, float* dclass_matrix, float* dbf,
                 float* dWeff, float* gbar,
                 float* workspace, size_t workspace_bytes, wc_stream_t stream);

/* ---- K5b  adjoint of the factorisation ------------------------------------------------------
 * Autodiff of K2 (TF _CholeskyGrad / _MatrixTriangularSolveGrad / _SvdGrad, un-vendored):
 *   A (G,g,g) = 2(1-eps) sym(dSigma')/(count-1),   cvec[C] = Weff^T gbar / count
 * so that dx = dy T^T + (x-mean) A - cvec.                                                 */
size_t wc_factor_bwd_workspace_bytes(int C, int g, int decomposition, int renorm);
int wc_factor_bwd(const float* dWeff, const float* gbar,
                  const float* Weff, const float* Wm, const float* R,
                  const float* U, const float* lam,
                  const float* count_dev, double count_host,
                  int C, int g, float eps, int decomposition, int renorm,
                  float* A, float* cvec,
                  float* workspace, size_t workspace_bytes, wc_stream_t stream);

/* ---- K6  backward apply ----------------------------------------------------------------------
 * Second half of autodiff through K3/K1:  dx[n] = dy[n] T[cls[n]]^T + (x[n]-mean) A - cvec.
 * A == NULL (inference phase or no whitening): dx[n] = dy[n] T[cls[n]]^T.                  */
int wc_apply_bwd(const float* x, const float* dy, int N, int HW, int Cin, int Cout,
                 const float* mean, const float* T, const float* A, const void* TA_packed,
                 const float* cvec, int g,
                 const int32_t* cls, int K, float* dx, wc_stream_t stream);


/* ---- whole-layer calls ------------------------------------------------------------------------------
 * One call per direction enqueues every stage above (K1, K2, K2b, pack, K3 | K4, K5a, K5b, pack, K6) on
 * `stream`: what DecorelationNormalization.call + the coloring layer's call (gan/conditional_layers.py:646-712,
 * :790-848, :971-1034, generator.py:41-88) and TF's autodiff of them do per training step.  The host fills one
 * WcLayerCall; the library carves `saved` (written by forward, read by backward) and `workspace` (scratch of the
 * running call) itself.  Synchronised-batch whitening splits each direction in two phases around the host's
 * collective (WC_PHASE_*); a single GPU runs both phases in one call.                                      */
#define WC_PHASE_STATS   1   /* forward: K1 -> shard record {mean[C], S[G g g], count} at saved + layout.rec       */
#define WC_PHASE_APPLY   2   /* forward: K2, K2b, pack, K3 (statistics from the record or from mean_ext / S_ext)     */
#define WC_PHASE_REDUCE  1   /* backward: K4, K5a -> parameter gradients and red = {dWeff (G,g,g), gbar[Cin]}         */
#define WC_PHASE_DX      2   /* backward: K5b, pack, K6 -> dx                                                        */

typedef struct WcLayerCall {
  /* shapes and switches */
  int32_t N, HW, Cin, Cout, g, K, E;
  int32_t whiten, decomposition, renorm, training, triangular;
  int32_t phases, use_tensor_cores;
  int32_t relu;                 /* 1: y = max(0, .) -- the Activation('relu') that follows every WC layer in the reference's
                                   blocks (gan/layer_utils.py:170,178; generator.py:155) applied in K3's epilogue.  Backward then
                                   expects dy already multiplied by the mask (y > 0).                                          */
  int32_t reserved0;
  float eps, momentum;
  double count;                 /* rows of the global batch; count_dev[0] overrides it when non-NULL */
  const float* count_dev;
  /* parameters and state (any may be NULL) */
  const float *Wu, *bu, *Wc, *bc, *class_matrix, *D, *bf;
  float *moving_mean, *moving_cov;
  const int32_t* cls;
  int32_t* info;
  /* forward */
  const float* x;
  float* y;
  const float *mean_ext, *S_ext;   /* merged statistics (synchronised batch); NULL: use this shard's record */
  /* backward */
  const float* dy;
  float* dx;                       /* NULL: no input gradient wanted */
  float *dWu, *dbu, *dWc, *dbc, *dclass_matrix, *dD, *dbf;
  float* red;                      /* (G g g + Cin) floats, all-reduced by the host between the backward phases */
  /* caller-owned memory */
  void* saved;     size_t saved_bytes;
  void* workspace; size_t workspace_bytes;
  /* synchronised-batch statistics over peer memory (one node, NVLink P2P): every rank maps a symmetric buffer of
     wc_exchange_bytes() zero-initialised bytes; xchg_peers is a DEVICE array of xchg_world pointers to those buffers in
     rank order, xchg_state 16 zero-initialised device bytes owned by this rank.  With these set, training calls with
     both phases exchange the shard records (forward) and the adjoint `red` (backward) inside the call; count is this
     shard's row count.  NULL: single process, or the host runs the collectives between the phases. */
  void* const* xchg_peers;
  void* xchg_state;
  int32_t xchg_rank, xchg_world;
} WcLayerCall;

/* byte offsets into `saved` ((size_t)-1 when the piece does not exist for this configuration) */
typedef struct WcSavedLayout {
  size_t rec, mean, S, rec_count;   /* shard record: mean | S | count, contiguous from `rec` */
  size_t Weff, Wm, R, U, lam, Wk, T, b, ehat, Tp_fwd, Tp_bwd, A, cvec, total;
} WcSavedLayout;

/* bytes of the symmetric exchange buffer for one (C, g) on `world` ranks (0: unsupported) */
size_t wc_exchange_bytes(int C, int g, int world);
int wc_layer_saved_layout(const WcLayerCall* c, WcSavedLayout* out);
size_t wc_layer_saved_bytes(const WcLayerCall* c);
size_t wc_layer_workspace_bytes(const WcLayerCall* c, int backward);
int wc_layer_fwd(const WcLayerCall* c, wc_stream_t stream);
int wc_layer_bwd(const WcLayerCall* c, wc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WC_B200_H_ */
