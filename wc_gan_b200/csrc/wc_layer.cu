// Whole-layer entry points: one C call per direction sequences the stages of the WC layer on the caller's stream.
// Replaces, per training step, DecorelationNormalization.call + the coloring layer's call
// (gan/conditional_layers.py:646-712, :790-848, :971-1034; generator.py:41-88) and TF's autodiff through them.
// The stage functions are the ones exported individually in include/wc_b200.h; this file only carves the caller's
// `saved` / `workspace` blocks and calls them in order, so the host pays for one foreign call, not ten.
#include "wc_common.cuh"

namespace {

constexpr size_t NONE = (size_t)-1;
constexpr int CHAIN_TAIL = 4;      // floats after P | s zeroed together with them (u32 grid-barrier counter of the chain)

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off += wc_align_up(bytes, 256); return o; }
};

inline bool conditional(const WcLayerCall& c) { return c.K > 1 && c.cls != nullptr; }
inline bool has_coloring(const WcLayerCall& c) { return c.Wu || c.Wc || c.D; }
// only the unconditional kernel, one class: the per-class matrix IS Wu and the bias IS bu (nothing to build or to save;
// the caller keeps both alive and unmodified between the two directions: they are autograd-saved tensors)
inline bool plain_u(const WcLayerCall& c) { return c.K == 1 && c.Wu && !c.Wc && !c.D && !c.class_matrix && !c.bc && !c.bf; }
// statistics and adjoint exchanged over peer memory inside the call (training, whole call, more than one rank)
inline bool xchg_fwd(const WcLayerCall& c) {
  return c.xchg_peers && c.xchg_state && c.xchg_world > 1 && c.whiten && c.training && !c.renorm &&
         (c.phases & WC_PHASE_STATS) && (c.phases & WC_PHASE_APPLY);
}
inline bool xchg_bwd(const WcLayerCall& c) {
  return c.xchg_peers && c.xchg_state && c.xchg_world > 1 && c.whiten && c.training && !c.renorm && c.dx &&
         (c.phases & WC_PHASE_REDUCE) && (c.phases & WC_PHASE_DX);
}
inline bool use_tc(const WcLayerCall& c) {
  return c.use_tensor_cores && (wc_path_for((int64_t)c.N * c.HW, c.Cin, c.Cout, c.g) & 1);
}

int check(const WcLayerCall* c) {
  if (!c) return WC_E_ARG;
  if (c->N <= 0 || c->HW <= 0 || c->Cin <= 0 || c->Cout <= 0 || c->K <= 0 || c->E < 0) return WC_E_ARG;
  if (c->whiten && (c->g <= 0 || c->Cin % c->g)) return WC_E_ARG;
  if (c->class_matrix && (!c->D || c->E <= 0)) return WC_E_ARG;
  if (!has_coloring(*c) && c->Cin != c->Cout) return WC_E_ARG;
  return 0;
}

void layout(const WcLayerCall& c, WcSavedLayout& L) {
  Carver cv;
  const size_t f = sizeof(float);
  const size_t gg = c.whiten ? (size_t)c.Cin * c.g : 0;
  const size_t kcc = (size_t)c.K * c.Cin * c.Cout;
  const bool train_w = c.whiten && c.training;
  L.rec = L.mean = L.S = L.rec_count = L.Weff = L.Wm = L.R = L.U = L.lam = L.Wk = L.T = L.b = L.ehat = L.Tp_fwd = L.Tp_bwd =
      L.A = L.cvec = NONE;
  if (train_w) {
    L.rec = cv.take(((size_t)c.Cin + gg + 1) * f);
    L.mean = L.rec; L.S = L.rec + (size_t)c.Cin * f; L.rec_count = L.S + gg * f;
  }
  if (c.whiten) {
    L.Weff = cv.take(gg * f);
    L.Wm = (train_w && c.renorm) ? cv.take(gg * f) : L.Weff;
    if (train_w && c.renorm) L.R = cv.take(gg * f);
    if (c.decomposition == WC_DECOMP_ZCA) { L.U = cv.take(gg * f); L.lam = cv.take((size_t)c.Cin * f); }
  }
  L.Wk = cv.take(kcc * f);
  L.T = c.whiten ? cv.take(kcc * f) : L.Wk;
  if (c.bu || c.bc || c.bf) L.b = cv.take((size_t)c.K * c.Cout * f);
  if (c.class_matrix) L.ehat = cv.take((size_t)c.K * c.E * f);
  if (use_tc(c)) {
    L.Tp_fwd = cv.take(wc_pack_bytes(c.Cin, c.K, 0));
    L.Tp_bwd = cv.take(wc_pack_bytes(c.Cin, c.K, 1));
  }
  if (train_w) { L.A = cv.take(gg * f); L.cvec = cv.take((size_t)c.Cin * f); }
  L.total = cv.off;
}

struct FwdWs { size_t stats, stats_bytes, factor, factor_bytes, compose, total; };
FwdWs fwd_ws(const WcLayerCall& c) {
  Carver cv; FwdWs w{};
  if (c.whiten && c.training) { w.stats_bytes = wc_stats_workspace_bytes((int64_t)c.N * c.HW, c.Cin, c.g); w.stats = cv.take(w.stats_bytes); }
  if (c.whiten) { w.factor_bytes = wc_factor_workspace_bytes(c.Cin, c.g, c.decomposition, c.renorm && c.training); w.factor = cv.take(w.factor_bytes); }
  w.compose = NONE;
  if (c.whiten && c.Cin == c.Cout) {
    const size_t nb = wc_compose_fact_workspace_bytes(c.Cin, c.K, c.E, c.class_matrix != nullptr, c.Wc != nullptr);
    if (nb) w.compose = cv.take(nb);
  }
  w.total = cv.off ? cv.off : 256;
  return w;
}
struct BwdWs { size_t Ps, reduce, reduce_bytes, color, color_bytes, factor, factor_bytes, total; };
BwdWs bwd_ws(const WcLayerCall& c) {
  Carver cv; BwdWs w{};
  // P | s | grid-barrier counter of the chain launch, contiguous: one memset
  w.Ps = cv.take(((size_t)c.K * c.Cin * c.Cout + (size_t)c.K * c.Cout + CHAIN_TAIL) * sizeof(float));
  w.reduce_bytes = wc_reduce_workspace_bytes(c.N, c.HW, c.Cin, c.Cout, c.K); w.reduce = cv.take(w.reduce_bytes);
  w.color_bytes = wc_color_bwd_workspace_bytes(c.Cin, c.Cout, c.g > 0 ? c.g : c.Cin, c.K, c.E); w.color = cv.take(w.color_bytes);
  if (c.whiten && c.training) { w.factor_bytes = wc_factor_bwd_workspace_bytes(c.Cin, c.g, c.decomposition, c.renorm); w.factor = cv.take(w.factor_bytes); }
  w.total = cv.off;
  return w;
}

template <typename T> inline T* at(void* base, size_t off) { return off == NONE ? nullptr : reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off); }

}  // namespace

extern "C" size_t wc_sizeof_layer_call(void) { return sizeof(WcLayerCall); }
extern "C" size_t wc_sizeof_saved_layout(void) { return sizeof(WcSavedLayout); }

extern "C" int wc_layer_saved_layout(const WcLayerCall* c, WcSavedLayout* out) {
  int rc = check(c);
  if (rc || !out) return rc ? rc : WC_E_ARG;
  layout(*c, *out);
  return 0;
}
extern "C" size_t wc_layer_saved_bytes(const WcLayerCall* c) {
  if (check(c)) return 0;
  WcSavedLayout L; layout(*c, L);
  return L.total;
}
extern "C" size_t wc_layer_workspace_bytes(const WcLayerCall* c, int backward) {
  if (check(c)) return 0;
  return backward ? bwd_ws(*c).total : fwd_ws(*c).total;
}

extern "C" int wc_layer_fwd(const WcLayerCall* cp, wc_stream_t stream) {
  int rc = check(cp);
  if (rc) return rc;
  const WcLayerCall& c = *cp;
  if (!c.x || !c.saved) return WC_E_ARG;
  WcSavedLayout L; layout(c, L);
  if (c.saved_bytes < L.total) return WC_E_WORKSPACE;
  FwdWs W = fwd_ws(c);
  if ((c.whiten) && (!c.workspace || c.workspace_bytes < W.total)) return WC_E_WORKSPACE;
  if (!wc_aligned16(c.saved) || (c.workspace && !wc_aligned16(c.workspace))) return WC_E_ALIGN;
  const int64_t M = (int64_t)c.N * c.HW;
  const bool train_w = c.whiten && c.training;
  if (c.whiten && (!c.moving_mean || !c.moving_cov)) return WC_E_ARG;

  int prep_done = 0;
  unsigned* fwd_barrier = nullptr;       // zeroed by K1's accumulator memset: grid-barrier counter of the K2b chain launch
  WcRawStats raw{nullptr, nullptr, nullptr};   // peer-memory exchange: K1's finalize step runs inside the exchange kernel
  if ((c.phases & WC_PHASE_STATS) && train_w) {
    // single process, no renorm: the finalize kernel of the statistics also scales / shrinks the covariance and
    // updates the moving averages (K2's prep step), one launch less on the serial path
    const bool can_fuse = (c.phases & WC_PHASE_APPLY) && c.y && !c.renorm && !c.mean_ext && !c.S_ext && !c.count_dev && !xchg_fwd(c);
    WcPrepFuse f{c.eps, c.decomposition == WC_DECOMP_ZCA ? c.eps : 0.f, c.momentum, c.moving_mean, c.moving_cov,
                 at<float>(c.workspace, W.factor)};
    rc = wc_stats_fwd_impl(c.x, M, c.Cin, c.g, at<float>(c.workspace, W.stats), W.stats_bytes, at<float>(c.saved, L.mean),
                           at<float>(c.saved, L.S), can_fuse ? &f : nullptr, &prep_done, &fwd_barrier, xchg_fwd(c) ? &raw : nullptr,
                           (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (!(c.phases & WC_PHASE_APPLY)) return 0;
  if (!c.y) return WC_E_ARG;
  const float* count_dev = c.count_dev;
  if (xchg_fwd(c)) {
    // shard record -> every rank; Chan merge + prep from the complete set; saved mean | S | count become global
    rc = wc_stats_exchange(at<float>(c.saved, L.rec), (float)c.count, c.Cin, c.g, c.xchg_peers, c.xchg_state, c.xchg_rank,
                           c.xchg_world, c.eps, c.decomposition == WC_DECOMP_ZCA ? c.eps : 0.f, c.momentum, c.moving_mean,
                           c.moving_cov, at<float>(c.workspace, W.factor), c.info, raw.partS, raw.psum, raw.pivot, (double)M,
                           (cudaStream_t)stream);
    if (rc) return rc;
    prep_done = 1;
    count_dev = at<float>(c.saved, L.rec_count);
  }
  const float* mean = nullptr;
  if (c.whiten) {
    const float* S = nullptr;
    if (train_w) {
      mean = c.mean_ext ? c.mean_ext : at<float>(c.saved, L.mean);
      S = c.S_ext ? c.S_ext : at<float>(c.saved, L.S);
      if (c.mean_ext)   // keep the statistics the layer actually used where backward looks for them
        WC_CUDA(cudaMemcpyAsync(at<float>(c.saved, L.mean), c.mean_ext, (size_t)c.Cin * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
      mean = at<float>(c.saved, L.mean);
    } else {
      mean = c.moving_mean;
    }
    rc = wc_factor_fwd_impl(train_w ? mean : nullptr, S, count_dev, c.count, c.Cin, c.g, c.eps, c.momentum, c.decomposition,
                            c.renorm, c.training, c.moving_mean, c.moving_cov, at<float>(c.saved, L.Weff),
                            at<float>(c.saved, train_w ? L.Wm : L.Weff), at<float>(c.saved, L.R), at<float>(c.saved, L.U),
                            at<float>(c.saved, L.lam), c.info, at<float>(c.workspace, W.factor), W.factor_bytes, prep_done,
                            (cudaStream_t)stream);
    if (rc) return rc;
  }
  // identity coloring (after_norm == 'n', generator.py:79-80): W_k = I is written by the compose kernel from Wu == NULL
  void* Tp = at<void>(c.saved, L.Tp_fwd);
  rc = wc_compose_fwd_impl(at<float>(c.saved, L.Weff), c.Cin, c.Cout, c.whiten ? c.g : c.Cin, c.K, c.Wu, c.bu, c.Wc, c.bc,
                           c.triangular, c.class_matrix, c.D, c.E, c.bf, at<float>(c.saved, L.ehat), at<float>(c.saved, L.Wk),
                           at<float>(c.saved, L.T), at<float>(c.saved, L.b), Tp, Tp ? at<void>(c.saved, L.Tp_bwd) : nullptr,
                           plain_u(c) ? c.Wu : nullptr, fwd_barrier, c.workspace ? at<float>(c.workspace, W.compose) : nullptr,
                           (cudaStream_t)stream);
  if (rc) return rc;
  const int band = (c.whiten && !has_coloring(c) && c.g < c.Cin) ? c.g : 0;
  return wc_apply_fwd_impl(c.x, c.N, c.HW, c.Cin, c.Cout, mean, at<float>(c.saved, L.T), Tp, plain_u(c) ? c.bu : at<float>(c.saved, L.b),
                           conditional(c) ? c.cls : nullptr, c.K, band, c.y, c.info, c.relu, (cudaStream_t)stream,
                           c.whiten ? 1 : 0);      // x, mean and the labels were read by this call's own kernels before
}

extern "C" int wc_layer_bwd(const WcLayerCall* cp, wc_stream_t stream) {
  int rc = check(cp);
  if (rc) return rc;
  const WcLayerCall& c = *cp;
  if (!c.x || !c.dy || !c.saved || !c.workspace) return WC_E_ARG;
  WcSavedLayout L; layout(c, L);
  if (c.saved_bytes < L.total) return WC_E_WORKSPACE;
  BwdWs W = bwd_ws(c);
  if (c.workspace_bytes < W.total) return WC_E_WORKSPACE;
  if (!wc_aligned16(c.saved) || !wc_aligned16(c.workspace)) return WC_E_ALIGN;
  const bool train_w = c.whiten && c.training;
  const bool stats_grad = train_w && c.dx != nullptr;
  if (stats_grad && !c.red) return WC_E_ARG;
  const int g = c.whiten ? c.g : c.Cin;
  const size_t gg = (size_t)c.Cin * g;
  const float* mean = c.whiten ? (train_w ? at<float>(c.saved, L.mean) : c.moving_mean) : nullptr;
  const int32_t* cls = conditional(c) ? c.cls : nullptr;
  float* P = at<float>(c.workspace, W.Ps);
  float* s = P + (size_t)c.K * c.Cin * c.Cout;

  // Cholesky whitening without renorm on tile-aligned shapes: K5a + K5b run as ONE persistent launch (wc_bwd_chain);
  // with the adjoint exchanged between ranks, as two (stage 0 | exchange | stages 1-3)
  unsigned* barrier = reinterpret_cast<unsigned*>(s + (size_t)c.K * c.Cout);
  const bool chain_ok = stats_grad && c.decomposition == WC_DECOMP_CHOLESKY && !c.renorm && c.Cin == c.Cout &&
                        (g % 32) == 0 && !(c.triangular && c.dWc) && !(wc_debug_flags() & 262144);
  const float* Wk = plain_u(c) ? c.Wu : at<float>(c.saved, L.Wk);
  float *A = stats_grad ? at<float>(c.saved, L.A) : nullptr, *cvec = stats_grad ? at<float>(c.saved, L.cvec) : nullptr;
  const float* count_dev = c.count_dev;
  const bool xb = stats_grad && xchg_bwd(c);
  if (xb) count_dev = at<float>(c.saved, L.rec_count);       // global row count, left there by the forward exchange
  // the factorised forward (many classes, fconv / ufconv) does not form W_k: whoever reads it below rebuilds it first
  const int wk_missing = stats_grad && !plain_u(c) && c.Cin == c.Cout && !c.bc &&
                         wc_bwd_factorised_ok(c.Cin, c.K, c.E, c.class_matrix != nullptr, c.Wc != nullptr);
  // the sums over classes (or basis matrices) that form dWeff are split over several CTAs which add into c.red: the kernel
  // that zeroes P and s clears it too
  // (experiment bit 2^26 = 67108864: one CTA per tile walks the whole sum, as before)
  // Only with use_tensor_cores: the FFMA path (use_tensor_cores = 0) stays deterministic, and sums added in arrival order
  // are not.
  const bool split_sum = chain_ok && c.use_tensor_cores && (c.phases & WC_PHASE_REDUCE) && c.K > 1 && wc_aligned16(c.red) &&
                         !(wc_debug_flags() & 67108864);
  // The products of stages 1 and 2 (Psi, Y) are split along K over several CTAs that add their partial tiles into the two
  // scratch matrices at the head of the factor workspace: cleared by the same kernel, when the whole chain runs in this call.
  // (experiment bit 2^28 = 268435456: one CTA per tile)
  const bool whole_call = (c.phases & WC_PHASE_REDUCE) && (c.phases & WC_PHASE_DX) && c.dx;
  const bool split_k = chain_ok && c.use_tensor_cores && whole_call && W.factor_bytes >= 2 * wc_align_up(gg * 4, 256) &&
                       wc_aligned16(at<float>(c.workspace, W.factor)) && !(wc_debug_flags() & 268435456);
  auto chain = [&](int lo, int hi, bool exchange) {
    return wc_bwd_chain(P, s, at<float>(c.saved, L.Weff), Wk, c.Wu, c.Cin, c.Cout, g, c.K, c.dWu, c.dbu, c.dWc, c.dbc, c.triangular,
                        c.class_matrix, at<float>(c.saved, L.ehat), c.D, c.E, c.dD, c.dclass_matrix, c.dbf, c.red, A, cvec,
                        count_dev, c.count, c.eps, at<void>(c.saved, L.Tp_bwd), c.K, at<float>(c.workspace, W.color),
                        at<float>(c.workspace, W.factor), barrier, lo, hi, exchange ? c.xchg_peers : nullptr,
                        exchange ? c.xchg_state : nullptr, c.xchg_rank, c.xchg_world, c.info, wk_missing, split_sum && lo == 0,
                        split_k && lo == 0 && hi == 4, (cudaStream_t)stream);
  };
  bool chained = false, all_done = false;
  if (c.phases & WC_PHASE_REDUCE) {
    rc = wc_reduce_bwd_impl(c.x, c.dy, c.N, c.HW, c.Cin, c.Cout, mean, cls, c.K, P, s, at<float>(c.workspace, W.reduce),
                            W.reduce_bytes, CHAIN_TAIL, split_sum ? c.red : nullptr, split_sum ? gg : 0,
                            split_k ? at<float>(c.workspace, W.factor) : nullptr, split_k ? 2 * wc_align_up(gg * 4, 256) / 4 : 0,
                            (cudaStream_t)stream);
    if (rc) return rc;
    if (chain_ok) {
      // one launch for the whole chain; with the adjoint exchanged over peer memory, the all-reduce is a stage of it
      const bool whole = (c.phases & WC_PHASE_DX) && c.dx;
      rc = chain(0, whole ? 4 : 1, whole && xb);
      if (rc == WC_E_UNSUPPORTED && whole && xb) { rc = chain(0, 1, false); if (rc == 0) { chained = true; } }
      else if (rc == 0) { chained = true; all_done = whole; }
      if (rc && rc != WC_E_UNSUPPORTED) return rc;
    }
    if (!chained) {
      if (wk_missing) { rc = wc_rebuild_wk(c.Cin, c.Cout, c.K, c.Wu, c.class_matrix, c.D, c.E, at<float>(c.saved, L.Wk), (cudaStream_t)stream); if (rc) return rc; }
      rc = wc_color_bwd(P, s, at<float>(c.saved, L.Weff), Wk, c.Cin, c.Cout, g, c.K, c.dWu, c.dbu, c.dWc,
                        c.dbc, c.triangular, c.class_matrix, at<float>(c.saved, L.ehat), c.D, c.E, c.dD, c.dclass_matrix, c.dbf,
                        stats_grad ? c.red : nullptr, stats_grad ? c.red + gg : nullptr, at<float>(c.workspace, W.color),
                        W.color_bytes, stream);
      if (rc) return rc;
    }
  }
  if (!(c.phases & WC_PHASE_DX) || !c.dx) return 0;
  if (xb && !all_done) {
    rc = wc_adjoint_exchange(c.red, c.Cin, g, c.xchg_peers, c.xchg_state, c.xchg_rank, c.xchg_world, c.info, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (stats_grad && !all_done) {
    rc = WC_E_UNSUPPORTED;
    // (the barrier counter is still zero here: a one-stage launch never touches it)
    if (chain_ok && (chained || !(c.phases & WC_PHASE_REDUCE))) rc = chain(1, 4, false);
    if (rc == WC_E_UNSUPPORTED)
      rc = wc_factor_bwd_impl(c.red, c.red + gg, at<float>(c.saved, L.Weff), at<float>(c.saved, L.Wm), at<float>(c.saved, L.R),
                              at<float>(c.saved, L.U), at<float>(c.saved, L.lam), count_dev, c.count, c.Cin, g, c.eps,
                              c.decomposition, c.renorm, A, cvec, at<float>(c.workspace, W.factor), W.factor_bytes,
                              at<void>(c.saved, L.Tp_bwd), c.K, (cudaStream_t)stream);
    if (rc) return rc;
  }
  void* Tp = at<void>(c.saved, L.Tp_bwd);
  // (x, dy, the mean and the labels went through K4 of this call; T's operand image dates from the forward pass: only A's
  // image and c are written by the kernel right before K6)
  const int early = (c.phases & WC_PHASE_REDUCE) ? 3 : 0;
  return wc_apply_bwd_impl(c.x, c.dy, c.N, c.HW, c.Cin, c.Cout, mean, at<float>(c.saved, L.T), A, Tp, cvec, g, cls, c.K, c.dx,
                           (cudaStream_t)stream, early);
}
