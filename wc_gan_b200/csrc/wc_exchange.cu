// Synchronised-batch whitening without a library collective: the two small exchanges of a WC layer run inside the
// library's kernels over peer memory (NVLink / NVSwitch P2P stores into a symmetric buffer that every rank of the node
// has mapped).  Protocol: wc_xch.cuh.
//
//   forward : every rank holds the raw statistics of ITS rows.  stats_exchange_kernel finalizes them into the shard record
//             {mean[C], S[G g g], count} (K1's finalize step), pushes the record to every rank, and -- reading every rank's
//             record from its own buffer as the words arrive -- applies Chan's parallel update (mean / scatter of the
//             concatenated batch, gan/conditional_layers.py:658,:683 on the global batch) and K2's prep step (covariance
//             scaling, moving averages, shrinkage, :684-691) in the same pass.  Replaces finalize + all_gather +
//             wc_stats_merge + factor_prep.
//   backward: the adjoint {dWeff[G g g], gbar[C]} is pushed the same way and summed over ranks in rank order (all-reduce,
//             bit-identical on every rank).  Normally a stage of the backward chain kernel (wc_small.cu);
//             adjoint_exchange_kernel is the stand-alone version for the configurations the chain does not cover.
#include "wc_common.cuh"
#include "wc_xch.cuh"
#include <stdlib.h>

namespace {

struct XchParams {
  char* const* peers;        // device array of R pointers: every rank's exchange buffer (index = rank)
  uint32_t* state;           // local {epoch, ticket} of this direction
  int rank, R;
  size_t data_off, slot_words;   // byte offset of this direction's slots, 8-byte words per (slot, rank) record
  int n;                     // floats actually exchanged
  float tail;                // forward: the record's last element (this shard's row count) is supplied by value
  int32_t* info;             // nullable: info[WC_INFO_XCHG_TIMEOUT] is set when a peer's words never arrived
  unsigned long long timeout_ns;   // bound of every poll (0 = wait forever)
  unsigned long long* dbg;   // debug marks (nullptr in normal use): slots 28..31 of the mark buffer
};

__device__ __forceinline__ uint32_t xch_epoch(const XchParams& p) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(p.state);
  __syncthreads();
  return s_epoch;
}
__device__ __forceinline__ unsigned long long* xch_slot(const XchParams& p, int buffer_of, uint32_t slot, int sender) {
  return reinterpret_cast<unsigned long long*>(p.peers[buffer_of] + p.data_off) + ((size_t)slot * p.R + sender) * p.slot_words;
}
// the last CTA to finish advances the epoch (every CTA read it at its start)
__device__ __forceinline__ void xch_finish(const XchParams& p, uint32_t e, bool ok) {
  if (!ok && p.info) atomicMax(&p.info[WC_INFO_XCHG_TIMEOUT], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(p.state + 1, 1u);
    if (done == gridDim.x - 1) {
      p.state[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(p.state) = e + 1;
    }
  }
}

// K1's raw accumulators (tensor-core path) or NULL partS: the record is already in m.rec
struct FinalizeSrc { const float* partS; const float* psum; const float* pivot; double M; };

struct FwdMerge {
  int C, g;
  float eps, extra_eps, momentum;
  float* moving_mean; float* moving_cov;
  float* rec;        // in (no FinalizeSrc): this shard's record; out: mean | S | count of the concatenated batch
  float* Ashr;       // out: shrunk covariance for the factorisation
};

__global__ void __launch_bounds__(XCH_THREADS) stats_exchange_kernel(XchParams p, FwdMerge m, FinalizeSrc fin) {
  wc_pdl_wait();
  if (threadIdx.x == 0) wc_mark(p.dbg, 28);
  __shared__ float s_dd[512];                      // sum(x - pivot) / M per channel (C <= 512 on this path)
  const int C = m.C, g = m.g;
  const long long total = (long long)C * g;
  const uint32_t e = xch_epoch(p), slot = e & 1u, tag = e + 1u;
  if (fin.partS) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_dd[i] = (float)((double)fin.psum[i] / fin.M);
    __syncthreads();
  }
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  // ---- finalize my elements of the record and push each to every rank (own buffer included)
  for (int i = gtid; i < p.n; i += gstride) {
    float v;
    if (!fin.partS) v = (i == p.n - 1) ? p.tail : m.rec[i];
    else if (i < C) v = fin.pivot[i] + s_dd[i];
    else if (i < C + total) {
      const int idx = i - C, grp = idx / (g * g), rem = idx - grp * g * g;
      const int ri = grp * g + rem / g, ci = grp * g + rem % g;
      // S = S_p - M d d^T: the product is formed in double (M d_i d_j can be of the size of S_p itself)
      v = (float)((double)fin.partS[(long long)max(ri, ci) * C + min(ri, ci)] - fin.M * (double)s_dd[ri] * (double)s_dd[ci]);
    } else v = p.tail;
    for (int q = 0; q < p.R; ++q) {
      const int r = (p.rank + 1 + q) % p.R;               // start with the neighbour: spreads the NVLink traffic
      ll_store(xch_slot(p, r, slot, p.rank) + i, v, tag);
    }
  }
  if (threadIdx.x == 0) wc_mark(p.dbg, 29);
  // ---- merge + prep from the R records in my own buffer, word by word as they arrive
  bool ok = true;
  const unsigned long long* mine = xch_slot(p, p.rank, slot, 0);
  double nr[XCH_MAXR], n = 0;
  for (int r0 = 0; r0 < p.R; r0 += 4) {
    const unsigned long long* src[4]; float v[4];
    const int cnt4 = min(4, p.R - r0);
#pragma unroll
    for (int k = 0; k < 4; ++k) src[k] = mine + (size_t)min(r0 + k, p.R - 1) * p.slot_words + C + total;
    ll_load_many<4>(src, v, cnt4, tag, p.timeout_ns, &ok);
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k < cnt4) { nr[r0 + k] = (double)v[k]; n += nr[r0 + k]; }
  }
  if (threadIdx.x == 0) wc_mark(p.dbg, 30);
  const float cnt = (float)n;
  for (long long idx = gtid; idx < total; idx += gstride) {
    const int grp = (int)(idx / ((long long)g * g));
    const int rem = (int)(idx % ((long long)g * g));
    const int rr = rem / g, cc = rem % g;
    const int i = grp * g + rr, j = grp * g + cc;
    float mir[XCH_MAXR], mjr[XCH_MAXR], sr[XCH_MAXR];
    double mi = 0, mj = 0;
    for (int r0 = 0; r0 < p.R; r0 += 4) {            // 4 ranks x {mean_i, mean_j, S_ij}: 12 words in flight
      const unsigned long long* src[12]; float v[12];
      const int cnt4 = min(4, p.R - r0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned long long* rec = mine + (size_t)min(r0 + k, p.R - 1) * p.slot_words;
        src[3 * k] = rec + i; src[3 * k + 1] = rec + j; src[3 * k + 2] = rec + C + idx;
      }
      ll_load_many<12>(src, v, 3 * cnt4, tag, p.timeout_ns, &ok);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < cnt4) {
          mir[r0 + k] = v[3 * k]; mjr[r0 + k] = v[3 * k + 1]; sr[r0 + k] = v[3 * k + 2];
          mi += nr[r0 + k] * v[3 * k]; mj += nr[r0 + k] * v[3 * k + 1];
        }
    }
    mi /= n; mj /= n;
    double s = 0;
    for (int r = 0; r < p.R; ++r) s += (double)sr[r] + nr[r] * ((double)mir[r] - mi) * ((double)mjr[r] - mj);
    const float sv = (float)s;
    m.rec[C + idx] = sv;
    const float cov = sv / (cnt - 1.f);
    m.moving_cov[idx] = m.moving_cov[idx] * m.momentum + cov * (1.f - m.momentum);          // conditional_layers.py:687-689
    if (rr == cc) {
      const float mu = (float)mi;
      m.rec[i] = mu;
      m.moving_mean[i] = m.moving_mean[i] * m.momentum + mu * (1.f - m.momentum);            // :684-686
    }
    m.Ashr[idx] = (1.f - m.eps) * cov + (rr == cc ? m.eps + m.extra_eps : 0.f);              // :691 (+ :671 for ZCA)
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) m.rec[C + total] = cnt;
  if (threadIdx.x == 0) wc_mark(p.dbg, 31);
  xch_finish(p, e, ok);
}

__global__ void __launch_bounds__(XCH_THREADS) adjoint_exchange_kernel(XchParams p, float* __restrict__ red) {
  wc_pdl_wait();
  const uint32_t e = xch_epoch(p), slot = e & 1u, tag = e + 1u;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  bool ok = true;
  const unsigned long long* mine = xch_slot(p, p.rank, slot, 0);
  for (int i = gtid; i < p.n; i += gstride) {
    const float v = red[i];
    for (int q = 0; q < p.R; ++q) ll_store(xch_slot(p, (p.rank + 1 + q) % p.R, slot, p.rank) + i, v, tag);
  }
  for (int i = gtid; i < p.n; i += gstride) {      // the same elements this thread pushed: nobody else touches red[i]
    float s = 0.f;
    for (int r0 = 0; r0 < p.R; r0 += 4) {            // rank order
      const unsigned long long* src[4]; float v[4];
      const int cnt4 = min(4, p.R - r0);
#pragma unroll
      for (int k = 0; k < 4; ++k) src[k] = mine + (size_t)min(r0 + k, p.R - 1) * p.slot_words + i;
      ll_load_many<4>(src, v, cnt4, tag, p.timeout_ns, &ok);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (k < cnt4) s += v[k];
    }
    red[i] = s;
  }
  xch_finish(p, e, ok);
}

}  // namespace

extern "C" size_t wc_exchange_bytes(int C, int g, int world) {
  if (C <= 0 || g <= 0 || C % g || world < 1 || world > XCH_MAXR) return 0;
  return xch_layout(C, g, world).total;
}

unsigned long long wc_xch_timeout_ns() {
  static const unsigned long long v = [] {
    const char* e = getenv("WC_B200_XCHG_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 10000;
    return ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
  }();
  return v;
}

// rec: mean | S | count of this shard (the count is passed by value; with partS the kernel computes mean and S itself from
// K1's raw accumulators); on return the merged record.  state: 4 zero-initialised uint32 owned by the layer ({epoch,
// ticket} for forward, then backward).
int wc_stats_exchange(float* rec, float count_local, int C, int g, void* const* peers, void* state, int rank, int world, float eps,
                      float extra_eps, float momentum, float* moving_mean, float* moving_cov, float* Ashr, int32_t* info,
                      const float* partS, const float* psum, const float* pivot, double M, cudaStream_t st) {
  if (!rec || !peers || !state || !moving_mean || !moving_cov || !Ashr || rank < 0 || rank >= world || world > XCH_MAXR) return WC_E_ARG;
  if (partS && C > 512) return WC_E_UNSUPPORTED;
  const XchLayout L = xch_layout(C, g, world);
  XchParams p{};
  p.peers = reinterpret_cast<char* const*>(peers); p.state = reinterpret_cast<uint32_t*>(state);
  p.rank = rank; p.R = world; p.data_off = L.fwd_off; p.slot_words = L.fwd_words;
  p.n = C + C * g + 1; p.tail = count_local;
  p.info = info; p.timeout_ns = wc_xch_timeout_ns(); p.dbg = wc_debug_buf();
  FwdMerge m{C, g, eps, extra_eps, momentum, moving_mean, moving_cov, rec, Ashr};
  FinalizeSrc fin{partS, psum, pivot, M};
  WC_CUDA(wc_launch(stats_exchange_kernel, XCH_CTAS, XCH_THREADS, 0, st, p, m, fin));
  WC_CHECK_LAUNCH();
  return 0;
}

int wc_adjoint_exchange(float* red, int C, int g, void* const* peers, void* state, int rank, int world, int32_t* info,
                        cudaStream_t st) {
  if (!red || !peers || !state || rank < 0 || rank >= world || world > XCH_MAXR) return WC_E_ARG;
  const XchLayout L = xch_layout(C, g, world);
  XchParams p{};
  p.peers = reinterpret_cast<char* const*>(peers); p.state = reinterpret_cast<uint32_t*>(state) + 2;
  p.rank = rank; p.R = world; p.data_off = L.bwd_off; p.slot_words = L.bwd_words;
  p.n = C * g + C;
  p.info = info; p.timeout_ns = wc_xch_timeout_ns(); p.dbg = nullptr;
  WC_CUDA(wc_launch(adjoint_exchange_kernel, XCH_CTAS, XCH_THREADS, 0, st, p, red));
  WC_CHECK_LAUNCH();
  return 0;
}
