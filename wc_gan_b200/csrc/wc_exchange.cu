// Synchronised-batch whitening without a library collective: the two small exchanges of a WC layer run as ONE kernel
// each, over peer memory (NVLink / NVSwitch P2P stores into a symmetric buffer that every rank of the node has mapped).
//
//   forward : every rank holds one shard record {mean[C], S[G g g], count} of ITS rows.  The kernel pushes the record
//             into slot [epoch & 1][rank] of EVERY rank's buffer, signals, waits for all peers, then -- from its own,
//             now complete, buffer -- applies Chan's parallel update (mean / scatter of the concatenated batch,
//             gan/conditional_layers.py:658,:683 on the global batch) and K2's prep step (covariance scaling, moving
//             averages, shrinkage, :684-691) in the same pass.  Replaces all_gather + wc_stats_merge + factor_prep.
//   backward: the adjoint {dWeff[G g g], gbar[C]} is pushed the same way and summed over ranks (all-reduce).
//
// Protocol (per direction): counters only grow.  Exchange number e uses data slot e & 1; each of the NCTA CTAs of rank
// q adds 1 (release, system scope) to flag[e & 1][q] on every rank after its slice has been stored there, and every
// CTA waits (acquire, system scope) until all R flags of its own buffer reach NCTA * (e / 2 + 1).  A rank cannot reach
// exchange e + 2 (which reuses slot e & 1) before every peer has finished reading exchange e, because exchange e + 1
// needs every peer's push of e + 1, which a peer issues only after its reads of e (stream order).  The own flag
// reaching its target also means every local CTA has finished READING the local record, so the merged statistics may
// overwrite it in place.  All CTAs of the grid must be co-resident (they spin): the grid is far below the SM count.
#include "wc_common.cuh"
#include "wc_xch.cuh"
#include <stdlib.h>

namespace {

struct XchParams {
  char* const* peers;        // device array of R pointers: every rank's exchange buffer (index = rank)
  uint32_t* state;           // local {epoch, ticket} of this direction
  int rank, R, dir;          // dir 0 forward, 1 backward
  size_t data_off, slot_floats;   // byte offset of this direction's slots, floats per (slot, rank) record
  int n;                     // floats actually exchanged
  int has_tail; float tail;  // forward: the record's last element (this shard's row count) is supplied by value
  int32_t* info;             // nullable: info[WC_INFO_XCHG_TIMEOUT] is set when a peer never arrived
  unsigned long long timeout_ns;   // bound of the spin (0 = wait forever)
  unsigned long long* dbg;   // debug marks (nullptr in normal use): slots 28..31 of the mark buffer
  int dbg_flags;
};

// push `src[0..n)` into slot [e & 1][rank] of every peer, signal, wait for all peers.  Returns the epoch.
__device__ __forceinline__ uint32_t xch_push_wait(const XchParams& p, const float* __restrict__ src) {
  __shared__ uint32_t s_epoch;
  const int t = threadIdx.x;
  if (t == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(p.state);
  __syncthreads();
  const uint32_t e = s_epoch, slot = e & 1u;
  // slice of this CTA (float4 granularity; n is a multiple of 4 or the tail is handled scalar by the last CTA)
  const int n4 = p.n >> 2;
  const int per = (n4 + gridDim.x - 1) / gridDim.x;
  const int b4 = blockIdx.x * per, e4 = min(n4, b4 + per);
  const size_t rec_off = p.data_off + ((size_t)slot * p.R + p.rank) * p.slot_floats * 4;
  for (int q = 0; q < p.R; ++q) {
    const int r = (p.rank + 1 + q) % p.R;               // start with the neighbour: spreads the NVLink traffic
    float4* dst = reinterpret_cast<float4*>(p.peers[r] + rec_off);
    if ((p.dbg_flags & 4194304) && r != p.rank) continue;     // experiment: no remote stores (the peers then time out / read stale data)
    // plain (coherent) loads: the same kernel later overwrites `src` in place with the merged result
    for (int i = b4 + t; i < e4; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(src)[i];
    if (blockIdx.x == gridDim.x - 1) {
      float* dsts = reinterpret_cast<float*>(p.peers[r] + rec_off);
      for (int i = (n4 << 2) + t; i < p.n; i += blockDim.x) dsts[i] = src[i];
      if (p.has_tail) {
        __syncthreads();                       // after whichever store covered the last element
        if (t == 0) dsts[p.n - 1] = p.tail;
      }
    }
  }
  if (t == 0) wc_mark(p.dbg, 30);
  // release pattern: the CTA barrier orders every thread's stores before the signalling threads' release-RMW at
  // system scope (release is cumulative over what the barrier made visible); no per-thread system fence is needed
  __syncthreads();
  const size_t flag_off = ((size_t)p.dir * 2 + slot) * XCH_MAXR * 4;
  if (t < p.R) red_release_sys_add(reinterpret_cast<uint32_t*>(p.peers[t] + flag_off) + p.rank, 1u);
  const uint32_t target = gridDim.x * (e / 2u + 1u);
  if (t < p.R) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p.peers[p.rank] + flag_off) + t;
    // bounded spin: a peer that died (or whose exchange is queued behind a kernel that waits for us) must not hang
    // the GPU.  The deadline is only consulted every 1024 polls.
    const unsigned long long t0 = p.timeout_ns ? xch_now() : 0ull;
    uint32_t polls = 0;
    while ((int32_t)(ld_acquire_sys(f) - target) < 0) {
      if (p.timeout_ns && ((++polls & 1023u) == 0) && xch_now() - t0 > p.timeout_ns) {
        if (p.info) atomicMax(&p.info[WC_INFO_XCHG_TIMEOUT], 1 + t);     // 1 + rank of the peer that is missing
        break;
      }
    }
  }
  __syncthreads();
  if (t == 0) wc_mark(p.dbg, 31);
  return e;
}

// the last CTA to finish advances the epoch (every CTA read it at its start)
__device__ __forceinline__ void xch_finish(const XchParams& p, uint32_t e) {
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

// K1's raw accumulators (tensor-core path): when given, the exchange kernel computes the shard record itself -- the finalize
// step of wc_stats_fwd, d = sum(x-p)/M, mean = p + d, S = S_p - M d d^T -- right before pushing it (one launch less on the
// serial path of every synchronised layer)
struct FinalizeSrc { const float* partS; const float* psum; const float* pivot; double M; };

struct FwdMerge {
  int C, g;
  float eps, extra_eps, momentum;
  float* moving_mean; float* moving_cov;
  float* rec;        // in: this shard's record; out: mean | S | count of the concatenated batch
  float* Ashr;       // out: shrunk covariance for the factorisation
};

__global__ void __launch_bounds__(XCH_THREADS) stats_exchange_kernel(XchParams p, FwdMerge m, FinalizeSrc fin) {
  wc_pdl_wait();
  if (threadIdx.x == 0) wc_mark(p.dbg, 28);
  if (fin.partS) {
    // every CTA finalizes exactly the slice of the record it is about to push (same float4 partition as xch_push_wait)
    const int C = m.C, g = m.g;
    const int n4 = p.n >> 2;
    const int per = (n4 + gridDim.x - 1) / gridDim.x;
    const int b = 4 * blockIdx.x * per, e2 = (blockIdx.x == gridDim.x - 1) ? p.n : min(p.n, b + 4 * per);
    for (int i = b + threadIdx.x; i < e2; i += blockDim.x) {
      float v;
      if (i < C) {
        v = (float)((double)fin.pivot[i] + (double)fin.psum[i] / fin.M);
      } else if (i < C + C * g) {
        const int idx = i - C, grp = idx / (g * g), rem = idx % (g * g);
        const int ri = grp * g + rem / g, ci = grp * g + rem % g;
        const double ddi = (double)fin.psum[ri] / fin.M, ddj = (double)fin.psum[ci] / fin.M;
        v = (float)((double)fin.partS[(long long)max(ri, ci) * C + min(ri, ci)] - fin.M * ddi * ddj);
      } else {
        v = p.tail;
      }
      m.rec[i] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) wc_mark(p.dbg, 29);
  const uint32_t e = xch_push_wait(p, m.rec);
  const int C = m.C, g = m.g;
  const long long total = (long long)C * g;
  const float* base = reinterpret_cast<const float*>(p.peers[p.rank] + p.data_off + (size_t)(e & 1u) * p.R * p.slot_floats * 4);
  double nr[XCH_MAXR], n = 0;
  for (int r = 0; r < p.R; ++r) { nr[r] = (double)base[(size_t)r * p.slot_floats + C + total]; n += nr[r]; }
  const float cnt = (float)n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int grp = (int)(idx / ((long long)g * g));
    const int rem = (int)(idx % ((long long)g * g));
    const int rr = rem / g, cc = rem % g;
    const int i = grp * g + rr, j = grp * g + cc;
    double mi = 0, mj = 0;
    for (int r = 0; r < p.R; ++r) {
      const float* rec = base + (size_t)r * p.slot_floats;
      mi += nr[r] * rec[i]; mj += nr[r] * rec[j];
    }
    mi /= n; mj /= n;
    double s = 0;
    for (int r = 0; r < p.R; ++r) {
      const float* rec = base + (size_t)r * p.slot_floats;
      const double di = rec[i] - mi, dj = rec[j] - mj;
      s += rec[C + idx] + nr[r] * di * dj;
    }
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
  xch_finish(p, e);
}

__global__ void __launch_bounds__(XCH_THREADS) adjoint_exchange_kernel(XchParams p, float* __restrict__ red) {
  wc_pdl_wait();
  const uint32_t e = xch_push_wait(p, red);
  const float* base = reinterpret_cast<const float*>(p.peers[p.rank] + p.data_off + (size_t)(e & 1u) * p.R * p.slot_floats * 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.n; idx += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < p.R; ++r) s += base[(size_t)r * p.slot_floats + idx];     // rank order: identical on every rank
    red[idx] = s;
  }
  xch_finish(p, e);
}

}  // namespace

extern "C" size_t wc_exchange_bytes(int C, int g, int world) {
  if (C <= 0 || g <= 0 || C % g || world < 1 || world > XCH_MAXR) return 0;
  return xch_layout(C, g, world).total;
}

// rec: mean | S | count of this shard (the count is passed by value); on return the merged record.  state: 4 zero-initialised
// uint32 owned by the layer ({epoch, ticket} for forward, then backward).
unsigned long long wc_xch_timeout_ns() {
  static const unsigned long long v = [] {
    const char* e = getenv("WC_B200_XCHG_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 10000;
    return ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
  }();
  return v;
}

int wc_stats_exchange(float* rec, float count_local, int C, int g, void* const* peers, void* state, int rank, int world, float eps,
                      float extra_eps, float momentum, float* moving_mean, float* moving_cov, float* Ashr, int32_t* info,
                      const float* partS, const float* psum, const float* pivot, double M, cudaStream_t st) {
  if (!rec || !peers || !state || !moving_mean || !moving_cov || !Ashr || rank < 0 || rank >= world || world > XCH_MAXR) return WC_E_ARG;
  const XchLayout L = xch_layout(C, g, world);
  XchParams p{};
  p.peers = reinterpret_cast<char* const*>(peers); p.state = reinterpret_cast<uint32_t*>(state);
  p.rank = rank; p.R = world; p.dir = 0; p.data_off = L.fwd_off; p.slot_floats = L.fwd_floats;
  p.n = C + C * g + 1; p.has_tail = 1; p.tail = count_local;
  p.info = info; p.timeout_ns = wc_xch_timeout_ns(); p.dbg = wc_debug_buf(); p.dbg_flags = wc_debug_flags();
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
  p.rank = rank; p.R = world; p.dir = 1; p.data_off = L.bwd_off; p.slot_floats = L.bwd_floats;
  p.n = C * g + C;
  p.info = info; p.timeout_ns = wc_xch_timeout_ns();
  WC_CUDA(wc_launch(adjoint_exchange_kernel, XCH_CTAS, XCH_THREADS, 0, st, p, red));
  WC_CHECK_LAUNCH();
  return 0;
}
