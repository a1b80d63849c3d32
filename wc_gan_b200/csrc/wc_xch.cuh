// Peer-memory exchange protocol shared by wc_exchange.cu (forward statistics; stand-alone adjoint kernel) and the chain
// kernel of wc_small.cu (adjoint all-reduce as a stage of the backward chain).
//
// "LL" protocol (data and flag travel in the same 8-byte word, as in NCCL's low-latency protocol): every float is pushed
// as one aligned 64-bit store {value, tag} with tag = epoch + 1, into slot [epoch & 1][sender rank] of every rank's buffer
// (P2P stores over NVLink / NVSwitch); a reader polls the word it needs until its tag matches.  A 64-bit aligned store
// is single-copy atomic, so a matching tag means the value beside it has arrived: no separate flag, no system-scope
// release fence after the data (which, measured, cost ~10 us per exchange: the fence waits for the acknowledgement of
// every outstanding remote write before the flag can leave), one-way NVLink latency instead of a round trip.
//
// Slot reuse: exchange e uses slot e & 1 and tag e + 1.  A rank cannot reach exchange e + 2 (same slot) before every peer
// has finished READING exchange e there: its exchange e + 1 needs every peer's words of e + 1, which a peer pushes only
// after its kernel of exchange e has completed (stream order).  Buffers start zeroed; tag 0 is never used.
#pragma once
#include "wc_common.cuh"

// 128 CTAs x 512 threads: one record element per thread at C = 256.  Every CTA pushes before it waits and waits only for
// remote words, but the CTAs of the peers' kernels must be able to START for those words to leave: the grid has to be fully
// resident (it is: 4 such CTAs fit on an SM, 2 next to a CTA of the preceding statistics kernel).
constexpr int XCH_CTAS = 128, XCH_THREADS = 512, XCH_MAXR = 16;

struct XchLayout { size_t fwd_words, bwd_words, fwd_off, bwd_off, total; };   // words of 8 bytes; offsets in bytes

static inline XchLayout xch_layout(int C, int g, int R) {
  XchLayout L;
  const size_t gg = (size_t)C * g;
  L.fwd_words = wc_align_up(C + gg + 1, 64);        // one padded shard record {mean[C], S[G g g], count}
  L.bwd_words = wc_align_up(gg + C, 64);            // the adjoint {dWeff[G g g], gbar[C]}
  L.fwd_off = 0;
  L.bwd_off = L.fwd_off + 2 * (size_t)R * L.fwd_words * 8;
  L.total = L.bwd_off + 2 * (size_t)R * L.bwd_words * 8;
  return L;
}

__device__ __forceinline__ unsigned long long xch_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ll_store(unsigned long long* dst, float v, uint32_t tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
}
// poll one word until its tag matches; *ok is cleared on time-out (the value returned is then meaningless)
__device__ __forceinline__ float ll_load(const unsigned long long* src, uint32_t tag, unsigned long long timeout_ns, bool* ok) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
  if ((uint32_t)(w >> 32) != tag) {
    const unsigned long long t0 = timeout_ns ? xch_now() : 0ull;
    uint32_t polls = 0;
    do {
      asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
      if (timeout_ns && ((++polls & 255u) == 0) && xch_now() - t0 > timeout_ns) { *ok = false; break; }
    } while ((uint32_t)(w >> 32) != tag);
  }
  return __uint_as_float((uint32_t)w);
}
// K words at once: all loads are issued before the first tag is looked at (one L2 round trip instead of K), the words that
// have not arrived yet are then polled one by one
template <int K>
__device__ __forceinline__ void ll_load_many(const unsigned long long* const (&src)[K], float (&out)[K], int count, uint32_t tag,
                                             unsigned long long timeout_ns, bool* ok) {
  unsigned long long w[K];
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (k < count) asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w[k]) : "l"(src[k]) : "memory");
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (k < count) out[k] = ((uint32_t)(w[k] >> 32) == tag) ? __uint_as_float((uint32_t)w[k]) : ll_load(src[k], tag, timeout_ns, ok);
}
unsigned long long wc_xch_timeout_ns();
