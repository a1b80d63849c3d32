// Peer-memory exchange protocol shared by wc_exchange.cu (stand-alone exchange kernels) and the chain kernel of wc_small.cu
// (adjoint all-reduce as a stage of the backward chain).  See wc_exchange.cu for the protocol description.
#pragma once
#include "wc_common.cuh"

constexpr int XCH_CTAS = 32, XCH_THREADS = 512, XCH_MAXR = 16;
constexpr size_t XCH_FLAG_BYTES = 1024;     // 2 directions x 2 slots x 16 ranks x u32 = 256 B, padded

struct XchLayout { size_t fwd_floats, bwd_floats, fwd_off, bwd_off, total; };

static inline XchLayout xch_layout(int C, int g, int R) {
  XchLayout L;
  const size_t gg = (size_t)C * g;
  L.fwd_floats = wc_align_up(C + gg + 1, 64);       // one padded shard record
  L.bwd_floats = wc_align_up(gg + C, 64);
  L.fwd_off = XCH_FLAG_BYTES;
  L.bwd_off = L.fwd_off + 2 * (size_t)R * L.fwd_floats * 4;
  L.total = L.bwd_off + 2 * (size_t)R * L.bwd_floats * 4;
  return L;
}

__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long xch_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded spin until *f >= target (wrap-safe); false on time-out
__device__ __forceinline__ bool xch_spin(const uint32_t* f, uint32_t target, unsigned long long timeout_ns) {
  const unsigned long long t0 = timeout_ns ? xch_now() : 0ull;
  uint32_t polls = 0;
  while ((int32_t)(ld_acquire_sys(f) - target) < 0) {
    if (timeout_ns && ((++polls & 1023u) == 0) && xch_now() - t0 > timeout_ns) return false;
  }
  return true;
}
unsigned long long wc_xch_timeout_ns();
