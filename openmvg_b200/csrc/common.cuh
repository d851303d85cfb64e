// common.cuh — error plumbing and the sm_100a PTX wrappers shared by match.cu and ba.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <map>
#include <mutex>
#include <unordered_map>

#include "../../include/omvg_b200.h"

namespace omvg {

inline std::string &last_error() { static thread_local std::string e; return e; }
inline int fail(int code, const char *fmt, ...) {
  char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  last_error() = buf; return code;
}
#define OMVG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return ::omvg::fail(OMVG_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

// ----------------------------------------------------------------------------- caching device allocator
// Adjust()/Match() are called many times by the SfM engines (BA / outlier-rejection / BA loops): cudaMalloc and
// cudaFree of ~100 work buffers cost more than a whole solve, so released buffers are kept per device and
// handed out again (best fit within 25 % slack).  omvg_trim_cache() returns everything to the driver.
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void *> free_[16];
  std::unordered_map<void *, std::pair<size_t, int>> live;
  size_t cached = 0;
};
inline DevPool &dev_pool() { static DevPool p; return p; }
inline cudaError_t pool_malloc(void **out, size_t bytes) {
  int dev = 0; cudaGetDevice(&dev); dev &= 15;
  bytes = (bytes + 511) & ~size_t(511);
  DevPool &P = dev_pool();
  { std::lock_guard<std::mutex> g(P.mu);
    auto it = P.free_[dev].lower_bound(bytes);
    if (it != P.free_[dev].end() && it->first <= bytes + bytes / 4 + 4096) {
      *out = it->second; P.live[*out] = {it->first, dev}; P.cached -= it->first; P.free_[dev].erase(it); return cudaSuccess; } }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {                                   // out of memory: drop the cache and retry once
    { std::lock_guard<std::mutex> g(P.mu); for (auto &kv : P.free_[dev]) cudaFree(kv.second); for (auto &kv : P.free_[dev]) P.cached -= kv.first; P.free_[dev].clear(); }
    cudaGetLastError(); e = cudaMalloc(out, bytes);
  }
  if (e == cudaSuccess) { std::lock_guard<std::mutex> g(P.mu); P.live[*out] = {bytes, dev}; }
  return e;
}
inline void pool_free(void *p) {
  if (!p) return;
  DevPool &P = dev_pool();
  std::lock_guard<std::mutex> g(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) { cudaFree(p); return; }
  const size_t bytes = it->second.first; const int dev = it->second.second; P.live.erase(it);
  if (P.cached + bytes > (size_t(16) << 30)) { cudaFree(p); return; }
  P.free_[dev].emplace(bytes, p); P.cached += bytes;
}
inline void pool_trim() {
  DevPool &P = dev_pool();
  std::lock_guard<std::mutex> g(P.mu);
  for (int d = 0; d < 16; ++d) { for (auto &kv : P.free_[d]) cudaFree(kv.second); P.free_[d].clear(); }
  P.cached = 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU box (a hang is a gpurun strike).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) { printf("mbar_wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); } }
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void *tmap) { asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory"); }

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival when all MMAs issued so far by this thread finish.
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, u8 x u8 -> s32   (SASS: UTCIMMA)
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base+i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, int32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr) : "memory");
}

// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes under the
// 128-byte swizzle (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B): 8-row x 128-B atoms, atom
// stride (SBO) 1024 B; LBO is unused for swizzled K-major; version 1 (sm_100); layout 2 = SW128.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)0 << 16;                              // LBO            [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO            [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::i8: D=S32, A=B=UINT8, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc_u8(int M, int N) {
  return (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace omvg
