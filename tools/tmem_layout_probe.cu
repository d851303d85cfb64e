// Probe: which (lane=row, column) does each thread/register receive for tcgen05.ld shapes on sm_100a?
// One CTA computes D[i][n] = i + 128*n with a single u8 tcgen05.mma (M128 N256 K32) and dumps the registers.
#include "../openmvg_b200/csrc/common.cuh"
#include <cstdio>
#include <vector>
using namespace omvg;

__device__ __forceinline__ uint32_t sw128_off(int row, int byte) {   // K-major, 128-B rows, 128-B swizzle
  const int chunk = byte >> 4, within = byte & 15;
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + within);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int *out /* [3 shapes][128 threads][8] */) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a = smem, *b = smem + 16384;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 32768);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  for (int i = threadIdx.x; i < 16384 + 32768; i += 128) smem[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += 128) { a[sw128_off(r, 0)] = (uint8_t)r; a[sw128_off(r, 1)] = 128; }
  for (int r = threadIdx.x; r < 256; r += 128) { b[sw128_off(r, 0)] = 1; b[sw128_off(r, 1)] = (uint8_t)r; }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 256);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    umma_i8(tb, make_kmajor_sw128_desc(smem_u32(a)), make_kmajor_sw128_desc(smem_u32(b)), make_idesc_u8(128, 256), 0);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t t0 = tb + ((warp * 32) << 16);
  int r[8];
  // shape A: 32x32b.x8  (thread = lane, 8 consecutive columns)
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 8; ++k) out[(0 * 128 + threadIdx.x) * 8 + k] = r[k];
  // shape B: 16x256b.x2 at lane offset 0 : 8 registers
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 8; ++k) out[(1 * 128 + threadIdx.x) * 8 + k] = r[k];
  // shape C: 16x256b.x2 at lane offset 16
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0 + (16u << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 8; ++k) out[(2 * 128 + threadIdx.x) * 8 + k] = r[k];
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 256);
}

int main() {
  int *d; cudaMalloc(&d, 3 * 128 * 8 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  probe_kernel<<<1, 128, 60000>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  std::vector<int> h(3 * 128 * 8); cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  const char *names[3] = {"32x32b.x8 @lane0", "16x256b.x2 @lane0", "16x256b.x2 @lane16"};
  for (int s = 0; s < 3; ++s) {
    printf("== %s : thread -> (row,col) per register\n", names[s]);
    for (int t = 0; t < 40; ++t) { if (t == 36) t = 124; printf(" t%3d:", t); for (int k = 0; k < 8; ++k) { int v = h[(s * 128 + t) * 8 + k]; printf(" (%3d,%3d)", v % 128, v / 128); } printf("\n"); }
  }
  return 0;
}
