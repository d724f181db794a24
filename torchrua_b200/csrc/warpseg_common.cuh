// shared by reduce_warpseg.cu (forward) and reduce_warpseg_bwd.cu (backward): window geometry and the TMA bulk-copy /
// mbarrier helpers of the warp-per-32-segments kernels.
#pragma once

#include "reduce_common.cuh"

namespace rua {

constexpr int kWsThreads = 128;                          // small CTAs: 1 M segments are only 4.4 waves of 256-thread CTAs
constexpr int kWsWarps = kWsThreads / 32;
constexpr int kWsChunkBytes = 4096;                      // 1024 fp32 scalars: the ~32 segments of a warp in ONE window
constexpr int kWsVecPerLane = kWsChunkBytes / 16 / 32;   // 8

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared, completion on an mbarrier ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <typename A, int HE, int OP>
__device__ __forceinline__ State<A, HE, OP> ws_shfl_xor(const State<A, HE, OP>& v, int d) {
  State<A, HE, OP> o;
#pragma unroll
  for (int h = 0; h < HE; ++h) {
    o.a[h] = __shfl_xor_sync(kFullMask, v.a[h], d);
    if constexpr (OpInfo<OP>::kIsLse) o.s[h] = __shfl_xor_sync(kFullMask, v.s[h], d);
  }
  if constexpr (!OpInfo<OP>::kIsLse) o.s[0] = A(0);
  return o;
}

}  // namespace rua
