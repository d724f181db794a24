// shared by reduce_warpseg.cu (forward) and reduce_warpseg_bwd.cu (backward): window geometry and the TMA bulk-copy /
// mbarrier helpers of the warp-per-32-segments kernels.
#pragma once

#include "reduce_common.cuh"
#include "tma.cuh"

namespace rua {

constexpr int kWsThreads = 128;                          // small CTAs: 1 M segments are only 4.4 waves of 256-thread CTAs
constexpr int kWsWarps = kWsThreads / 32;
constexpr int kWsChunkBytes = 4096;                      // 1024 fp32 scalars: the ~32 segments of a warp in ONE window
constexpr int kWsVecPerLane = kWsChunkBytes / 16 / 32;   // 8

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
