// Shared device helpers for the rua_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rua_b200.h"

namespace rua {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

extern int g_last_cuda_error;
extern long long g_launch_count;

// per-device counter of out-of-range explicit indices (index.cu); nullptr if it cannot be allocated
unsigned long long* index_error_counter();

inline int check_launch() {
  ++g_launch_count;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    (void)cudaGetLastError();
    return RUA_ERR_CUDA;
  }
  return RUA_OK;
}

inline int check_cuda(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    (void)cudaGetLastError();
    return RUA_ERR_CUDA;
  }
  return RUA_OK;
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Largest s in [0, S) with off[s] <= j, where off is non-decreasing, off[0] == 0 <= j < off[S].
// With empty segments (off[s] == off[s+1]) this returns the unique segment that owns row j.
template <typename OffFn>
__device__ __forceinline__ int64_t owner_search(OffFn off, int64_t S, int64_t j) {
  int64_t lo = 0, hi = S;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (off(mid) <= j) lo = mid; else hi = mid;
  }
  return lo;
}

// The same search done by a whole warp: 32 probes per step instead of 1, so a tile prologue costs
// ~log32(S) dependent memory round trips instead of log2(S).
template <typename OffFn>
__device__ __forceinline__ int64_t warp_owner_search(OffFn off, int64_t S, int64_t j, int lane) {
  int64_t lo = 0, hi = S;
  while (hi - lo > 1) {
    const int64_t n = hi - lo - 1;                 // candidates lo+1 .. hi-1
    const int64_t step = (n + 31) / 32;
    const int64_t probe = lo + (int64_t)(lane + 1) * step;
    const bool ok = probe < hi && off(probe) <= j;
    const int k = __popc(__ballot_sync(kFullMask, ok));   // probes are monotone: k leading successes
    const int64_t new_lo = lo + (int64_t)k * step;
    const int64_t next = new_lo + step;
    hi = next < hi ? next : hi;
    lo = new_lo;
  }
  return lo;
}

struct GlobalOff {
  const int64_t* __restrict__ p;
  __device__ __forceinline__ int64_t operator()(int64_t i) const { return __ldg(p + i); }
};

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
  int lo = __shfl_sync(kFullMask, (int)(v & 0xffffffffll), src);
  int hi = __shfl_sync(kFullMask, (int)(v >> 32), src);
  return ((int64_t)hi << 32) | (uint32_t)lo;
}

}  // namespace rua
