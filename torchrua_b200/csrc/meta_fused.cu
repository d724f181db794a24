// K0f -- one-launch metadata for batches of up to 8192 sequences (BASELINE configs 1, 2 and 4's
// per-step micro-batches): ONE CTA computes, from the lengths alone,
//     off (exclusive scan), N, T,
//     sorted / unsorted (stable descending argsort), batch_sizes and poff
// so that `C.pack()` needs one kernel launch and one D2H instead of the reference's CPU sort, two
// device<->host round trips and a BxT int64 mask (torchrua/core/view.py:47-58), and instead of the
// ~12 small launches of the general path (scan + 2-3 radix passes x 3 kernels + batch_sizes + scan).
//
// A single SM issues 4 warp-instructions per cycle, so the sort must be O(B) work: an in-CTA LSD radix
// sort, 8-bit digits, ceil(bits(T)/8) passes (T is known inside the kernel by then), warp-private
// digit histograms in shared memory, stable ranking inside a warp with __match_any_sync.  (A bitonic
// network over the same 4096 keys measured 55 us -- 78 stages x 4096 compare-exchanges are issue-bound
// on one SM; this version is an order of magnitude less work.)  Keys are T - len, so an ascending stable
// sort is the stable DESCENDING order by length.  batch_sizes[t] = first rank whose length is <= t is a
// binary search over the sorted lengths staged in shared memory.  The host reads
// [N, T, batch_sizes[0..cap)] back in a single copy and only needs a second round trip when T > cap.
#include "common.cuh"

namespace rua {

constexpr int kFusedThreads = 1024;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxB = 8192;
constexpr int kFusedRounds = kFusedMaxB / kFusedThreads;  // 8: elements per lane per radix pass
constexpr int kDigits = 256;

__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t* s_part, int tid, int64_t* total) {
  const int lane = tid & 31, warp = tid >> 5;
  int64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t o = shfl_i64(incl, max(lane - d, 0));
    if (lane >= d) incl += o;
  }
  __syncthreads();  // s_part may still be in use by a previous call
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  int64_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kFusedWarps; ++w) {
    int64_t x = s_part[w];
    if (w < warp) base += x;
    tot += x;
  }
  if (total) *total = tot;
  return base + incl - v;
}

__global__ void __launch_bounds__(kFusedThreads)
meta_fused_kernel(const int64_t* __restrict__ len, int B, int Bpad, int64_t* __restrict__ off,
                  int64_t* __restrict__ sorted, int64_t* __restrict__ unsorted, int64_t* __restrict__ hostbuf,
                  int64_t* __restrict__ poff, int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* skey = reinterpret_cast<uint32_t*>(smem_raw);                  // Bpad: sorted lengths
  uint16_t* va = reinterpret_cast<uint16_t*>(smem_raw + 4 * (size_t)Bpad);  // Bpad: permutation ping
  uint16_t* vb = va + Bpad;                                                 // Bpad: permutation pong
  __shared__ uint32_t whist[kFusedWarps][kDigits];                          // 32 KB
  __shared__ uint32_t s_dbase[kDigits];
  __shared__ uint32_t s_wsum[8];
  __shared__ int64_t s_part[kFusedWarps];
  __shared__ int64_t s_max[kFusedWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- exclusive scan + max over the lengths ------------------------------------------------------
  const int per = (B + kFusedThreads - 1) / kFusedThreads;  // contiguous items per thread (<= 8)
  const int i0 = tid * per;
  int64_t tsum = 0, tmax = 0;
  for (int k = 0; k < per; ++k) {
    const int i = i0 + k;
    if (i < B) {
      const int64_t v = len[i];
      tsum += v;
      tmax = v > tmax ? v : tmax;
    }
  }
  int64_t wmax = tmax;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int64_t o = shfl_i64(wmax, lane ^ d);
    wmax = o > wmax ? o : wmax;
  }
  if (lane == 0) s_max[warp] = wmax;
  int64_t total = 0;
  int64_t run = block_excl_scan(tsum, s_part, tid, &total);
  int64_t T = 0;
#pragma unroll
  for (int w = 0; w < kFusedWarps; ++w) T = s_max[w] > T ? s_max[w] : T;
  for (int k = 0; k < per; ++k) {
    const int i = i0 + k;
    if (i < B) {
      off[i] = run;
      run += len[i];
    }
  }
  if (tid == 0) {
    off[B] = total;
    hostbuf[0] = total;
    hostbuf[1] = T;
  }
  if (sorted == nullptr) return;

  // ---- stable LSD radix sort of the permutation by key = T - len ---------------------------------
  for (int i = tid; i < B; i += kFusedThreads) va[i] = (uint16_t)i;
  int bits = 0;
  while (bits < 32 && (T >> bits) != 0) ++bits;
  const int passes = bits == 0 ? 1 : (bits + 7) / 8;
  const int chunk = ((B + kFusedWarps - 1) / kFusedWarps + 31) & ~31;  // elements per warp, multiple of 32
  const int rounds = chunk / 32;                                       // <= kFusedRounds
  const unsigned lt = (1u << lane) - 1u;
  uint16_t* src = va;
  uint16_t* dst = vb;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    for (int i = tid; i < kFusedWarps * kDigits; i += kFusedThreads) (&whist[0][0])[i] = 0;
    __syncthreads();
    uint32_t loc[kFusedRounds];
    uint16_t val[kFusedRounds];
    uint16_t dig[kFusedRounds];
#pragma unroll
    for (int r = 0; r < kFusedRounds; ++r) {
      if (r < rounds) {
        const int pos = warp * chunk + r * 32 + lane;
        const bool ok = pos < B;
        uint32_t v = 0, d = kDigits;
        if (ok) {
          v = src[pos];
          d = ((uint32_t)(T - __ldg(len + v)) >> shift) & (kDigits - 1);
        }
        const unsigned peers = __match_any_sync(kFullMask, d);
        const uint32_t rank = __popc(peers & lt);
        const uint32_t old = ok ? whist[warp][d] : 0;
        __syncwarp();
        if (ok && rank == 0) whist[warp][d] = old + __popc(peers);
        __syncwarp();
        loc[r] = old + rank;
        val[r] = (uint16_t)v;
        dig[r] = (uint16_t)d;
      }
    }
    __syncthreads();
    uint32_t dtotal = 0;
    if (tid < kDigits) {  // per digit: exclusive scan over the warps (keeps warp order => stability)
#pragma unroll 8
      for (int w = 0; w < kFusedWarps; ++w) {
        const uint32_t c = whist[w][tid];
        whist[w][tid] = dtotal;
        dtotal += c;
      }
      // exclusive scan of the 256 digit totals: shuffle scan inside each of the 8 warps ...
      uint32_t incl = dtotal;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_wsum[warp] = incl;
      s_dbase[tid] = incl - dtotal;
    }
    __syncthreads();
    if (tid < kDigits) {  // ... plus the sums of the warps before
      uint32_t add = 0;
      for (int w = 0; w < warp; ++w) add += s_wsum[w];
      s_dbase[tid] += add;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kFusedRounds; ++r) {
      if (r < rounds && dig[r] < kDigits) dst[s_dbase[dig[r]] + whist[warp][dig[r]] + loc[r]] = val[r];
    }
    __syncthreads();
    uint16_t* t = src; src = dst; dst = t;
  }
  // src now holds the sorted permutation
  for (int r = tid; r < B; r += kFusedThreads) {
    const uint32_t i = src[r];
    sorted[r] = i;
    unsorted[i] = r;
    skey[r] = (uint32_t)__ldg(len + i);  // non-increasing in r
  }
  __syncthreads();

  // ---- batch_sizes[t] = #{r : len_r > t} for t < min(T, cap), and their exclusive prefix sums -----
  const int Tc = (int)(T < cap ? T : cap);
  const int tper = (Tc + kFusedThreads - 1) / kFusedThreads;
  const int t0 = tid * tper;
  int64_t bsum = 0;
  for (int k = 0; k < tper; ++k) {
    const int t = t0 + k;
    if (t < Tc) {
      int lo = 0, hi = B;  // first rank whose length is <= t
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skey[mid] > (uint32_t)t) lo = mid + 1; else hi = mid;
      }
      hostbuf[2 + t] = lo;
      bsum += lo;
    }
  }
  int64_t prun = block_excl_scan(bsum, s_part, tid, nullptr);
  for (int k = 0; k < tper; ++k) {
    const int t = t0 + k;
    if (t < Tc) {
      poff[t] = prun;
      prun += hostbuf[2 + t];  // written by this same thread above
    }
  }
  if (tid == 0) poff[Tc] = Tc == T ? total : -1;  // exact only when nothing was cut off by cap
  // entries beyond T are N as well, so that a consumer launched BEFORE the host has read T (speculative C -> P
  // conversion) can search poff[0 .. cap] as if there were cap time steps: the extra ones are empty
  if (Tc == T)
    for (int t = Tc + 1 + tid; t <= cap; t += kFusedThreads) poff[t] = total;
}

}  // namespace rua

using namespace rua;

extern "C" {

int64_t rua_meta_fused_max_batch(void) { return kFusedMaxB; }

int rua_meta_fused(const int64_t* len, int64_t B, int64_t* off, int64_t* sorted, int64_t* unsorted,
                   int64_t* hostbuf, int64_t* poff, int64_t cap, rua_stream_t stream) {
  if (B <= 0 || B > kFusedMaxB || cap < 0 || cap > (1 << 20)) return RUA_ERR_INVALID;
  if (!len || !off || !hostbuf) return RUA_ERR_INVALID;
  if (sorted && (!unsorted || !poff)) return RUA_ERR_INVALID;
  int Bpad = 32;
  while (Bpad < B) Bpad <<= 1;
  size_t smem = (size_t)Bpad * 8;  // skey (4 B) + two uint16 permutation buffers
  static bool configured = false;
  if (!configured) {
    int rc = check_cuda(cudaFuncSetAttribute(meta_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kFusedMaxB * 8));
    if (rc) return rc;
    configured = true;
  }
  meta_fused_kernel<<<1, kFusedThreads, smem, (cudaStream_t)stream>>>(len, (int)B, Bpad, off, sorted, unsorted, hostbuf,
                                                                      poff, (int)cap);
  return check_launch();
}

}  // extern "C"
