// K0 -- metadata kernels: lengths -> offsets / max / stable descending sort / batch_sizes, and back.
//
// Replaces (reference file:line): get_offsets torchrua/utils.py:16-19, invert_permutation
// utils.py:22-26, size() layout/cat.py:61-66, pack_view core/view.py:47-58 (CPU torch.sort + D2H/H2D
// + a BxT int64 mask), cat_view/left_view/right_view token_sizes core/view.py:21-38,67-71.
//
// All of this is a few MB of int64 at most (config 5: B = 1M -> 8 MB), so the design goal is few
// launches and no host syncs rather than peak bandwidth:
//   * scan: one-pass chained scan (decoupled look-back), 2048 items per CTA, fused max reduction;
//   * sort: LSD radix sort, 8-bit digits, ceil(bits(T)/8) passes; keys are T - len so that an
//     ascending stable sort yields the stable DESCENDING order by length;
//   * batch_sizes / lengths_from_pack: one binary search per output element on a monotone array.
#include "common.cuh"

namespace rua {

int g_last_cuda_error = 0;
long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------------
// exclusive scan + max
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItemsOne = 16;                          // up to 4096 lengths: ONE CTA, no look-back, no memset
constexpr int kScanItemsMany = 4;                          // larger vectors: 1024 per CTA -- B = 1 M is 977 CTAs instead of
                                                           // 245 (1.65 CTAs per SM: the chained scan was latency-bound)
constexpr int kScanTileOne = kScanThreads * kScanItemsOne;
constexpr int kScanTileMany = kScanThreads * kScanItemsMany;

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t o = shfl_i64(v, max(lane - d, 0));
    if (lane >= d) v += o;
  }
  return v;
}

__device__ __forceinline__ int64_t warp_sum(int64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int lo = __shfl_xor_sync(kFullMask, (int)(v & 0xffffffffll), d);
    int hi = __shfl_xor_sync(kFullMask, (int)(v >> 32), d);
    v += ((int64_t)hi << 32) | (uint32_t)lo;
  }
  return v;
}

__device__ __forceinline__ int64_t warp_max(int64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int lo = __shfl_xor_sync(kFullMask, (int)(v & 0xffffffffll), d);
    int hi = __shfl_xor_sync(kFullMask, (int)(v >> 32), d);
    int64_t o = ((int64_t)hi << 32) | (uint32_t)lo;
    v = o > v ? o : v;
  }
  return v;
}

// status word: (value << 2) | flag; flag 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix
// optional extras of the scan (rua_scan_lengths_ex):
//   * lengths derived on the fly from a PackedSequence's (batch_sizes, unsorted_indices): len[i] = first t with
//     bs[t] <= unsorted[i] (bs is non-increasing) -- P -> token_sizes and its prefix sum in ONE launch;
//   * completion notice straight into pinned host memory: the last CTA to finish writes [sum, max, ticket] through
//     the UVA mapping, so the host learns N and T by polling one cache line instead of enqueueing a device->host copy
//     and synchronising the stream (the output shapes of most conversions depend on them).
struct ScanExtras {
  const int64_t* bs;
  const int64_t* unsorted;
  int64_t Tp;
  int64_t* len_out;
  volatile int64_t* notify;   // device-accessible pinned host memory, 3 x int64, or NULL
  int64_t ticket;
  int64_t tiles;
};

template <int kScanItems>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const int64_t* __restrict__ in, int64_t n, int64_t clamp_max, int64_t* __restrict__ out,
            int64_t* __restrict__ stats, unsigned long long* status, int multi_tile, const ScanExtras ex) {
  constexpr int kScanTile = kScanThreads * kScanItems;
  __shared__ int64_t s_warp[kScanThreads / 32];
  __shared__ int64_t s_wmax[kScanThreads / 32];
  __shared__ int64_t s_prefix;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  const int64_t base = tile * kScanTile + (int64_t)tid * kScanItems;

  int64_t v[kScanItems];
  int64_t tsum = 0, tmax = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t idx = base + k;
    if (in) {
      v[k] = idx < n ? in[idx] : 0;
    } else {
      v[k] = 0;
      if (idx < n) {
        const int64_t r = __ldg(ex.unsorted + idx);
        int64_t lo = 0, hi = ex.Tp;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(ex.bs + mid) > r) lo = mid + 1; else hi = mid;
        }
        v[k] = lo;
        ex.len_out[idx] = lo;
      }
    }
    tsum += v[k];
    tmax = v[k] > tmax ? v[k] : tmax;
  }
  int64_t incl = warp_incl_scan(tsum, lane);
  int64_t wmax = warp_max(tmax);
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0) s_wmax[warp] = wmax;
  __syncthreads();
  int64_t warp_base = 0, agg = 0, bmax = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    int64_t x = s_warp[w];
    if (w < warp) warp_base += x;
    agg += x;
    bmax = s_wmax[w] > bmax ? s_wmax[w] : bmax;
  }

  if (warp == 0) {
    int64_t excl = 0;
    if (multi_tile && tile > 0) {
      if (lane == 0) *(volatile unsigned long long*)(status + tile) = ((unsigned long long)agg << 2) | 1ull;
      int64_t pred = tile - 1;
      while (true) {
        int64_t idx = pred - lane;
        unsigned long long st = 2ull;  // virtual tile before the first: prefix 0
        if (idx >= 0) {
          do { st = *(volatile unsigned long long*)(status + idx); } while ((st & 3ull) == 0ull);
        }
        unsigned done = __ballot_sync(kFullMask, (st & 3ull) == 2ull);
        int first = done ? (__ffs(done) - 1) : 32;
        int64_t x = lane <= first ? (int64_t)(st >> 2) : 0;
        excl += warp_sum(x);
        if (done) break;
        pred -= 32;
      }
    }
    if (lane == 0) {
      if (multi_tile)
        *(volatile unsigned long long*)(status + tile) = ((unsigned long long)(excl + agg) << 2) | 2ull;
      s_prefix = excl;
    }
  }
  __syncthreads();

  int64_t run = s_prefix + warp_base + (incl - tsum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t idx = base + k;
    if (idx < n) out[idx] = run < clamp_max ? run : clamp_max;
    run += v[k];
    if (idx == n - 1) {
      out[n] = run < clamp_max ? run : clamp_max;
      stats[0] = run;
    }
  }
  if (tid == 0) {
    if (multi_tile) atomicMax((long long*)(stats + 1), (long long)bmax);
    else stats[1] = bmax;
  }
  if (ex.notify) {                                   // kernel-uniform
    __syncthreads();                                 // this CTA's stats contributions are issued
    if (tid == 0) {
      bool last = true;
      if (multi_tile) {
        __threadfence();
        last = atomicAdd(status + ex.tiles, 1ull) == (unsigned long long)(ex.tiles - 1);
        __threadfence();
      }
      if (last) {
        ex.notify[0] = *(volatile int64_t*)(stats);
        ex.notify[1] = *(volatile int64_t*)(stats + 1);
        __threadfence_system();
        ex.notify[2] = ex.ticket;
      }
    }
  }
}

__global__ void scan_empty_kernel(int64_t* out, int64_t* stats, volatile int64_t* notify, int64_t ticket) {
  out[0] = 0;
  stats[0] = 0;
  stats[1] = 0;
  if (notify) {
    notify[0] = 0;
    notify[1] = 0;
    __threadfence_system();
    notify[2] = ticket;
  }
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort of (T - len[i], i)
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 8;
constexpr int kSortTile = kSortThreads * kSortRounds;  // 2048 elements per CTA
constexpr int kRadix = 256;

struct SortSrc {
  const int64_t* len;     // pass 0: key = T - len[i] (descending sort) or len[i] (ascending), val = i
  const uint32_t* keys;   // later passes
  const uint32_t* vals;
  int64_t T;
  int ascending;
  __device__ __forceinline__ void load(int64_t i, uint32_t& k, uint32_t& v) const {
    if (len) { k = ascending ? (uint32_t)len[i] : (uint32_t)(T - len[i]); v = (uint32_t)i; }
    else { k = keys[i]; v = vals[i]; }
  }
};

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(SortSrc src, int64_t n, int shift, uint32_t* __restrict__ table, int nblk) {
  __shared__ uint32_t hist[kRadix];
  hist[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) {
      uint32_t k, v;
      src.load(i, k, v);
      atomicAdd(&hist[(k >> shift) & (kRadix - 1)], 1u);
    }
  }
  __syncthreads();
  table[(size_t)threadIdx.x * nblk + blockIdx.x] = hist[threadIdx.x];
}

// The digit table is digit-major: table[d * nblk + b] = count of digit d in block b.  Its exclusive scan
// is done in two levels so that no single CTA walks all 256 * nblk entries (a 1M-element sort has 125 K of
// them: 106 us on one CTA): here one CTA per digit scans its own row in place and records the row total;
// the scatter kernel turns the 256 totals into digit bases with a block scan of its own.
__global__ void __launch_bounds__(256) table_row_scan_kernel(uint32_t* table, int nblk, uint32_t* __restrict__ totals) {
  __shared__ uint32_t s_warp[8];
  uint32_t* row = table + (size_t)blockIdx.x * nblk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < nblk; base += 256) {
    const int i = base + tid;
    const uint32_t x = i < nblk ? row[i] : 0;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t t = s_warp[w];
      if (w < warp) wbase += t;
      total += t;
    }
    if (i < nblk) row[i] = carry + wbase + incl - x;
    carry += total;
    __syncthreads();
  }
  if (tid == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(SortSrc src, int64_t n, int shift, const uint32_t* __restrict__ table,
                     const uint32_t* __restrict__ totals, int nblk, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     int64_t* __restrict__ sorted, int64_t* __restrict__ unsorted) {
  __shared__ uint32_t whist[kSortWarps][kRadix];
  __shared__ uint32_t s_dwarp[kSortWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&whist[0][0])[i] = 0;
  // digit base = exclusive scan of the 256 row totals (one digit per thread)
  const uint32_t dtotal = totals[tid];
  uint32_t dincl = dtotal;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(kFullMask, dincl, d);
    if (lane >= d) dincl += o;
  }
  if (lane == 31) s_dwarp[warp] = dincl;
  __syncthreads();
  uint32_t dbase = dincl - dtotal;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w)
    if (w < warp) dbase += s_dwarp[w];

  // warp w owns the contiguous range [base + w*256, base + (w+1)*256), visited in 8 rounds of 32
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (32 * kSortRounds);
  uint32_t key[kSortRounds], val[kSortRounds], loc[kSortRounds];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t k = 0, v = 0;
    if (ok) src.load(i, k, v);
    key[r] = k; val[r] = v;
    uint32_t d = ok ? ((k >> shift) & (kRadix - 1)) : kRadix;  // kRadix = "no element"
    unsigned peers = __match_any_sync(kFullMask, d);
    uint32_t rank = __popc(peers & lt);
    uint32_t old = ok ? whist[warp][d] : 0;
    __syncwarp();
    if (ok && rank == 0) whist[warp][d] = old + __popc(peers);
    __syncwarp();
    loc[r] = old + rank;
  }
  __syncthreads();
  {  // one thread per digit: turn per-warp counts into global start positions
    uint32_t run = dbase + table[(size_t)tid * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = whist[w][tid];
      whist[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      uint32_t d = (key[r] >> shift) & (kRadix - 1);
      uint32_t pos = whist[warp][d] + loc[r];
      if (sorted) {  // last pass: emit the permutation and its inverse directly
        sorted[pos] = (int64_t)val[r];
        unsorted[val[r]] = (int64_t)pos;
      } else {
        keys_out[pos] = key[r];
        vals_out[pos] = val[r];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small element-wise metadata kernels
// ------------------------------------------------------------------------------------------------
__global__ void invert_perm_kernel(const int64_t* __restrict__ perm, int64_t n, int64_t* __restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[perm[j]] = j;
}

// bs[t] = first r with len[sorted[r]] <= t  (len o sorted is non-increasing).  One WARP per time step:
// a 32-ary search needs log32(B) rounds of two dependent loads instead of log2(B) (B = 1M: 4 rounds, not 20;
// this kernel is pure latency -- T searches over an L2-resident array).
__global__ void __launch_bounds__(256)
batch_sizes_kernel(const int64_t* __restrict__ len, const int64_t* __restrict__ sorted, int64_t B, int64_t T,
                   int64_t* __restrict__ bs) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  int64_t lo = 0, hi = B;  // invariant: all r < lo have len > t; all r >= hi have len <= t
  while (lo < hi) {
    const int64_t n = hi - lo;
    const int64_t step = (n + 31) / 32;
    const int64_t probe = lo + (int64_t)lane * step;       // probes lo, lo+step, ... (monotone predicate)
    const bool gt = probe < hi && __ldg(len + __ldg(sorted + probe)) > t;
    const int k = __popc(__ballot_sync(kFullMask, gt));    // k leading probes still have len > t
    if (k == 0) { hi = lo; break; }
    const int64_t last_gt = lo + (int64_t)(k - 1) * step;   // len > t here
    const int64_t nxt = last_gt + step;                     // first probe with len <= t (or beyond hi)
    lo = last_gt + 1;
    hi = nxt < hi ? nxt : hi;
  }
  if (lane == 0) bs[t] = lo;
}

// len[i] = first t with bs[t] <= unsorted[i]  (bs is non-increasing)
__global__ void lengths_from_pack_kernel(const int64_t* __restrict__ bs, const int64_t* __restrict__ unsorted,
                                         int64_t B, int64_t T, int64_t* __restrict__ len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int64_t r = unsorted[i];
  int64_t lo = 0, hi = T;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(bs + mid) > r) lo = mid + 1; else hi = mid;
  }
  len[i] = lo;
}

// off[m] = first rank r with keys[sorted[r]] >= m  (keys o sorted is non-decreasing), m in [0, M]
__global__ void bucket_offsets_kernel(const int64_t* __restrict__ keys, const int64_t* __restrict__ sorted,
                                      int64_t n, int64_t M, int64_t* __restrict__ off) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m > M) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + __ldg(sorted + mid)) < m) lo = mid + 1; else hi = mid;
  }
  off[m] = lo;
}

}  // namespace rua

using namespace rua;

extern "C" {

int rua_version(void) { return 100; }

const char* rua_error_string(int status) {
  switch (status) {
    case RUA_OK: return "ok";
    case RUA_ERR_INVALID: return "invalid argument";
    case RUA_ERR_WORKSPACE: return "workspace too small";
    case RUA_ERR_UNSUPPORTED: return "unsupported size or dtype";
    case RUA_ERR_CUDA: return "CUDA runtime error";
    default: return "unknown status";
  }
}

int rua_last_cuda_error(void) { return g_last_cuda_error; }
int64_t rua_launch_count(void) { return (int64_t)g_launch_count; }

size_t rua_scan_workspace_bytes(int64_t n) {
  int64_t tiles = n > kScanTileOne ? ceil_div(n, kScanTileMany) : 1;
  return (size_t)(tiles + 1) * sizeof(unsigned long long);   // tile status words + the completion counter
}

int rua_scan_lengths(const int64_t* sizes, int64_t n, int64_t clamp_max, int64_t* off, int64_t* stats, void* ws,
                     size_t ws_bytes, rua_stream_t stream) {
  return rua_scan_lengths_ex(sizes, n, clamp_max, off, stats, ws, ws_bytes, nullptr, nullptr, 0, nullptr, nullptr, 0,
                             stream);
}

int rua_scan_lengths_ex(const int64_t* sizes, int64_t n, int64_t clamp_max, int64_t* off, int64_t* stats, void* ws,
                        size_t ws_bytes, const int64_t* bs, const int64_t* unsorted, int64_t Tp, int64_t* len_out,
                        int64_t* notify_host_mapped, int64_t ticket, rua_stream_t stream) {
  if (n < 0 || !off || !stats) return RUA_ERR_INVALID;
  if (n > 0 && !sizes && (!unsorted || !len_out || Tp < 0 || (Tp > 0 && !bs))) return RUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  ScanExtras ex{bs, unsorted, Tp, len_out, notify_host_mapped, ticket, 0};
  if (n == 0) {
    scan_empty_kernel<<<1, 1, 0, st>>>(off, stats, ex.notify, ticket);
    return check_launch();
  }
  int64_t tiles = n > kScanTileOne ? ceil_div(n, kScanTileMany) : 1;
  ex.tiles = tiles;
  int multi = tiles > 1;
  if (multi) {
    if (!ws || ws_bytes < rua_scan_workspace_bytes(n)) return RUA_ERR_WORKSPACE;
    int rc;
    if ((char*)stats + 2 * sizeof(int64_t) == (char*)ws) {
      // stats sits right in front of the status words (the layout torchrua_b200 uses): one memset
      rc = check_cuda(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t) + (size_t)(tiles + 1) * sizeof(unsigned long long), st));
      if (rc) return rc;
    } else {
      rc = check_cuda(cudaMemsetAsync(ws, 0, (size_t)(tiles + 1) * sizeof(unsigned long long), st));
      if (rc) return rc;
      rc = check_cuda(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), st));
      if (rc) return rc;
    }
  }
  if (multi) scan_kernel<kScanItemsMany><<<(unsigned)tiles, kScanThreads, 0, st>>>(sizes, n, clamp_max, off, stats, (unsigned long long*)ws, multi, ex);
  else scan_kernel<kScanItemsOne><<<1, kScanThreads, 0, st>>>(sizes, n, clamp_max, off, stats, (unsigned long long*)ws, multi, ex);
  return check_launch();
}

size_t rua_sort_workspace_bytes(int64_t B) {
  if (B <= 0) return 16;
  int64_t nblk = ceil_div(B, kSortTile);
  // 2 x (keys, vals) ping-pong + digit table + digit totals
  return (size_t)(4 * B + (int64_t)kRadix * nblk + kRadix) * sizeof(uint32_t) + 64;
}

static int sort_impl(const int64_t* len, int64_t B, int64_t T, int ascending, int64_t* sorted, int64_t* unsorted,
                     void* ws, size_t ws_bytes, rua_stream_t stream);

int rua_sort_lengths(const int64_t* len, int64_t B, int64_t T, int64_t* sorted, int64_t* unsorted,
                     void* ws, size_t ws_bytes, rua_stream_t stream) {
  return sort_impl(len, B, T, 0, sorted, unsorted, ws, ws_bytes, stream);
}

int rua_sort_keys(const int64_t* keys, int64_t n, int64_t max_key, int64_t* sorted, int64_t* unsorted,
                  void* ws, size_t ws_bytes, rua_stream_t stream) {
  return sort_impl(keys, n, max_key, 1, sorted, unsorted, ws, ws_bytes, stream);
}

static int sort_impl(const int64_t* len, int64_t B, int64_t T, int ascending, int64_t* sorted, int64_t* unsorted,
                     void* ws, size_t ws_bytes, rua_stream_t stream) {
  if (B < 0 || T < 0) return RUA_ERR_INVALID;
  if (B == 0) return RUA_OK;
  if (!len || !sorted || !unsorted || !ws) return RUA_ERR_INVALID;
  if (B >= (1ll << 31) || T >= (1ll << 32)) return RUA_ERR_UNSUPPORTED;
  if (ws_bytes < rua_sort_workspace_bytes(B)) return RUA_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = (int)ceil_div(B, kSortTile);
  uint32_t* w = (uint32_t*)ws;
  uint32_t* keys[2] = {w, w + B};
  uint32_t* vals[2] = {w + 2 * B, w + 3 * B};
  uint32_t* table = w + 4 * B;
  uint32_t* totals = table + (size_t)kRadix * nblk;

  int bits = 0;
  while (bits < 32 && (T >> bits) != 0) ++bits;
  int passes = bits == 0 ? 1 : (bits + 7) / 8;
  for (int p = 0; p < passes; ++p) {
    SortSrc src;
    src.T = T;
    src.ascending = ascending;
    if (p == 0) { src.len = len; src.keys = nullptr; src.vals = nullptr; }
    else { src.len = nullptr; src.keys = keys[(p - 1) & 1]; src.vals = vals[(p - 1) & 1]; }
    int shift = 8 * p;
    bool last = p == passes - 1;
    radix_hist_kernel<<<nblk, kSortThreads, 0, st>>>(src, B, shift, table, nblk);
    int rc = check_launch();
    if (rc) return rc;
    table_row_scan_kernel<<<kRadix, 256, 0, st>>>(table, nblk, totals);
    rc = check_launch();
    if (rc) return rc;
    radix_scatter_kernel<<<nblk, kSortThreads, 0, st>>>(src, B, shift, table, totals, nblk, keys[p & 1], vals[p & 1],
                                                        last ? sorted : nullptr, last ? unsorted : nullptr);
    rc = check_launch();
    if (rc) return rc;
  }
  return RUA_OK;
}

int rua_invert_permutation(const int64_t* perm, int64_t B, int64_t* out, rua_stream_t stream) {
  if (B < 0) return RUA_ERR_INVALID;
  if (B == 0) return RUA_OK;
  if (!perm || !out) return RUA_ERR_INVALID;
  invert_perm_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(perm, B, out);
  return check_launch();
}

int rua_batch_sizes(const int64_t* len, const int64_t* sorted, int64_t B, int64_t T, int64_t* bs,
                    rua_stream_t stream) {
  if (B < 0 || T < 0) return RUA_ERR_INVALID;
  if (T == 0) return RUA_OK;
  if (!len || !sorted || !bs) return RUA_ERR_INVALID;
  batch_sizes_kernel<<<(unsigned)ceil_div(T, 8), 256, 0, (cudaStream_t)stream>>>(len, sorted, B, T, bs);
  return check_launch();
}

int rua_bucket_offsets(const int64_t* keys, const int64_t* sorted, int64_t n, int64_t M, int64_t* off,
                       rua_stream_t stream) {
  if (n < 0 || M < 0 || !off) return RUA_ERR_INVALID;
  if (n > 0 && (!keys || !sorted)) return RUA_ERR_INVALID;
  bucket_offsets_kernel<<<(unsigned)ceil_div(M + 1, 256), 256, 0, (cudaStream_t)stream>>>(keys, sorted, n, M, off);
  return check_launch();
}

int rua_lengths_from_pack(const int64_t* bs, const int64_t* unsorted, int64_t B, int64_t T, int64_t* len,
                          rua_stream_t stream) {
  if (B < 0 || T < 0) return RUA_ERR_INVALID;
  if (B == 0) return RUA_OK;
  if (!unsorted || !len || (T > 0 && !bs)) return RUA_ERR_INVALID;
  lengths_from_pack_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(bs, unsorted, B, T, len);
  return check_launch();
}

}  // extern "C"
