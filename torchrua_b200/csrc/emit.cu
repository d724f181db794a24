// K3 -- mask / index emit: write-only kernels, no payload reads.
//
// Replaces (reference file:line): mask / bmask / fmask torchrua/mask.py:6-32, get_mask
// core/view.py:11-18, major_sizes_to_ptr utils.py:7-13, C/L/R.ptr layout/cat.py:68-71,
// P.ptr layout/pack.py:23-27, L.idx layout/left.py:73-77, R.idx layout/right.py:74-79.
//
// The reference builds these from repeat_interleave (each an internal cumsum + a .item() sync),
// arange, new_full and index_put_.  Here every output element is a closed form of (segment, within)
// and is stored once with 16-byte coalesced stores.  HBM-write bound.
#include "common.cuh"

namespace rua {

// ------------------------------------------------------------------------------------------------
// mask: out[i, t] = t < len[i] ? one : zero          (left aligned for all layouts, mask.py:10)
// each thread produces 16 bytes = 16/E consecutive elements of the flattened (B, W) output
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void __launch_bounds__(256)
mask_kernel(const int64_t* __restrict__ len, int64_t B, int64_t W, E zero, E one, E* __restrict__ out) {
  constexpr int K = 16 / sizeof(E);
  const int64_t total = B * W;
  const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * K;
  if (g0 >= total) return;
  int64_t i, t;
  if (total < (1ll << 31)) {
    uint32_t q = (uint32_t)g0 / (uint32_t)W;
    i = q;
    t = (uint32_t)g0 - q * (uint32_t)W;
  } else {
    i = g0 / W;
    t = g0 - i * W;
  }
  int64_t li = __ldg(len + i);
  union { uint4 v; E e[K]; } u;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    u.e[k] = t < li ? one : zero;
    if (++t == W) {
      t = 0;
      ++i;
      li = i < B ? __ldg(len + i) : 0;
    }
  }
  if (g0 + K <= total) {
    __stcs(reinterpret_cast<uint4*>(out + g0), u.v);
  } else {
    for (int k = 0; g0 + k < total; ++k) out[g0 + k] = u.e[k];
  }
}

// ------------------------------------------------------------------------------------------------
// ptr / idx emit over a segmented range.  A CTA owns a tile of kEmitTile consecutive positions.
// The segment boundaries that fall inside the tile are staged in shared memory (one cooperative
// 32-ary search finds the first one), so each element resolves its segment with a short
// shared-memory binary search instead of a log2(S)-deep walk over global memory.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitGroup = 4;                                   // consecutive positions per thread = one 256-bit store
constexpr int kEmitRounds = 2;
constexpr int kEmitTile = kEmitThreads * kEmitGroup * kEmitRounds;  // 2048 positions per CTA
constexpr int kEmitCap = kEmitTile + 2;                            // segment starts staged per tile

// sm_100a has 256-bit global stores (SASS STG.E.ENL2.256): one full 32-byte sector per lane
__device__ __forceinline__ void st_v4_i64(int64_t* p, const int64_t* v) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}

__global__ void __launch_bounds__(kEmitThreads)
emit_ptr_kernel(const int64_t* __restrict__ off, int64_t S, int64_t n, const int64_t* __restrict__ relabel,
                int64_t* __restrict__ which, int64_t* __restrict__ within, int64_t* __restrict__ flat,
                int64_t stride, int right_align, int wide_stores) {
  __shared__ int s_rel[kEmitCap];          // tile-relative segment starts, clamped to [-1, tile + 1]
  __shared__ int64_t s_first, s_last;
  const int tid = threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.x * kEmitTile;
  const int nj = (int)(j0 + kEmitTile < n ? kEmitTile : n - j0);
  GlobalOff g{off};
  // two warp-cooperative 32-ary searches (log32 S dependent round trips instead of log2 S)
  if (tid < 32) {
    const int64_t a = warp_owner_search(g, S, j0, tid);
    if (tid == 0) s_first = a;
  } else if (tid < 64) {
    const int64_t b = warp_owner_search(g, S, j0 + nj - 1, tid - 32);
    if (tid == 32) s_last = b;
  }
  __syncthreads();
  const int64_t first = s_first, last = s_last;
  const int64_t cnt64 = last - first + 2;  // off[first .. last+1]
  const bool staged = cnt64 <= kEmitCap;   // many empty segments inside the tile can overflow the stage
  const int cnt = staged ? (int)cnt64 : 0;
  if (staged) {
    for (int k = tid; k < cnt; k += kEmitThreads) {
      const int64_t d = __ldg(off + first + k) - j0;
      s_rel[k] = d < -1 ? -1 : (d > kEmitTile + 1 ? kEmitTile + 1 : (int)d);
    }
  }
  __syncthreads();
  const int64_t off_first = __ldg(off + first);  // the only staged start that can lie before the tile

#pragma unroll
  for (int r = 0; r < kEmitRounds; ++r) {
    const int jb = (r * kEmitThreads + tid) * kEmitGroup;  // tile-relative, 4 consecutive positions
    if (jb >= nj) break;
    int64_t sv[kEmitGroup], wv[kEmitGroup], fv[kEmitGroup];
    if (staged) {
      int lo = 0, hi = cnt - 1;              // one search for the group's first position ...
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_rel[mid] <= jb) lo = mid; else hi = mid;
      }
#pragma unroll
      for (int k = 0; k < kEmitGroup; ++k) {  // ... then walk: a group rarely crosses more than one boundary
        const int jr = jb + k;
        while (lo + 2 < cnt && s_rel[lo + 1] <= jr) ++lo;
        const int64_t base = lo == 0 ? off_first : j0 + s_rel[lo];
        const int64_t j = j0 + jr;
        sv[k] = first + lo;
        wv[k] = j - base;
        if (flat) {
          const int64_t len = (s_rel[lo + 1] <= kEmitTile && lo > 0) ? (int64_t)(s_rel[lo + 1] - s_rel[lo])
                                                                      : __ldg(off + first + lo + 1) - base;
          fv[k] = sv[k] * stride + wv[k] + (right_align ? stride - len : 0);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < kEmitGroup; ++k) {
        const int64_t j = j0 + jb + k;
        const int64_t jj = j < n ? j : n - 1;
        const int64_t s = owner_search(g, S, jj);
        const int64_t base = __ldg(off + s);
        sv[k] = s;
        wv[k] = jj - base;
        fv[k] = s * stride + wv[k] + (right_align ? stride - (__ldg(off + s + 1) - base) : 0);
      }
    }
    if (relabel) {
#pragma unroll
      for (int k = 0; k < kEmitGroup; ++k)
        if (jb + k < nj) wv[k] = __ldg(relabel + wv[k]);
    }
    const int64_t j = j0 + jb;
    if (wide_stores && jb + kEmitGroup <= nj) {
      if (which) st_v4_i64(which + j, sv);
      if (within) st_v4_i64(within + j, wv);
      if (flat) st_v4_i64(flat + j, fv);
    } else {
      for (int k = 0; k < kEmitGroup && jb + k < nj; ++k) {
        if (which) which[j + k] = sv[k];
        if (within) within[j + k] = wv[k];
        if (flat) flat[j + k] = fv[k];
      }
    }
  }
}

}  // namespace rua

using namespace rua;

extern "C" {

int rua_mask(const int64_t* len, int64_t B, int64_t W, const void* zero_host, const void* one_host,
             int32_t elem_bytes, void* out, rua_stream_t stream) {
  if (B < 0 || W < 0) return RUA_ERR_INVALID;
  if (B == 0 || W == 0) return RUA_OK;
  if (!len || !zero_host || !one_host || !out) return RUA_ERR_INVALID;
  if (((uintptr_t)out & 15u) != 0) return RUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = B * W;
  const int64_t per = 16 / elem_bytes;
  const int64_t blocks = ceil_div(ceil_div(total, per), 256);
  if (blocks >= (1ll << 31)) return RUA_ERR_UNSUPPORTED;
  switch (elem_bytes) {
    case 1: mask_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(len, B, W, *(const uint8_t*)zero_host, *(const uint8_t*)one_host, (uint8_t*)out); break;
    case 2: mask_kernel<uint16_t><<<(unsigned)blocks, 256, 0, st>>>(len, B, W, *(const uint16_t*)zero_host, *(const uint16_t*)one_host, (uint16_t*)out); break;
    case 4: mask_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(len, B, W, *(const uint32_t*)zero_host, *(const uint32_t*)one_host, (uint32_t*)out); break;
    case 8: mask_kernel<uint64_t><<<(unsigned)blocks, 256, 0, st>>>(len, B, W, *(const uint64_t*)zero_host, *(const uint64_t*)one_host, (uint64_t*)out); break;
    default: return RUA_ERR_INVALID;
  }
  return check_launch();
}

int rua_emit_ptr(const int64_t* off, int64_t S, int64_t n, const int64_t* relabel, int64_t* which,
                 int64_t* within, int64_t* flat, int64_t stride, int32_t right_align, rua_stream_t stream) {
  if (S < 0 || n < 0) return RUA_ERR_INVALID;
  if (n == 0) return RUA_OK;
  if (!off || S == 0) return RUA_ERR_INVALID;
  if (!which && !within && !flat) return RUA_OK;
  const int64_t blocks = ceil_div(n, kEmitTile);
  if (blocks >= (1ll << 31)) return RUA_ERR_UNSUPPORTED;
  const uintptr_t align = (uintptr_t)which | (uintptr_t)within | (uintptr_t)flat;
  emit_ptr_kernel<<<(unsigned)blocks, kEmitThreads, 0, (cudaStream_t)stream>>>(off, S, n, relabel, which, within,
                                                                            flat, stride, right_align,
                                                                            (align & 31u) == 0);
  return check_launch();
}

}  // extern "C"
