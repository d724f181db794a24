// Tile decode for segmented ranges with SHORT segments (narrow rows, index emit -- BASELINE config 5).
//
// A CTA owns kDecTile consecutive positions [p0, p0 + np) of a range cut into S segments by a monotone
// offset function f (f(0) = 0; segment s = [f(s), f(s+1)); empty segments allowed).  Afterwards every
// thread can resolve ANY position of the tile in O(1):
//     k = s_seg[p - p0]        segment index relative to `first`  (owner = first + k)
//     s_rel[k]                 tile-relative start of that segment (k == 0: may lie before the tile -> off_first)
// How: two warp-cooperative 32-ary searches find the first / last segment that intersects the tile; their
// starts are staged tile-relative in shared memory; every start inside the tile bumps a counter at its
// position; a block-wide inclusive scan of the counters is exactly "how many segments start at or before
// me".  A per-position binary search (6 dependent shared-memory rounds for 64 segments per tile, in 64-bit)
// is what this replaces: these kernels are issue-bound, not DRAM-bound.
#pragma once

#include "common.cuh"

namespace rua {

constexpr int kDecThreads = 256;
constexpr int kDecItems = 8;
constexpr int kDecTile = kDecThreads * kDecItems;   // 2048 positions per CTA
constexpr int kDecCap = kDecTile + 2;               // staged segment starts

struct TileDecode {
  int64_t first;      // first segment that intersects the tile
  int64_t off_first;  // f(first) (<= p0)
  int cnt;            // staged entries: f(first) .. f(last + 1)
  bool staged;        // false: too many (empty) segments inside the tile -> caller falls back to owner_search
};

struct alignas(16) TileDecodeSmem {
  int seg[kDecTile];    // first: accessed as int4
  int64_t bounds[2];
  int rel[kDecCap];
  int warp_tot[kDecThreads / 32];
};

template <typename OffFn>
__device__ __forceinline__ TileDecode tile_decode(OffFn f, int64_t S, int64_t p0, int np, TileDecodeSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    const int64_t a = warp_owner_search(f, S, p0, lane);
    if (lane == 0) sm.bounds[0] = a;
  } else if (warp == 1) {
    const int64_t b = warp_owner_search(f, S, p0 + np - 1, lane);
    if (lane == 0) sm.bounds[1] = b;
  }
  // zero the per-position counters meanwhile
  reinterpret_cast<int4*>(sm.seg)[2 * tid] = make_int4(0, 0, 0, 0);
  reinterpret_cast<int4*>(sm.seg)[2 * tid + 1] = make_int4(0, 0, 0, 0);
  __syncthreads();
  TileDecode d;
  d.first = sm.bounds[0];
  const int64_t cnt64 = sm.bounds[1] - d.first + 2;
  d.staged = cnt64 <= kDecCap;
  d.cnt = d.staged ? (int)cnt64 : 0;
  d.off_first = f(d.first);
  if (!d.staged) return d;
  for (int k = tid; k < d.cnt; k += kDecThreads) {
    const int64_t r = f(d.first + k) - p0;
    sm.rel[k] = r < -1 ? -1 : (r > kDecTile + 1 ? kDecTile + 1 : (int)r);
    if (k > 0 && k < d.cnt - 1 && r >= 0 && r < kDecTile) atomicAdd(&sm.seg[(int)r], 1);   // starts of first+1 .. last lie in [1, np)
  }
  __syncthreads();
  // block-wide inclusive scan of the counters (8 consecutive positions per thread)
  int4 a = reinterpret_cast<int4*>(sm.seg)[2 * tid], b = reinterpret_cast<int4*>(sm.seg)[2 * tid + 1];
  a.y += a.x; a.z += a.y; a.w += a.z;
  b.x += a.w; b.y += b.x; b.z += b.y; b.w += b.z;
  int incl = b.w;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const int o = __shfl_up_sync(kFullMask, incl, s);
    if (lane >= s) incl += o;
  }
  if (lane == 31) sm.warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - b.w;
#pragma unroll
  for (int w = 0; w < kDecThreads / 32; ++w)
    if (w < warp) base += sm.warp_tot[w];
  a.x += base; a.y += base; a.z += base; a.w += base;
  b.x += base; b.y += base; b.z += base; b.w += base;
  reinterpret_cast<int4*>(sm.seg)[2 * tid] = a;
  reinterpret_cast<int4*>(sm.seg)[2 * tid + 1] = b;
  __syncthreads();
  return d;
}

}  // namespace rua
