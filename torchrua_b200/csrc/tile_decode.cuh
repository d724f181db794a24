// Tile decode for segmented ranges with SHORT segments (narrow rows, index emit -- BASELINE config 5).
//
// A CTA of 256 threads owns 256 * ITEMS consecutive positions [p0, p0 + np) of a range cut into S segments by
// a monotone offset function f (f(0) = 0; segment s = [f(s), f(s+1)); empty segments allowed).  Afterwards
// every thread can resolve ANY position of the tile in O(1):
//     k = sm.seg[p - p0]       segment index relative to `first`  (owner = first + k)
//     sm.rel[k]                tile-relative start of that segment (k == 0: may lie before the tile -> off_first)
// How: two warp-cooperative 32-ary searches find the first / last segment that intersects the tile; their
// starts are staged tile-relative in shared memory; every start inside the tile bumps a counter at its
// position; a block-wide inclusive scan of the counters is exactly "how many segments start at or before
// me".  A per-position binary search (6 dependent shared-memory rounds for 64 segments per tile, in 64-bit)
// is what this replaces: these kernels are issue-bound, not DRAM-bound.
#pragma once

#include "common.cuh"

namespace rua {

constexpr int kDecThreads = 256;

struct TileDecode {
  int64_t first;      // first segment that intersects the tile
  int64_t off_first;  // f(first) (<= p0)
  int cnt;            // staged entries: f(first) .. f(last + 1)
  bool staged;        // false: too many (empty) segments inside the tile -> caller falls back to owner_search
};

template <int ITEMS>
struct alignas(16) TileDecodeSmemT {
  static constexpr int kTile = kDecThreads * ITEMS;
  static constexpr int kCap = kTile + 2;
  int seg[kTile];    // first: accessed as int4
  int64_t bounds[2];
  int rel[kCap];
  int warp_tot[kDecThreads / 32];
};

template <int ITEMS, typename OffFn>
__device__ __forceinline__ TileDecode tile_decode(OffFn f, int64_t S, int64_t p0, int np, TileDecodeSmemT<ITEMS>& sm) {
  static_assert(ITEMS % 4 == 0, "positions per thread are scanned as int4 groups");
  constexpr int kTile = kDecThreads * ITEMS, kCap = kTile + 2, Q = ITEMS / 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    const int64_t a = warp_owner_search(f, S, p0, lane);
    if (lane == 0) sm.bounds[0] = a;
  } else if (warp == 1) {
    const int64_t b = warp_owner_search(f, S, p0 + np - 1, lane);
    if (lane == 0) sm.bounds[1] = b;
  }
  // zero the per-position counters meanwhile
  int4* seg4 = reinterpret_cast<int4*>(sm.seg) + Q * tid;
#pragma unroll
  for (int q = 0; q < Q; ++q) seg4[q] = make_int4(0, 0, 0, 0);
  __syncthreads();
  TileDecode d;
  d.first = sm.bounds[0];
  const int64_t cnt64 = sm.bounds[1] - d.first + 2;
  d.staged = cnt64 <= kCap;
  d.cnt = d.staged ? (int)cnt64 : 0;
  d.off_first = f(d.first);
  if (!d.staged) return d;
  if (d.cnt == 2) {   // ONE segment covers the whole tile (P.ptr: segments are time steps): the zeroed counters are the answer
    if (tid < 2) {
      const int64_t r = f(d.first + tid) - p0;
      sm.rel[tid] = r < -1 ? -1 : (r > kTile + 1 ? kTile + 1 : (int)r);
    }
    __syncthreads();
    return d;
  }
  for (int k = tid; k < d.cnt; k += kDecThreads) {
    const int64_t r = f(d.first + k) - p0;
    sm.rel[k] = r < -1 ? -1 : (r > kTile + 1 ? kTile + 1 : (int)r);
    if (k > 0 && k < d.cnt - 1 && r >= 0 && r < kTile) atomicAdd(&sm.seg[(int)r], 1);   // starts of first+1 .. last lie in [1, np)
  }
  __syncthreads();
  // block-wide inclusive scan of the counters (ITEMS consecutive positions per thread)
  int4 v[Q];
  int run = 0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    v[q] = seg4[q];
    v[q].x += run; v[q].y += v[q].x; v[q].z += v[q].y; v[q].w += v[q].z;
    run = v[q].w;
  }
  int incl = run;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const int o = __shfl_up_sync(kFullMask, incl, s);
    if (lane >= s) incl += o;
  }
  if (lane == 31) sm.warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - run;
#pragma unroll
  for (int w = 0; w < kDecThreads / 32; ++w)
    if (w < warp) base += sm.warp_tot[w];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    v[q].x += base; v[q].y += base; v[q].z += base; v[q].w += base;
    seg4[q] = v[q];
  }
  __syncthreads();
  return d;
}

// the shape the row-map tile kernels and the index emitters use: 2048 positions per CTA
constexpr int kDecItems = 8;
constexpr int kDecTile = kDecThreads * kDecItems;
constexpr int kDecCap = kDecTile + 2;
using TileDecodeSmem = TileDecodeSmemT<kDecItems>;

}  // namespace rua
