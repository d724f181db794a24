// K4-flat -- segment reduce for FEATURELESS data (H == 1: per-token scalars such as log-probabilities
// or token ids), where the wide-row kernel of reduce.cu would leave 31 of 32 lanes idle.
//
// Rows map to lanes.  A CTA owns a tile of 256*G consecutive rows (G = 16 / sizeof(T): one 128-bit load
// per thread).  The segments that intersect the tile are found with two warp-cooperative 32-ary searches
// and their offsets staged in shared memory.  Each thread reduces its G rows locally, cutting at
// segment boundaries: runs that start AND end inside the thread are complete segments and are stored
// directly; the first run may continue one from the threads before, the last may continue into the
// threads after.  Those open ends are stitched by a block-wide SEGMENTED SCAN BY KEY over
// (segment id, partial) pairs -- warp shuffles, then one shared-memory hop across the 8 warps.  Pieces
// of segments that cross tile boundaries go to the same head/tail scratch as in reduce.cu and are
// merged in tile order by segreduce_span_kernel, so results stay deterministic.
#pragma once

#include "reduce_common.cuh"

namespace rua {

constexpr int kFlatThreads = 256;
constexpr int kFlatWarps = kFlatThreads / 32;

template <typename A, int OP, bool kFast>
__device__ __forceinline__ void flat_merge(State<A, 1, OP>& later, const State<A, 1, OP>& earlier) {
  State<A, 1, OP> t = earlier;
  t.template merge<kFast>(later);
  later = t;
}

template <typename A, int OP>
__device__ __forceinline__ State<A, 1, OP> flat_shfl_up(const State<A, 1, OP>& v, int d) {
  State<A, 1, OP> o;
  o.a[0] = __shfl_up_sync(kFullMask, v.a[0], d);
  o.s[0] = A(0);
  if constexpr (OpInfo<OP>::kIsLse) o.s[0] = __shfl_up_sync(kFullMask, v.s[0], d);
  return o;
}

template <typename T, int OP>
__global__ void __launch_bounds__(kFlatThreads)
segreduce_flat_kernel(const T* __restrict__ data, const int64_t* __restrict__ off, int64_t N, int64_t S,
                      T* __restrict__ out, typename Store<T>::Acc* __restrict__ head,
                      typename Store<T>::Acc* __restrict__ tail, int64_t* __restrict__ tail_seg, RedHeader* hdr,
                      int vector_loads) {
  using A = typename Store<T>::Acc;
  using St = State<A, 1, OP>;
  constexpr bool kFast = sizeof(T) == 2;
  constexpr int G = 16 / sizeof(T);
  constexpr int R = kFlatThreads * G;          // rows per tile
  constexpr int kCap = R + 2;
  constexpr int P = OpInfo<OP>::kParts;
  __shared__ int64_t s_off[kCap];
  __shared__ int64_t s_bounds[2];
  __shared__ long long s_wkey[kFlatWarps];
  __shared__ A s_wval[kFlatWarps][2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  const int64_t row0 = tile * R;
  const int64_t row1 = row0 + R < N ? row0 + R : N;

  // ---- which segments intersect the tile ------------------------------------------------------
  GlobalOff g{off};
  if (warp == 0) {
    int64_t a = warp_owner_search(g, S, row0, lane);
    if (lane == 0) s_bounds[0] = a;
  } else if (warp == 1) {
    int64_t b = warp_owner_search(g, S, row1 - 1, lane);
    if (lane == 0) s_bounds[1] = b;
  }
  __syncthreads();
  const int64_t first = s_bounds[0], last = s_bounds[1];
  const int64_t cnt = last - first + 2;
  const bool staged = cnt <= kCap;
  if (staged)
    for (int64_t k = tid; k < cnt; k += kFlatThreads) s_off[k] = __ldg(off + first + k);
  __syncthreads();
  auto off_at = [&](int64_t k) -> int64_t { return staged ? s_off[k] : __ldg(off + first + k); };

  // ---- thread-local reduction of G consecutive rows, cut at segment boundaries ------------------
  const int64_t tr0 = row0 + (int64_t)tid * G;
  A x[G];
  if (tr0 + G <= row1 && vector_loads) {
    Raw<T, G> raw;
    raw.r = __ldcs(reinterpret_cast<const uint4*>(data + tr0));
    Store<T>::unpack(raw.r, x);
  } else {
#pragma unroll
    for (int k = 0; k < G; ++k) x[k] = tr0 + k < row1 ? Store<T>::to_acc(data[tr0 + k]) : A(0);
  }

  A ext = OP == RUA_MIN ? -inf_of<A>() : inf_of<A>();
  bool saw_nan = false;
  St firstRun, acc;
  acc.reset();
  firstRun.reset();
  int64_t firstSeg = -1, firstLen = 0;   // the first CLOSED run of this thread (may need a carry-in)
  int closed = 0;
  int64_t lo = 0, seg_beg = 0, seg_end = 0;
  bool have = false;                     // acc holds at least one row
  if (tr0 < row1) {
    // segment of the first row: binary search over the staged offsets
    int64_t a = 0, b = cnt - 1;
    while (b - a > 1) {
      const int64_t mid = (a + b) >> 1;
      if (off_at(mid) <= tr0) a = mid; else b = mid;
    }
    lo = a;
    seg_beg = off_at(lo);
    seg_end = off_at(lo + 1);
  }
  auto close_run = [&]() {
    if (closed == 0) {
      firstRun = acc;
      firstSeg = first + lo;
      firstLen = seg_end - seg_beg;
    } else {  // started and ended inside this thread: a complete segment
      A o[1];
      acc.finalize(seg_end - seg_beg, o);
      if (OpInfo<OP>::kNeedsExt) saw_nan |= (o[0] != o[0]);
      out[first + lo] = Store<T>::from_acc(o[0]);
    }
    ++closed;
    acc.reset();
    have = false;
  };
#pragma unroll
  for (int k = 0; k < G; ++k) {
    const int64_t row = tr0 + k;
    if (row < row1) {
      if (row >= seg_end) {
        if (have) close_run();
        while (row >= seg_end && lo + 2 < cnt) {  // next non-empty segment
          ++lo;
          seg_beg = seg_end;
          seg_end = off_at(lo + 1);
        }
        if (row >= seg_end) break;  // rows past the last segment (sum of sizes < N): not reduced
      }
      acc.template add<kFast>(&x[k]);
      have = true;
      if (OpInfo<OP>::kNeedsExt) ext = OP == RUA_MIN ? max_num(ext, x[k]) : min_num(ext, x[k]);
    }
  }
  const int64_t my_end = tr0 + G < row1 ? tr0 + G : row1;  // one past this thread's last row
  if (have && seg_end == my_end) close_run();              // the segment ends exactly with this thread
  // what is left in `acc` (if `have`) is an OPEN run of segment first+lo, continuing to the right
  long long key = have ? (long long)(first + lo) : -1ll;
  St val = acc;

  // ---- block-wide segmented inclusive scan by key -----------------------------------------------
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long k2 = __shfl_up_sync(kFullMask, key, d);
    St v2 = flat_shfl_up<A, OP>(val, d);
    if (lane >= d && k2 == key && key >= 0) flat_merge<A, OP, kFast>(val, v2);
  }
  if (lane == 31) {
    s_wkey[warp] = key;
    s_wval[warp][0] = val.a[0];
    s_wval[warp][1] = OpInfo<OP>::kIsLse ? val.s[0] : A(0);
  }
  __syncthreads();
  long long ckey = -1;   // carry from the warps before this one
  St cval;
  cval.reset();
  for (int w = 0; w < warp; ++w) {
    St wv;
    wv.a[0] = s_wval[w][0];
    wv.s[0] = s_wval[w][1];
    const long long wk = s_wkey[w];
    if (wk >= 0 && wk == ckey) flat_merge<A, OP, kFast>(wv, cval);
    ckey = wk;
    cval = wv;
  }
  // equal keys are contiguous, so every lane whose key equals the carry's key is connected to the warp start
  if (key >= 0 && key == ckey) flat_merge<A, OP, kFast>(val, cval);
  // exclusive value = block-wide inclusive value of the previous thread
  long long ekey = __shfl_up_sync(kFullMask, key, 1);
  St eval = flat_shfl_up<A, OP>(val, 1);
  if (lane == 0) {
    ekey = ckey;
    eval = cval;
  }

  // ---- emit -----------------------------------------------------------------------------------
  if (closed > 0) {  // first closed run: fold in what the threads before contributed to the same segment
    St v = firstRun;
    if (ekey == firstSeg) flat_merge<A, OP, kFast>(v, eval);
    const int64_t beg = __ldg(off + firstSeg);
    if (beg >= row0) {
      A o[1];
      v.finalize(firstLen, o);
      if (OpInfo<OP>::kNeedsExt) saw_nan |= (o[0] != o[0]);
      out[firstSeg] = Store<T>::from_acc(o[0]);
    } else {  // the segment began in an earlier tile: this is the tile's head piece
      store_partial<A, 1, OP>(head + tile * P, 1, 0, v);
    }
  }
  const int last_thread = (int)((row1 - 1 - row0) / G);
  if (tid == last_thread) {
    int64_t spans = -1;
    if (key >= 0) {  // the tile ends inside a segment
      const int64_t beg = __ldg(off + key);
      store_partial<A, 1, OP>((beg >= row0 ? tail : head) + tile * P, 1, 0, val);
      if (beg >= row0) spans = key;
    }
    tail_seg[tile] = spans;  // read by segreduce_span_kernel: the segment that starts here and runs on
  }

  if (OpInfo<OP>::kNeedsExt) {
    __shared__ unsigned long long s_key[kFlatWarps];
    __shared__ int s_nan;
    if (tid == 0) s_nan = 0;
    __syncthreads();
    unsigned long long okey = order_key(ext);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      unsigned long long o = __shfl_xor_sync(kFullMask, okey, d);
      okey = OP == RUA_MIN ? (o > okey ? o : okey) : (o < okey ? o : okey);
    }
    if (lane == 0) s_key[warp] = okey;
    if (saw_nan) s_nan = 1;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kFlatWarps; ++w) {
        unsigned long long o = s_key[w];
        okey = OP == RUA_MIN ? (o > okey ? o : okey) : (o < okey ? o : okey);
      }
      if (OP == RUA_MIN) atomicMax(&hdr->ext_key, okey); else atomicMin(&hdr->ext_key, okey);
      if (s_nan) atomicOr(&hdr->nan_flag, 1u);
    }
  }
}

}  // namespace rua
