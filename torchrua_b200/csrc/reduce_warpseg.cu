// K4-ws -- segment reduce for NARROW rows (H * elem <= 16 bytes) and MANY SHORT segments: per-token scalars reduced
// per sequence (BASELINE config 5: 1 M segments of 1..64 fp32 values).
//
// Replaces (reference file:line): the same segment_* functions as reduce.cu (torchrua/reduce.py:34-61) on the
// featureless / few-column inputs that mask.py / segment.py style pooling produces.
//
// The rows-on-lanes kernel of reduce_flat.cu spends its time in a per-tile decode chain (two searches, a staged
// offset table, a block scan, three barriers: 49 M warp instructions for 32.5 M rows, 16 % of the HBM peak).  Here a
// WARP owns 32 consecutive SEGMENTS instead:
//   * lane i reads off[s0 + i], off[s0 + i + 1] -- two coalesced loads, no search, no decode, no block barrier;
//   * the rows of those 32 segments are one contiguous range; the warp streams it through a private 2 KB
//     shared-memory window with coalesced 128-bit loads (the next window is requested before the current one is reduced);
//   * every lane then reduces ITS OWN segment out of the window: conflict-tolerant LDS + one accumulate per row,
//     all 32 lanes busy -- about 10x fewer instructions than the tile kernel;
//   * a window that lies entirely inside ONE long segment is reduced by the whole warp from registers (shuffle tree), so
//     a long sequence among short ones costs bandwidth, not a serial loop;
//   * every segment is finished by the lane that owns it: no cross-tile pieces, no span kernel, results stored coalesced.
// The launcher picks this kernel when there are enough segments to fill the machine (reduce.cu: plan_reduce).
#include "reduce_common.cuh"

namespace rua {

constexpr int kWsThreads = 256;
constexpr int kWsWarps = kWsThreads / 32;
constexpr int kWsChunkBytes = 2048;
constexpr int kWsVecPerLane = kWsChunkBytes / 16 / 32;   // 4

template <typename T, int HE>
__device__ __forceinline__ void ws_load_row(const T* p, typename Store<T>::Acc* x) {
  constexpr int kBytes = HE * (int)sizeof(T);
  if constexpr (kBytes == 16) {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    Store<T>::unpack(w, x);
  } else if constexpr (kBytes == 8) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    typename Store<T>::Acc y[16 / sizeof(T)];
    Store<T>::unpack(make_uint4(w.x, w.y, 0u, 0u), y);
#pragma unroll
    for (int h = 0; h < HE; ++h) x[h] = y[h];
  } else if constexpr (kBytes == 4) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    typename Store<T>::Acc y[16 / sizeof(T)];
    Store<T>::unpack(make_uint4(w, 0u, 0u, 0u), y);
#pragma unroll
    for (int h = 0; h < HE; ++h) x[h] = y[h];
  } else {
#pragma unroll
    for (int h = 0; h < HE; ++h) x[h] = Store<T>::to_acc(p[h]);
  }
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

template <typename T, int HE, int OP>
__global__ void __launch_bounds__(kWsThreads)
segreduce_warpseg_kernel(const T* __restrict__ data, const int64_t* __restrict__ off, int64_t N, int64_t S,
                         T* __restrict__ out, RedHeader* hdr) {
  using A = typename Store<T>::Acc;
  using St = State<A, HE, OP>;
  constexpr bool kFast = sizeof(T) == 2;
  constexpr int E = 16 / (int)sizeof(T);            // elements per 16-byte vector
  constexpr int G = E / HE;                         // rows per vector
  constexpr int CE = kWsChunkBytes / (int)sizeof(T);
  constexpr int CR = CE / HE;                       // rows per window
  __shared__ uint4 s_buf[kWsWarps][kWsChunkBytes / 16];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = ((int64_t)blockIdx.x * kWsWarps + warp) * 32;
  St acc;
  acc.reset();
  A ext = OP == RUA_MIN ? -inf_of<A>() : inf_of<A>();
  bool saw_nan = false;

  if (s0 < S) {                                     // warp-uniform
    const int64_t covered = __ldg(off + S) < N ? __ldg(off + S) : N;   // rows past the last segment / past N belong to nobody
    const int64_t s = s0 + lane;
    int64_t beg = s < S ? __ldg(off + s) : covered, end = s < S ? __ldg(off + s + 1) : covered;
    const int64_t len = end - beg;
    beg = beg < covered ? beg : covered;
    end = end < covered ? end : covered;
    const int64_t wbeg = shfl_i64(beg, 0), wend = shfl_i64(end, 31);
    const int64_t total_e = N * HE;
    const uint4* s_mine = s_buf[warp];
    uint4 raw[kWsVecPerLane];

    // request the vectors of the window that starts at element e0 (a multiple of E: data is 16-byte aligned)
    auto request = [&](int64_t e0) {
#pragma unroll
      for (int j = 0; j < kWsVecPerLane; ++j) {
        const int64_t e = e0 + (int64_t)(lane + 32 * j) * E;
        raw[j] = make_uint4(0u, 0u, 0u, 0u);
        if (e < wend * HE) {
          if (e + E <= total_e) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(raw[j].x), "=r"(raw[j].y), "=r"(raw[j].z), "=r"(raw[j].w) : "l"(data + e));
          } else {                                   // the last, partial vector of the array
            T* tmp = reinterpret_cast<T*>(&raw[j]);
#pragma unroll
            for (int k = 0; k < E; ++k)
              if (e + k < total_e) tmp[k] = data[e + k];
          }
        }
      }
    };

    const int64_t e_first = (wbeg * HE) & ~(int64_t)(E - 1);
    if (wbeg < wend) request(e_first);
    for (int64_t e0 = e_first; e0 < wend * HE; e0 += CE) {
      const int64_t r_lo = e0 / HE;                  // first row of the window (e0 is a multiple of E, HE divides E)
      const int64_t r_hi = r_lo + CR < wend ? r_lo + CR : wend;
      uint4* s_w = s_buf[warp];
#pragma unroll
      for (int j = 0; j < kWsVecPerLane; ++j) s_w[lane + 32 * j] = raw[j];
      // does ONE segment cover the whole window?  (at most one lane can say yes)
      const bool covers = beg <= r_lo && end >= r_lo + CR;
      const unsigned who = __ballot_sync(kFullMask, covers);
      if (who) {
        St part;
        part.reset();
#pragma unroll
        for (int j = 0; j < kWsVecPerLane; ++j) {
          A x[E];
          Store<T>::unpack(raw[j], x);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            part.template add<kFast>(&x[g * HE]);
            if (OpInfo<OP>::kNeedsExt) {
#pragma unroll
              for (int h = 0; h < HE; ++h) ext = OP == RUA_MIN ? max_num(ext, x[g * HE + h]) : min_num(ext, x[g * HE + h]);
            }
          }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const St other = ws_shfl_xor<A, HE, OP>(part, d);
          part.template merge<kFast>(other);
        }
        if (lane == __ffs(who) - 1) acc.template merge<kFast>(part);
      }
      if (e0 + CE < wend * HE) request(e0 + CE);     // the next window is in flight while this one is reduced
      __syncwarp();
      if (!who) {
        const int64_t lo = beg > r_lo ? beg : r_lo, hi = end < r_hi ? end : r_hi;
        const T* sb = reinterpret_cast<const T*>(s_mine) + (lo - r_lo) * HE;
        const int n = hi > lo ? (int)(hi - lo) : 0;
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
          A x[HE];
          ws_load_row<T, HE>(sb + r * HE, x);
          acc.template add<kFast>(x);
          if (OpInfo<OP>::kNeedsExt) {
#pragma unroll
            for (int h = 0; h < HE; ++h) ext = OP == RUA_MIN ? max_num(ext, x[h]) : min_num(ext, x[h]);
          }
        }
      }
      __syncwarp();
    }

    if (s < S) {
      A o[HE];
      if (len > 0) {
        acc.finalize(OP == RUA_MEAN ? len : 1, o);
        if (OpInfo<OP>::kNeedsExt) saw_nan |= acc.any_nan_out(o);
      } else {                                       // empty: sum / mean -> 0, prod -> 1; max / min / logsumexp are patched
#pragma unroll
        for (int h = 0; h < HE; ++h) o[h] = OP == RUA_PROD ? A(1) : A(0);
      }
      T packed[HE];
#pragma unroll
      for (int h = 0; h < HE; ++h) packed[h] = Store<T>::from_acc(o[h]);
      if constexpr (HE * sizeof(T) == 16) *reinterpret_cast<uint4*>(out + s * HE) = *reinterpret_cast<const uint4*>(packed);
      else if constexpr (HE * sizeof(T) == 8) *reinterpret_cast<uint2*>(out + s * HE) = *reinterpret_cast<const uint2*>(packed);
      else if constexpr (HE * sizeof(T) == 4) *reinterpret_cast<uint32_t*>(out + s * HE) = *reinterpret_cast<const uint32_t*>(packed);
      else {
#pragma unroll
        for (int h = 0; h < HE; ++h) out[s * HE + h] = packed[h];
      }
    }
  }

  if (OpInfo<OP>::kNeedsExt) {                       // global extreme + NaN flag: one atomic per CTA, and only if it can matter
    __shared__ unsigned long long s_key[kWsWarps];
    __shared__ int s_nan;
    if (threadIdx.x == 0) s_nan = 0;
    __syncthreads();
    unsigned long long key = order_key(ext);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(kFullMask, key, d);
      key = OP == RUA_MIN ? (o > key ? o : key) : (o < key ? o : key);
    }
    if (lane == 0) s_key[warp] = key;
    if (saw_nan) s_nan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kWsWarps; ++w) {
        const unsigned long long o = s_key[w];
        key = OP == RUA_MIN ? (o > key ? o : key) : (o < key ? o : key);
      }
      const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&hdr->ext_key);
      if (OP == RUA_MIN ? key > cur : key < cur) { if (OP == RUA_MIN) atomicMax(&hdr->ext_key, key); else atomicMin(&hdr->ext_key, key); }
      if (s_nan) atomicOr(&hdr->nan_flag, 1u);
    }
  }
}

template <typename T, int HE>
static int ws_launch2(int op, const void* data, const int64_t* off, int64_t N, int64_t S, void* out, RedHeader* hdr,
                      cudaStream_t st) {
  const int64_t blocks = ceil_div(S, (int64_t)kWsWarps * 32);
  if (blocks >= (1ll << 31)) return RUA_ERR_UNSUPPORTED;
  const unsigned nb = (unsigned)blocks;
#define RUA_WS(OP_) segreduce_warpseg_kernel<T, HE, OP_><<<nb, kWsThreads, 0, st>>>((const T*)data, off, N, S, (T*)out, hdr)
  switch (op) {
    case RUA_SUM: RUA_WS(RUA_SUM); break;
    case RUA_MEAN: RUA_WS(RUA_MEAN); break;
    case RUA_PROD: RUA_WS(RUA_PROD); break;
    case RUA_MAX: RUA_WS(RUA_MAX); break;
    case RUA_MIN: RUA_WS(RUA_MIN); break;
    case RUA_LOGSUMEXP: RUA_WS(RUA_LOGSUMEXP); break;
    default: return RUA_ERR_INVALID;
  }
#undef RUA_WS
  return check_launch();
}

template <typename T>
static int ws_launch1(int he, int op, const void* data, const int64_t* off, int64_t N, int64_t S, void* out, RedHeader* hdr,
                      cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  if (he == 1) return ws_launch2<T, 1>(op, data, off, N, S, out, hdr, st);
  if constexpr (E >= 2) if (he == 2) return ws_launch2<T, 2>(op, data, off, N, S, out, hdr, st);
  if constexpr (E >= 4) if (he == 4) return ws_launch2<T, 4>(op, data, off, N, S, out, hdr, st);
  if constexpr (E >= 8) if (he == 8) return ws_launch2<T, 8>(op, data, off, N, S, out, hdr, st);
  return RUA_ERR_UNSUPPORTED;
}

// enough short segments to fill the machine with warps of 32 segments each?  (host-side choice, reduce.cu)
bool warpseg_applies(int64_t N, int64_t S) {
  static const int64_t min_s = [] { const char* e = getenv("RUA_WARPSEG_MIN_S"); return e ? atoll(e) : 32768ll; }();
  return S >= min_s && N <= 256 * S;
}

int warpseg_launch(int32_t dtype, int64_t H, int32_t op, const void* data, const int64_t* off, int64_t N, int64_t S,
                   void* out, void* hdr, cudaStream_t st) {
  RedHeader* h = (RedHeader*)hdr;
  switch (dtype) {
    case RUA_F32: return ws_launch1<float>((int)H, op, data, off, N, S, out, h, st);
    case RUA_F64: return ws_launch1<double>((int)H, op, data, off, N, S, out, h, st);
    case RUA_F16: return ws_launch1<__half>((int)H, op, data, off, N, S, out, h, st);
    default: return ws_launch1<__nv_bfloat16>((int)H, op, data, off, N, S, out, h, st);
  }
}

}  // namespace rua
