/*
 * rua_oracle.c -- plain C restatement of the reference's ragged hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same closed forms as oracle/rua_oracle.py (which is pinned to the live reference through
 * tests/golden/), written as straight loops so it can (a) check the CUDA path at BASELINE.json's full
 * sizes in seconds and (b) serve as the multi-threaded CPU baseline of bench.py (`cpu_baseline`,
 * `--impl reference`): one OpenMP thread team over sequences / segments, every host core busy.
 * tests/test_oracle_golden.py::test_c_oracle_* pins this file to the numpy oracle and the golden vectors.
 * Nothing under torchrua_b200/ links, loads or calls it.
 *
 * Reference being restated (speedcell4/torchrua v0.5.1):
 *   positions          torchrua/core/get.py:21-79, torchrua/layout/{cat,left,right,pack}.py
 *   pack metadata      torchrua/core/view.py:47-58
 *   conversions        torchrua/core/cast.py:8-71
 *   selects            torchrua/select/{rev,roll,trunc,head,last}.py
 *   masks              torchrua/mask.py:6-12
 *   segment reduce     torchrua/reduce.py:34-61 (ATen segment_reduce: strict left-to-right)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORA_CAT = 0, ORA_LEFT = 1, ORA_PACK = 2, ORA_RIGHT = 3 };
enum { ORA_SUM = 0, ORA_MEAN = 1, ORA_PROD = 2, ORA_MAX = 3, ORA_MIN = 4, ORA_LOGSUMEXP = 5 };
enum { ORA_MAP_ID = 0, ORA_MAP_REV = 1, ORA_MAP_ROLL = 2 };

/* exclusive prefix sum, n+1 entries (torchrua/utils.py:16-19 plus the total) */
void ora_excl_scan(const int64_t* sizes, int64_t n, int64_t* off) {
  int64_t run = 0;
  for (int64_t i = 0; i < n; ++i) { off[i] = run; run += sizes[i]; }
  off[n] = run;
}

/* core/view.py:47-58: batch_sizes[t] = #{i: len[i] > t}; sorted = stable descending argsort (ties by
 * ascending index: the documented deviation from the reference's non-stable sort); unsorted = inverse.
 * If sorted_in != NULL that permutation is used instead (injected-permutation parity mode). */
void ora_pack_meta(const int64_t* len, int64_t B, int64_t T, const int64_t* sorted_in, int64_t* batch_sizes,
                   int64_t* sorted, int64_t* unsorted) {
  int64_t* cnt = (int64_t*)calloc((size_t)T + 2, sizeof(int64_t));
  for (int64_t i = 0; i < B; ++i) cnt[len[i]]++;
  /* start[l] = number of sequences strictly longer than l */
  int64_t run = 0;
  int64_t* start = (int64_t*)malloc(((size_t)T + 2) * sizeof(int64_t));
  for (int64_t l = T; l >= 0; --l) { start[l] = run; run += cnt[l]; }
  for (int64_t t = 0; t < T; ++t) batch_sizes[t] = start[t];
  if (sorted_in) {
    for (int64_t r = 0; r < B; ++r) sorted[r] = sorted_in[r];
  } else {
    for (int64_t i = 0; i < B; ++i) sorted[start[len[i]]++] = i;
  }
  for (int64_t r = 0; r < B; ++r) unsorted[sorted[r]] = r;
  free(cnt);
  free(start);
}

/* core/view.py:21-25 on a P source: len[i] = #{t: bs[t] > unsorted[i]} */
void ora_lengths_from_pack(const int64_t* bs, const int64_t* unsorted, int64_t B, int64_t T, int64_t* len) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i) {
    int64_t lo = 0, hi = T, r = unsorted[i];
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (bs[mid] > r) lo = mid + 1; else hi = mid; }
    len[i] = lo;
  }
}

typedef struct {
  int kind;
  int64_t width;           /* L/R: padded width */
  const int64_t* off;      /* C */
  const int64_t* poff;     /* P */
  const int64_t* unsorted; /* P */
} ora_side;

static inline int64_t ora_row(const ora_side* s, int64_t i, int64_t t, int64_t len_i) {
  switch (s->kind) {
    case ORA_CAT: return s->off[i] + t;                          /* get.py:25-26 */
    case ORA_LEFT: return i * s->width + t;                      /* get.py:41-42 */
    case ORA_RIGHT: return i * s->width + (s->width - len_i) + t; /* get.py:73-74 */
    default: return s->poff[t] + s->unsorted[i];                 /* get.py:57-58 */
  }
}

/*
 * General ragged move: for every token (i, t) of the destination (lengths len[i]),
 *   dst[row_dst(i,t)] = src[row_src(i, map(t))],   map in {identity, rev, roll(shift)};
 * padded destinations are first filled with `fill` (cast.py:19-23: new_full then index_put), or with
 * flat source row 0 when pad_row0 != 0 (L/R.roll quirk, select/roll.py:19-34).
 * bs / unsorted describe the P side(s); src_width / dst_width the L/R side(s).
 */
void ora_move(const uint8_t* src, uint8_t* dst, int64_t row_bytes, int src_kind, int dst_kind, const int64_t* len,
              int64_t B, const int64_t* bs, int64_t T, const int64_t* unsorted, int64_t src_width,
              int64_t dst_width, int map, int64_t shift, const uint8_t* fill, int fill_bytes, int pad_row0) {
  int64_t* off = (int64_t*)malloc(((size_t)B + 1) * sizeof(int64_t));
  int64_t* poff = NULL;
  ora_excl_scan(len, B, off);
  if (src_kind == ORA_PACK || dst_kind == ORA_PACK) {
    poff = (int64_t*)malloc(((size_t)T + 1) * sizeof(int64_t));
    ora_excl_scan(bs, T, poff);
  }
  ora_side s = {src_kind, src_width, off, poff, unsorted};
  ora_side d = {dst_kind, dst_width, off, poff, unsorted};
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < B; ++i) {
    const int64_t n = len[i];
    if (dst_kind == ORA_LEFT || dst_kind == ORA_RIGHT) {
      const int64_t p0 = dst_kind == ORA_LEFT ? n : 0, p1 = dst_kind == ORA_LEFT ? dst_width : dst_width - n;
      for (int64_t p = p0; p < p1; ++p) {
        uint8_t* q = dst + (i * dst_width + p) * row_bytes;
        if (pad_row0) memcpy(q, src, (size_t)row_bytes);
        else for (int64_t k = 0; k < row_bytes; k += fill_bytes) memcpy(q + k, fill, (size_t)fill_bytes);
      }
    }
    for (int64_t t = 0; t < n; ++t) {
      int64_t ts = t;
      if (map == ORA_MAP_REV) ts = n - 1 - t;                                   /* rev.py */
      else if (map == ORA_MAP_ROLL) { ts = (t - shift) % n; if (ts < 0) ts += n; } /* roll.py:11 */
      memcpy(dst + ora_row(&d, i, t, n) * row_bytes, src + ora_row(&s, i, ts, n) * row_bytes, (size_t)row_bytes);
    }
  }
  free(off);
  free(poff);
}

/* mask.py:6-12: (B, W) elements of elem_bytes; `one` where t < len[i] else `zero` (left aligned) */
void ora_mask(const int64_t* len, int64_t B, int64_t W, const uint8_t* zero, const uint8_t* one, int elem_bytes,
              uint8_t* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
    for (int64_t t = 0; t < W; ++t)
      memcpy(out + (i * W + t) * elem_bytes, t < len[i] ? one : zero, (size_t)elem_bytes);
}

static inline float bf16_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static inline uint16_t f32_to_bf16(float f) { /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

/*
 * reduce.py:34-61 over (N, H) fp32 rows (bf16 rows when is_bf16: upcast, accumulate in fp32, round
 * once -- the bf16 parity contract of SURVEY.md 8c hazard 2).  Strict left-to-right accumulation from
 * `initial`: 0 (sum, mean), 1 (prod), the GLOBAL min / max of the whole tensor (max, logsumexp / min).
 */
void ora_segment_reduce(const void* data, const int64_t* sizes, int64_t S, int64_t H, int op, int is_bf16,
                        void* out) {
  int64_t* off = (int64_t*)malloc(((size_t)S + 1) * sizeof(int64_t));
  ora_excl_scan(sizes, S, off);
  const int64_t total = off[S] * H;
  const float* xf = (const float*)data;
  const uint16_t* xb = (const uint16_t*)data;
#define LOAD(idx) (is_bf16 ? bf16_to_f32(xb[idx]) : xf[idx])
  float gmin = INFINITY, gmax = -INFINITY;
  int has_nan = 0;
  if (op == ORA_MAX || op == ORA_MIN || op == ORA_LOGSUMEXP) {
#pragma omp parallel for reduction(min : gmin) reduction(max : gmax) reduction(| : has_nan) schedule(static)
    for (int64_t k = 0; k < total; ++k) {
      float v = LOAD(k);
      if (v != v) has_nan = 1;
      if (v < gmin) gmin = v;
      if (v > gmax) gmax = v;
    }
    if (has_nan) gmin = gmax = NAN; /* tensor.min() / .max() return NaN if any element is NaN */
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t s = 0; s < S; ++s) {
    const int64_t n = sizes[s], r0 = off[s];
    for (int64_t h = 0; h < H; ++h) {
      float acc, m = 0.f;
      switch (op) {
        case ORA_PROD: acc = 1.f; break;
        case ORA_MAX: case ORA_LOGSUMEXP: acc = gmin; break;
        case ORA_MIN: acc = gmax; break;
        default: acc = 0.f;
      }
      if (op == ORA_LOGSUMEXP) {
        for (int64_t r = 0; r < n; ++r) { float v = LOAD((r0 + r) * H + h); acc = (v != v) ? v : (acc < v ? v : acc); }
        m = acc;
        float sum = 0.f;
        for (int64_t r = 0; r < n; ++r) sum += expf(LOAD((r0 + r) * H + h) - m);
        acc = logf(sum + (n == 0 ? 1.f : 0.f)) + m;
      } else {
        for (int64_t r = 0; r < n; ++r) {
          float v = LOAD((r0 + r) * H + h);
          switch (op) {
            case ORA_PROD: acc = acc * v; break;
            case ORA_MAX: acc = (v != v) ? v : (acc < v ? v : acc); break;
            case ORA_MIN: acc = (v != v) ? v : (v < acc ? v : acc); break;
            default: acc = acc + v;
          }
        }
        if (op == ORA_MEAN && n > 0 && acc == acc) acc = acc / (float)n;
      }
      if (is_bf16) ((uint16_t*)out)[s * H + h] = f32_to_bf16(acc);
      else ((float*)out)[s * H + h] = acc;
    }
  }
#undef LOAD
  free(off);
}

void ora_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int ora_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}
