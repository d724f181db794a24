/*
 * rua_b200.h -- C ABI of the B200-native (sm_100a) ragged-sequence kernels.
 *
 * The reference (speedcell4/torchrua v0.5.1) is pure Python over stock ATen ops and has NO native
 * plugin / FFI interface of its own.  This header therefore *defines* the boundary a
 * maintainer would bind: one entry point per kernel family, each citing the reference functions
 * whose bodies it replaces.  The Python mirror of the reference API (package torchrua_b200) calls
 * these through ctypes; INTEGRATION.md shows the stub for binding them from the reference itself.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns and allocates every buffer, including workspaces (query the *_workspace_bytes
 *     functions); kernels are stateless, enqueue on `stream` and never synchronise;
 *   - return value: 0 on success, <0 on error (rua_error_string); nothing throws across the ABI;
 *   - int64 metadata everywhere, exactly like the reference (torchrua/layout/cat.py:75,
 *     torchrua/core/view.py:15).
 *
 * Notation: B sequences with base lengths len[i]; off = exclusive prefix sum of len (B+1 entries);
 * T = max len; batch_sizes bs[t] = #{i: len[i] > t}; poff = exclusive prefix sum of bs (T+1
 * entries); sorted = argsort-descending permutation, unsorted = its inverse.
 * Row of token (i,t):  C: off[i]+t   L: i*W+t   R: i*W+(W-len[i])+t   P: poff[t]+unsorted[i]
 * (torchrua/core/get.py:21-79, torchrua/layout/{cat,left,right,pack}.py).
 */
#ifndef RUA_B200_H_
#define RUA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rua_stream_t; /* a cudaStream_t */

enum rua_status {
  RUA_OK = 0,
  RUA_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, bad enum, misalignment) */
  RUA_ERR_WORKSPACE = -2,   /* workspace too small */
  RUA_ERR_UNSUPPORTED = -3, /* size / dtype outside what the kernels implement */
  RUA_ERR_CUDA = -4         /* a CUDA runtime call failed; see rua_last_cuda_error() */
};

enum rua_layout { RUA_CAT = 0, RUA_LEFT = 1, RUA_PACK = 2, RUA_RIGHT = 3 };

/* how one side's per-sequence lengths derive from the base lengths */
enum rua_len_xform {
  RUA_LEN_SAME = 0,  /* len'[i] = len[i]                                   (conversions, rev, roll) */
  RUA_LEN_CONST = 1, /* len'[i] = arg                                      (head(n), last)          */
  RUA_LEN_MINUS = 2  /* len'[i] = len[i] - arg                             (trunc((a,b)), arg=a+b)  */
};

/* token map dst t_d -> src t_s (validity: 0 <= t_s < len_src[i]) */
enum rua_tmap {
  RUA_MAP_SHIFT = 0, /* t_s = t_d + arg                 identity / trunc / head and their inverses */
  RUA_MAP_REV = 1,   /* t_s = len[i] - 1 - t_d          rev, last (torchrua/select/rev.py, last.py) */
  RUA_MAP_ROLL = 2   /* t_s = (t_d - arg) mod len[i]    roll (torchrua/select/roll.py:11)           */
};

enum rua_pad {
  RUA_PAD_FILL = 0, /* padding / unmapped rows receive the fill pattern */
  RUA_PAD_ROW0 = 1, /* ... receive a copy of flat source row 0 (L/R.roll quirk, select/roll.py:19-34) */
  RUA_PAD_WRAP = 2  /* FILL, except that source position -1 wraps like a negative Python index: what
                       last() / segment_last return for an EMPTY sequence (select/last.py:11-13)   */
};

enum rua_dtype { RUA_F32 = 0, RUA_F64 = 1, RUA_F16 = 2, RUA_BF16 = 3 };

enum rua_reduce {
  RUA_SUM = 0, RUA_MEAN = 1, RUA_PROD = 2, RUA_MAX = 3, RUA_MIN = 4, RUA_LOGSUMEXP = 5
};

/* base ragged structure shared by both sides of a row map */
typedef struct {
  int64_t B;               /* number of sequences                                             */
  const int64_t* off;      /* B+1: exclusive prefix sum of the base lengths                   */
  const int64_t* poff;     /* Tp+1: exclusive prefix sum of batch_sizes; NULL if no P side    */
  const int64_t* sorted;   /* B: sorted_indices; NULL if no P side                            */
  const int64_t* unsorted; /* B: unsorted_indices; NULL if no P side                          */
  int64_t Tp;              /* number of time steps in poff (len(batch_sizes)); 0 if no P side */
} rua_ragged_t;

/* one side (source or destination) of a row map */
typedef struct {
  int32_t layout;    /* enum rua_layout                                            */
  int32_t len_xform; /* enum rua_len_xform                                         */
  int64_t len_arg;   /* n for CONST, a+b for MINUS                                 */
  int64_t width;     /* L/R: rows per sequence in storage (data.size(1)); else 0   */
  int64_t rows;      /* total rows in this side's flattened storage (N' or B*W)    */
} rua_side_t;

/* ------------------------------------------------------------------------------------------- */
/* library                                                                                       */
/* ------------------------------------------------------------------------------------------- */
int rua_version(void);
const char* rua_error_string(int status);
int rua_last_cuda_error(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t rua_launch_count(void);
/* host-only self test of launch-time arithmetic (magic-number division of the narrow-row kernels): 0 = ok.  No GPU. */
int rua_selftest(void);

/* ------------------------------------------------------------------------------------------- */
/* K0  metadata from lengths                                                                     */
/* replaces get_offsets (torchrua/utils.py:16-19), size() (layout/cat.py:61-66),                 */
/* invert_permutation (utils.py:22-26), pack_view (core/view.py:47-58), and the                  */
/* get_mask(...).sum(dim) detours of cat_view/left_view/right_view (core/view.py:11-38,67-71)    */
/* ------------------------------------------------------------------------------------------- */

/* off[0..n] = min(exclusive prefix sum of sizes[0..n), clamp_max); stats[0] = sum, stats[1] =
 * max(sizes) (0 if n==0).  clamp_max = INT64_MAX for the plain scan; C.offsets()/P.offsets() pass N-1
 * (torchrua/layout/cat.py:81, pack.py:45).  ws: rua_scan_workspace_bytes(n) bytes. */
size_t rua_scan_workspace_bytes(int64_t n);
int rua_scan_lengths(const int64_t* sizes, int64_t n, int64_t clamp_max, int64_t* off, int64_t* stats,
                     void* ws, size_t ws_bytes, rua_stream_t stream);

/* the same scan with two optional extras (either may be left out by passing NULL):
 *  - sizes == NULL: the lengths are those of a PackedSequence, len[i] = #{t : bs[t] > unsorted[i]} (core/view.py:21-25
 *    on a P source), computed on the fly and also written to len_out (n entries): P -> token_sizes + offsets in ONE launch;
 *  - notify_host_mapped != NULL: 3 int64 of PINNED HOST memory reachable from the device (UVA).  When the scan is
 *    complete the device writes [sum, max, ticket] there (ticket last, after a system-scope fence), so the host can
 *    learn N = sum and T = max by polling notify[2] == ticket instead of copying stats back and synchronising the
 *    stream.  ws: rua_scan_workspace_bytes(n). */
int rua_scan_lengths_ex(const int64_t* sizes, int64_t n, int64_t clamp_max, int64_t* off, int64_t* stats, void* ws,
                        size_t ws_bytes, const int64_t* bs, const int64_t* unsorted, int64_t Tp, int64_t* len_out,
                        int64_t* notify_host_mapped, int64_t ticket, rua_stream_t stream);

/* pinned host memory a kernel can write to directly (cudaHostAlloc mapped + portable): *device_ptr is what to pass as
 * notify_host_mapped, *host_ptr is where the host polls.  Set-up calls (they synchronise like cudaMalloc / cudaFree). */
int rua_pinned_alloc(size_t bytes, void** host_ptr, void** device_ptr);
int rua_pinned_free(void* host_ptr);

/* stable descending argsort of the lengths (ties by ascending index -- the documented deviation from
 * the reference's non-stable CPU sort, SURVEY.md 8c hazard 1) and its inverse.  T = max length. */
size_t rua_sort_workspace_bytes(int64_t B);
int rua_sort_lengths(const int64_t* len, int64_t B, int64_t T, int64_t* sorted, int64_t* unsorted,
                     void* ws, size_t ws_bytes, rua_stream_t stream);

/* stable ASCENDING argsort of int64 keys in [0, max_key] and its inverse (scatter_* sort their index
 * with this; same workspace as rua_sort_lengths), and the bucket boundaries of the sorted order:
 * off[m] = #{k : keys[k] < m} for m in [0, M]. */
int rua_sort_keys(const int64_t* keys, int64_t n, int64_t max_key, int64_t* sorted, int64_t* unsorted,
                  void* ws, size_t ws_bytes, rua_stream_t stream);
int rua_bucket_offsets(const int64_t* keys, const int64_t* sorted, int64_t n, int64_t M, int64_t* off,
                       rua_stream_t stream);

/* out[perm[j]] = j */
int rua_invert_permutation(const int64_t* perm, int64_t B, int64_t* out, rua_stream_t stream);

/* bs[t] = #{i : len[i] > t}, t in [0,T), given ANY permutation `sorted` that orders len
 * non-increasingly (ours or an injected one). */
int rua_batch_sizes(const int64_t* len, const int64_t* sorted, int64_t B, int64_t T, int64_t* bs,
                    rua_stream_t stream);

/* len[i] = #{t : bs[t] > unsorted[i]}  (P -> token_sizes; core/view.py:21-25 on a P source) */
int rua_lengths_from_pack(const int64_t* bs, const int64_t* unsorted, int64_t B, int64_t T,
                          int64_t* len, rua_stream_t stream);

/* One-launch metadata for B <= rua_meta_fused_max_batch() sequences (a single CTA; bitonic sort in
 * shared memory): off[B+1]; hostbuf[0] = N, hostbuf[1] = T; and, when `sorted` != NULL, the stable
 * descending permutation + inverse, hostbuf[2 + t] = batch_sizes[t] and poff[t] for t < min(T, cap),
 * poff[min(T,cap)] = N (or -1 if T > cap: re-run with a larger cap).  hostbuf has 2 + cap entries so
 * the caller can fetch N, T and batch_sizes with ONE device->host copy (pack_view, core/view.py:47-58). */
int64_t rua_meta_fused_max_batch(void);
int rua_meta_fused(const int64_t* len, int64_t B, int64_t* off, int64_t* sorted, int64_t* unsorted,
                   int64_t* hostbuf, int64_t* poff, int64_t cap, rua_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* K1/K2  ragged row map: the 12 layout conversions, the selects and their backward passes       */
/* replaces to_cat / cat_pack_to_left / right_to_left / to_pack / cat_pack_to_right /            */
/* left_to_right (torchrua/core/cast.py:8-71), {cat,left,pack,right}_getitem/_setitem with       */
/* (batch_ptr, token_ptr) keys (core/get.py:21-79, core/set.py:23-92), head / last / rev / roll  */
/* / trunc (torchrua/select/*.py), segment_head / segment_last (reduce.py:64-69)                 */
/* ------------------------------------------------------------------------------------------- */
/* For every row j of the destination storage (dst->rows rows of row_bytes bytes):
 *   decode j -> (i, t_d) in the destination layout; if it is padding -> pad;
 *   t_s = tmap(t_d); if t_s is outside the source's length -> pad;
 *   else copy source row src_row(i, t_s).
 * Every destination byte is written exactly once (padding fill is fused). */
int rua_row_map(const void* src, void* dst, int64_t row_bytes, const rua_ragged_t* ragged,
                const rua_side_t* src_side, const rua_side_t* dst_side, int32_t tmap,
                int64_t tmap_arg, int32_t pad_mode, const void* fill_host, int32_t fill_bytes,
                rua_stream_t stream);

/* fused consumer pattern: `X.left(fill)` AND `X.mask(zero, one, dtype)` (torchrua/core/cast.py:19-32 +
 * torchrua/mask.py:6-32, e.g. padded activations + the additive attention bias of fmask) from ONE decode of the
 * destination rows: rua_row_map towards a LEFT destination that also writes mask_out[i, t] (dst_side->rows elements of
 * mask_elem_bytes) = one where the row holds a token, zero where it is padding.  Rows of >= 128 bytes only (narrow
 * rows: RUA_ERR_UNSUPPORTED -- launch rua_row_map and rua_mask). */
int rua_row_map_mask(const void* src, void* dst, int64_t row_bytes, const rua_ragged_t* ragged,
                     const rua_side_t* src_side, const rua_side_t* dst_side, const void* fill_host, int32_t fill_bytes,
                     const void* zero_host, const void* one_host, int32_t mask_elem_bytes, void* mask_out,
                     rua_stream_t stream);

/* constructors C/L/P/R.new(list) (torchrua/core/__init__.py:9-36): rua_row_map with an identity token map whose
 * SOURCE is a list -- sequence i is its own contiguous (len[i], row_bytes) allocation src_list[i].  src_list is a
 * DEVICE array of B device pointers; src_align = a power of two that divides every non-null pointer in it (the
 * kernel picks its vector width from it).  Replaces torch.cat + conversion: every payload byte moves once. */
int rua_row_map_list(const void* const* src_list, int32_t src_align, void* dst, int64_t row_bytes,
                     const rua_ragged_t* ragged, const rua_side_t* dst_side, const void* fill_host,
                     int32_t fill_bytes, rua_stream_t stream);

/* compose (torchrua/compose.py:9-33: torch.cat(data)[indices]) without materialising the concatenation: dst[j] = row
 * index[j] of the VIRTUAL concatenation of n_src tensors.  src_list: DEVICE array of n_src device pointers (whole
 * contiguous (rows_k, row_bytes) tensors); bases: DEVICE array of n_src + 1 int64, bases[k] = first row of tensor k in the
 * concatenation, bases[n_src] = total rows; src_align as in rua_row_map_list.  Every payload byte moves once. */
int rua_gather_rows_multi(const void* const* src_list, const int64_t* bases, int32_t n_src, int32_t src_align,
                          const int64_t* index, int64_t n, int64_t row_bytes, void* dst, rua_stream_t stream);

/* dst[j] = src[index[j]] for j < n   (tensor_getitem / Z-keyed getitem, core/get.py:11-31).
 * Negative indices wrap (index + src_rows) like torch advanced indexing; an index that is still out of range is
 * not dereferenced: the row is zero-filled (gather) or skipped (scatter) and counted (rua_index_error_count). */
int rua_gather_rows(const void* src, int64_t src_rows, const int64_t* index, int64_t n,
                    int64_t row_bytes, void* dst, rua_stream_t stream);
/* dst[index[j]] = src[j] for j < n   (tensor_setitem / Z-keyed setitem, core/set.py:10-31) */
int rua_scatter_rows(const void* src, const int64_t* index, int64_t n, int64_t row_bytes, void* dst,
                     int64_t dst_rows, rua_stream_t stream);

/* flat storage rows of layout position keys -- the (batch_ptr, token_ptr) branches of
 * {cat,left,pack,right}_getitem / _setitem (torchrua/core/get.py:21-82, core/set.py:23-92):
 *   C: min(off[b], rows-1) + t     L: b*W + t     R: b*W + (T - len[b]) + t     P: unsorted[b] + min(poff[t], rows-1)
 * with the clamps of C.offsets() / P.offsets() (layout/cat.py:79-81, pack.py:43-45), T = max length for R
 * (layout/right.py:62-67: size()[1], not the storage width W = side->width) and torch's wrap-around of negative
 * indices where the reference's indexing applies it.  batch_ptr == NULL: token_ptr holds flat rows already; they
 * are wrapped and bounds-checked only.  Out-of-range keys yield side->rows (one past the end: rua_gather_rows zero-fills it,
 * rua_scatter_rows skips it, an ascending sort keeps it last) and bump the error counter below.  Feed rows_out to rua_gather_rows / rua_scatter_rows. */
int rua_token_rows(const rua_ragged_t* ragged, const rua_side_t* side, int64_t T, const int64_t* batch_ptr,
                   const int64_t* token_ptr, int64_t n, int64_t* rows_out, rua_stream_t stream);

/* number of out-of-range indices seen by rua_gather_rows / rua_scatter_rows / rua_token_rows on the current device
 * since the last reset.  Those rows were NOT dereferenced (ATen device-asserts in the same situation).  This call
 * synchronises (one 8-byte device->host copy): diagnostics, not data path. */
int rua_index_error_count(int64_t* count_host, int32_t reset);

/* ------------------------------------------------------------------------------------------- */
/* K3  mask / index emit (no payload reads)                                                      */
/* replaces mask / bmask / fmask (torchrua/mask.py:6-32), get_mask (core/view.py:11-18),         */
/* major_sizes_to_ptr (utils.py:7-13), C/L/R.ptr, P.ptr, L.idx, R.idx (layout/*.py)              */
/* ------------------------------------------------------------------------------------------- */
/* out (B, W) elements of elem_bytes in {1,2,4,8}: `one` where t < len[i] else `zero` (left aligned
 * for every layout, mask.py:10). */
int rua_mask(const int64_t* len, int64_t B, int64_t W, const void* zero_host, const void* one_host,
             int32_t elem_bytes, void* out, rua_stream_t stream);

/* enumerate the n = off[S] positions of a segmented range: which[k] = segment id, within[k] = k -
 * off[which[k]].  Either output may be NULL.  If `relabel` is non-NULL, within[k] = relabel[k -
 * off[which[k]]] (P.ptr: segments are time steps, within = sorted_indices[rank]).  With stride != 0 also emits
 * flat[k] = which*stride + within + (align ? (stride - len[which]) : 0)   (L.idx / R.idx). */
int rua_emit_ptr(const int64_t* off, int64_t S, int64_t n, const int64_t* relabel, int64_t* which,
                 int64_t* within, int64_t* flat, int64_t stride, int32_t right_align,
                 rua_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* K4  segment reduce                                                                            */
/* replaces segment_max/min/sum/mean/prod/logsumexp (torchrua/reduce.py:34-61)                   */
/* ------------------------------------------------------------------------------------------- */
/* data (N, H) row-major, off (S+1) exclusive prefix of the segment sizes, out (S, H).
 * Accumulates in fp32 (fp64 for RUA_F64) whatever the storage dtype and rounds once.
 * Reference quirks reproduced (SURVEY.md 8c hazard 3): empty segments of max / logsumexp (min)
 * return the global min (max) of the whole tensor; a NaN anywhere poisons every output of
 * max / min / logsumexp; empty sum/mean -> 0, prod -> 1. */
size_t rua_segment_reduce_workspace_bytes(int64_t N, int64_t S, int64_t H, int32_t dtype, int32_t op);
int rua_segment_reduce(const void* data, const int64_t* off, int64_t N, int64_t S, int64_t H,
                       int32_t dtype, int32_t op, void* out, void* ws, size_t ws_bytes,
                       rua_stream_t stream);

/* the same reduction over GATHERED rows: segment s reduces data[row_index[r]] for r in [off[s], off[s+1]);
 * N = off[S] = number of entries of row_index.  scatter_* (torchrua/reduce.py:6-31) = stable sort of the
 * index + rua_bucket_offsets + this. */
int rua_segment_reduce_gather(const void* data, const int64_t* row_index, const int64_t* off, int64_t N,
                              int64_t S, int64_t H, int32_t dtype, int32_t op, void* out, void* ws,
                              size_t ws_bytes, rua_stream_t stream);

/* PARITY MODE for sum / mean / prod: the reference's order of operations replayed exactly -- per (segment,
 * column) strictly left to right in the STORAGE dtype with one rounding per step, mean divides by the length
 * converted to the storage dtype (what torch.segment_reduce computes for torchrua/reduce.py:44-53; SURVEY.md
 * 8c hazard 2).  Bit-identical to the reference for fp32 / fp64 / fp16 / bf16.  No workspace.  Other ops return
 * RUA_ERR_UNSUPPORTED (max / min are bit-exact in rua_segment_reduce already). */
int rua_segment_reduce_strict(const void* data, const int64_t* off, int64_t N, int64_t S, int64_t H,
                              int32_t dtype, int32_t op, void* out, rua_stream_t stream);

/* backward twin (ATen SegmentReduceBackward0 semantics): sum -> broadcast, mean ->
 * broadcast / len, max/min -> split evenly among ties, prod -> grad*out/x (exact when x != 0, else
 * product of the others), logsumexp -> softmax weights. */
size_t rua_segment_reduce_backward_workspace_bytes(int64_t N, int64_t S, int64_t H, int32_t dtype,
                                                   int32_t op);
int rua_segment_reduce_backward(const void* grad_out, const void* out, const void* data,
                                const int64_t* off, int64_t N, int64_t S, int64_t H, int32_t dtype,
                                int32_t op, void* grad_data, void* ws, size_t ws_bytes,
                                rua_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* K5  multi-GPU output gather over NVLink peer memory (the reference is         */
/* single-device).  One process per GPU; the batch is sharded by sequence; every rank stores its */
/* rows directly into the output buffers ("windows") of all ranks of the node.                   */
/* ------------------------------------------------------------------------------------------- */
#define RUA_MAX_DESTINATIONS 16
#define RUA_PEER_HANDLE_BYTES 64

/* window = device allocation that other processes of the node can map (CUDA IPC).  alloc returns the
 * local pointer and a 64-byte handle to ship to the peers; open maps a peer's window into this
 * process (enables peer access lazily); close unmaps; free releases the local allocation.  These four
 * calls synchronise the device (cudaMalloc / cudaFree semantics); they are set-up, not data path. */
int rua_peer_window_alloc(size_t bytes, void** ptr, void* handle_host);
int rua_peer_window_open(const void* handle_host, void** ptr);
int rua_peer_window_close(void* ptr);
int rua_peer_window_free(void* ptr);

/* fused "local layout -> global C on every rank": the n_tokens tokens of the local shard are walked in cat
 * order; token (i, t) is read ONCE from the local source layout (src_side: C, L, R or P over `ragged`) and
 * stored at row dst_base_host[k][i] + t of destination k, for every k < n_dst <= RUA_MAX_DESTINATIONS.
 * dst_base_host[k] is a DEVICE array of B int64 (first row of local sequence i in destination k), or NULL
 * for "local cat row" (a contiguous local copy).  dst_host[k] are device pointers (local or peer windows);
 * both arrays themselves live on the host.  Replaces to_cat (torchrua/core/cast.py:8-16) + all_gather. */
int rua_row_map_multi(const void* src, int64_t row_bytes, const rua_ragged_t* ragged,
                      const rua_side_t* src_side, int64_t n_tokens, void* const* dst_host,
                      const int64_t* const* dst_base_host, int32_t n_dst, rua_stream_t stream);

/* per-sequence results (segment reductions, last(), head(1)): row j of src (n, row_bytes) is stored at row
 * index[j] of every destination. */
int rua_scatter_rows_multi(const void* src, const int64_t* index, int64_t n, int64_t row_bytes,
                           void* const* dst_host, int32_t n_dst, rua_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RUA_B200_H_ */
