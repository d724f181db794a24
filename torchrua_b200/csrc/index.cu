// Layout position keys -- the index arithmetic of the reference's layout-aware __getitem__ / __setitem__.
//
// Replaces (reference file:line): the (batch_ptr, token_ptr) branches of core/get.py:21-82 and core/set.py:23-92
//   C: key = self.offsets()[b] + t                (offsets clamped to N-1, layout/cat.py:79-81)
//   L: self.data[b, t]
//   R: self.data[b, self.size()[1] - token_sizes[b] + t]     (size()[1] = max length, NOT data.size(1))
//   P: key = unsorted_indices[b] + self.offsets()[t]          (offsets of batch_sizes clamped to N-1, pack.py:43-45)
// i.e. four eager ATen ops (index, index, add, sub) and a cumsum per call.  Here one small kernel turns n key
// pairs into n flat storage rows, with torch's wrap-around of negative indices applied where the reference's
// indexing applies it (b over B; t over the row width for L / R, over T for P; the final row over the storage
// rows for C / P), and every row bounds-checked.  rua_gather_rows / rua_scatter_rows then move the payload.
//
// Also here: the per-device counter of out-of-range explicit indices (shared by rowmap.cu).
#include <cstring>

#include "common.cuh"

namespace rua {

static unsigned long long* g_index_errors[64] = {nullptr};

// one 8-byte device allocation per device, made the first time an indexed kernel is launched there
unsigned long long* index_error_counter() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_index_errors[dev]) {
    unsigned long long* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    if (cudaMemset(p, 0, sizeof(unsigned long long)) != cudaSuccess) { (void)cudaGetLastError(); cudaFree(p); return nullptr; }
    g_index_errors[dev] = p;
  }
  return g_index_errors[dev];
}

__global__ void __launch_bounds__(256)
token_rows_kernel(const rua_ragged_t rg, const rua_side_t sd, const int64_t T, const int64_t* __restrict__ bp,
                  const int64_t* __restrict__ tp, const int64_t n, int64_t* __restrict__ out,
                  unsigned long long* __restrict__ errors) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t t = __ldg(tp + k);
  const int64_t kBadRow = sd.rows;   // one past the end: rua_gather_rows zero-fills it, rua_scatter_rows skips it, a sort keeps it last
  int64_t row = kBadRow;
  if (!bp) {                                   // flat keys: wrap + bounds only
    if (t < 0) t += sd.rows;
    if (t >= 0 && t < sd.rows) row = t;
  } else {
    int64_t b = __ldg(bp + k);
    if (b < 0) b += rg.B;
    if (b >= 0 && b < rg.B) {
      if (sd.layout == RUA_CAT) {
        int64_t o = __ldg(rg.off + b);
        o = o < sd.rows - 1 ? o : sd.rows - 1;     // C.offsets() is clamped to N-1
        row = o + t;
        if (row < 0) row += sd.rows;
        if (row < 0 || row >= sd.rows) row = kBadRow;
      } else if (sd.layout == RUA_PACK) {
        if (t < 0) t += rg.Tp;
        if (t >= 0 && t < rg.Tp) {
          int64_t o = __ldg(rg.poff + t);
          o = o < sd.rows - 1 ? o : sd.rows - 1;   // P.offsets() is clamped to N-1
          row = __ldg(rg.unsorted + b) + o;
          if (row >= sd.rows) row = kBadRow;
        }
      } else {
        int64_t col = t;
        if (sd.layout == RUA_RIGHT) col += T - (__ldg(rg.off + b + 1) - __ldg(rg.off + b));
        if (col < 0) col += sd.width;
        if (col >= 0 && col < sd.width) row = b * sd.width + col;
      }
    }
  }
  if (row == kBadRow && errors) atomicAdd(errors, 1ull);
  out[k] = row;
}

}  // namespace rua

using namespace rua;

extern "C" {

int rua_token_rows(const rua_ragged_t* ragged, const rua_side_t* side, int64_t T, const int64_t* batch_ptr,
                   const int64_t* token_ptr, int64_t n, int64_t* rows_out, rua_stream_t stream) {
  if (!side || n < 0 || side->rows < 0) return RUA_ERR_INVALID;
  if (n == 0) return RUA_OK;
  if (!token_ptr || !rows_out) return RUA_ERR_INVALID;
  rua_ragged_t rg{};
  if (batch_ptr) {
    if (!ragged || !ragged->off || ragged->B <= 0) return RUA_ERR_INVALID;
    if (side->layout < RUA_CAT || side->layout > RUA_RIGHT) return RUA_ERR_INVALID;
    if (side->layout == RUA_PACK && (!ragged->poff || !ragged->unsorted)) return RUA_ERR_INVALID;
    if ((side->layout == RUA_LEFT || side->layout == RUA_RIGHT) && side->width <= 0) return RUA_ERR_INVALID;
    rg = *ragged;
  }
  const int64_t blocks = ceil_div(n, 256);
  if (blocks >= (1ll << 31)) return RUA_ERR_UNSUPPORTED;
  token_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rg, *side, T, batch_ptr, token_ptr, n, rows_out,
                                                                       index_error_counter());
  return check_launch();
}

/* pinned host memory that kernels can write directly (UVA mapping): the completion notices of rua_scan_lengths_ex */
int rua_pinned_alloc(size_t bytes, void** host_ptr, void** device_ptr) {
  if (!host_ptr || !device_ptr || bytes == 0) return RUA_ERR_INVALID;
  void* h = nullptr;
  int rc = check_cuda(cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  if (rc) return rc;
  void* d = nullptr;
  rc = check_cuda(cudaHostGetDevicePointer(&d, h, 0));
  if (rc) { cudaFreeHost(h); return rc; }
  memset(h, 0, bytes);
  *host_ptr = h;
  *device_ptr = d;
  return RUA_OK;
}

int rua_pinned_free(void* host_ptr) {
  if (!host_ptr) return RUA_OK;
  return check_cuda(cudaFreeHost(host_ptr));
}

int rua_index_error_count(int64_t* count_host, int32_t reset) {
  if (!count_host) return RUA_ERR_INVALID;
  unsigned long long* p = index_error_counter();
  if (!p) return RUA_ERR_CUDA;
  unsigned long long v = 0;
  int rc = check_cuda(cudaMemcpy(&v, p, sizeof(v), cudaMemcpyDeviceToHost));
  if (rc) return rc;
  *count_host = (int64_t)v;
  if (reset && v) return check_cuda(cudaMemset(p, 0, sizeof(v)));
  return RUA_OK;
}

}  // extern "C"
