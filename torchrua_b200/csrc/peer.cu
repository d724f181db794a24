// K5 (host side) -- peer-memory windows for the multi-GPU output gather.
//
// The reference has no multi-GPU support at all.  Here the batch is sharded by sequence (one process per
// GPU) and the only payload exchange is the final gather of outputs.  Instead of "all_gather padded
// shards, then permute" each rank's kernel stores its rows straight into the output buffers of all
// peers (rua_row_map_multi / rua_scatter_rows_multi in rowmap.cu).  For that every rank owns a WINDOW:
// device memory whose CUDA IPC handle is handed to the other processes of the node, which map it into
// their own address space (NVLink / NVSwitch peer mapping).  The handles travel through whatever
// control plane the host has (torch.distributed all_gather_object in torchrua_b200/shard.py).
#include <cstring>

#include "common.cuh"

using namespace rua;

extern "C" {

int rua_peer_window_alloc(size_t bytes, void** ptr, void* handle_host) {
  if (!ptr || !handle_host || bytes == 0) return RUA_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == RUA_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  int rc = check_cuda(cudaMalloc(&p, bytes));
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  rc = check_cuda(cudaIpcGetMemHandle(&h, p));
  if (rc) {
    cudaFree(p);
    return rc;
  }
  memcpy(handle_host, &h, sizeof(h));
  *ptr = p;
  return RUA_OK;
}

int rua_peer_window_open(const void* handle_host, void** ptr) {
  if (!handle_host || !ptr) return RUA_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  int rc = check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  if (rc) return rc;
  *ptr = p;
  return RUA_OK;
}

int rua_peer_window_close(void* ptr) {
  if (!ptr) return RUA_ERR_INVALID;
  return check_cuda(cudaIpcCloseMemHandle(ptr));
}

int rua_peer_window_free(void* ptr) {
  if (!ptr) return RUA_ERR_INVALID;
  return check_cuda(cudaFree(ptr));
}

}  // extern "C"
