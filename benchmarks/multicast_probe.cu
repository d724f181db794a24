// NVSwitch multicast probe: what does an all-gather of per-GPU slices cost when every row leaves the SM ONCE
// (multimem.st into a cuMulticast window bound on all GPUs) instead of once per destination (unicast peer stores)?
// Context: DESIGN.md section 6 / profiles/r2_gather_sweep.md -- the fused output gather (rua_row_map_multi) sits at
// ~595 GB/s out per rank at 8 GPUs with unicast stores, whatever the grid, the split or the store shape.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/_build/multicast_probe benchmarks/multicast_probe.cu -lcuda
//   benchmarks/_build/multicast_probe [slice_MiB=512] [reps=5]
//
// One process drives all visible GPUs.  Per GPU d: a source slice in local HBM; a window of n_gpus slices; slice d of every
// window receives GPU d's data.  Modes: "unicast" (n stores per 16 bytes through peer-mapped VAs) and "multicast" (one
// multimem.st per 16 bytes).  All GPUs send at once; time = max over GPUs of the event time; correctness by checksum.
// Prints one JSON line.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CU(x)                                                                         \
  do {                                                                                \
    CUresult r_ = (x);                                                                \
    if (r_ != CUDA_SUCCESS) {                                                         \
      const char* s_ = nullptr;                                                       \
      cuGetErrorString(r_, &s_);                                                      \
      snprintf(g_err, sizeof g_err, "%s -> %d %s", #x, (int)r_, s_ ? s_ : "?");       \
      return false;                                                                   \
    }                                                                                 \
  } while (0)
#define RT(x)                                                                         \
  do {                                                                                \
    cudaError_t r_ = (x);                                                             \
    if (r_ != cudaSuccess) {                                                          \
      snprintf(g_err, sizeof g_err, "%s -> %s", #x, cudaGetErrorString(r_));          \
      return false;                                                                   \
    }                                                                                 \
  } while (0)

static char g_err[512];
constexpr int kMaxGpus = 8;

struct Dsts {
  uint4* p[kMaxGpus];
  int n;
};

__global__ void fill_kernel(uint4* __restrict__ src, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)i * 2654435761u + seed;
    src[i] = make_uint4(x, x ^ 0x9e3779b9u, x + 17u, ~x);
  }
}

__global__ void unicast_kernel(const uint4* __restrict__ src, Dsts d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(src + i);
    const int k0 = (int)((i >> 7) % d.n);              // 2 KB runs start at different peers
    for (int kk = 0; kk < d.n; ++kk) {
      const int k = k0 + kk < d.n ? k0 + kk : k0 + kk - d.n;
      __stcs(d.p[k] + i, v);
    }
  }
}

__global__ void multicast_kernel(const uint4* __restrict__ src, uint4* mc, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(src + i);
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
  }
}

__global__ void checksum_kernel(const uint4* __restrict__ p, size_t n, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    s += (unsigned long long)v.x * 3u + v.y * 5ull + v.z * 7ull + v.w;
  }
  atomicAdd(out, s);
}

struct State {
  int n = 0;
  size_t slice = 0, win = 0, gran = 0;
  CUmemGenericAllocationHandle mem[kMaxGpus] = {}, mc = 0;
  CUdeviceptr uc_va[kMaxGpus] = {}, mc_va = 0;
  uint4* src[kMaxGpus] = {};
  unsigned long long* sum[kMaxGpus] = {};
  cudaStream_t st[kMaxGpus] = {};
  cudaEvent_t e0[kMaxGpus] = {}, e1[kMaxGpus] = {};
  bool have_mc = false;
};

static bool setup_windows(State& s, bool want_mc) {
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = (unsigned)s.n;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof ap);
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  ap.location.id = 0;
  CU(cuMemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (want_mc) {
    size_t g2 = 0;
    mp.size = s.win;
    CU(cuMulticastGetGranularity(&g2, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (g2 > gran) gran = g2;
  }
  s.gran = gran;
  s.win = (s.win + gran - 1) / gran * gran;
  if (want_mc) {
    mp.size = s.win;
    CU(cuMulticastCreate(&s.mc, &mp));
    for (int d = 0; d < s.n; ++d) {
      CUdevice dev;
      CU(cuDeviceGet(&dev, d));
      CU(cuMulticastAddDevice(s.mc, dev));
    }
  }
  std::vector<CUmemAccessDesc> acc(s.n);
  for (int d = 0; d < s.n; ++d) {
    acc[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[d].location.id = d;
    acc[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  for (int d = 0; d < s.n; ++d) {
    ap.location.id = d;
    CU(cuMemCreate(&s.mem[d], s.win, &ap, 0));
    CU(cuMemAddressReserve(&s.uc_va[d], s.win, gran, 0, 0));
    CU(cuMemMap(s.uc_va[d], s.win, 0, s.mem[d], 0));
    CU(cuMemSetAccess(s.uc_va[d], s.win, acc.data(), (size_t)s.n));
    if (want_mc) CU(cuMulticastBindMem(s.mc, 0, s.mem[d], 0, s.win, 0));
  }
  if (want_mc) {
    CU(cuMemAddressReserve(&s.mc_va, s.win, gran, 0, 0));
    CU(cuMemMap(s.mc_va, s.win, 0, s.mc, 0));
    CU(cuMemSetAccess(s.mc_va, s.win, acc.data(), (size_t)s.n));
    s.have_mc = true;
  }
  return true;
}

static bool run_mode(State& s, bool multicast, int reps, int blocks, double* ms_out, bool* ok_out) {
  const size_t nvec = s.slice / 16;
  for (int d = 0; d < s.n; ++d) {
    RT(cudaSetDevice(d));
    RT(cudaMemsetAsync((void*)s.uc_va[d], 0, s.win, s.st[d]));
    RT(cudaStreamSynchronize(s.st[d]));
  }
  double best = 1e30;
  for (int r = 0; r < reps + 1; ++r) {
    for (int d = 0; d < s.n; ++d) {
      RT(cudaSetDevice(d));
      RT(cudaDeviceSynchronize());
    }
    for (int d = 0; d < s.n; ++d) {
      RT(cudaSetDevice(d));
      RT(cudaEventRecord(s.e0[d], s.st[d]));
      if (multicast) {
        multicast_kernel<<<blocks, 256, 0, s.st[d]>>>(s.src[d], (uint4*)s.mc_va + (size_t)d * nvec, nvec);
      } else {
        Dsts ds;
        ds.n = s.n;
        for (int k = 0; k < s.n; ++k) ds.p[k] = (uint4*)s.uc_va[k] + (size_t)d * nvec;
        unicast_kernel<<<blocks, 256, 0, s.st[d]>>>(s.src[d], ds, nvec);
      }
      RT(cudaEventRecord(s.e1[d], s.st[d]));
    }
    double worst = 0;
    for (int d = 0; d < s.n; ++d) {
      RT(cudaSetDevice(d));
      RT(cudaStreamSynchronize(s.st[d]));
      RT(cudaGetLastError());
      float ms = 0;
      RT(cudaEventElapsedTime(&ms, s.e0[d], s.e1[d]));
      if (ms > worst) worst = ms;
    }
    if (r > 0 && worst < best) best = worst;              // first pass = warm-up
  }
  *ms_out = best;
  // every window must now hold every GPU's slice: compare checksums of window k, slice d with the source of d
  bool ok = true;
  std::vector<unsigned long long> want(s.n);
  for (int d = 0; d < s.n; ++d) {
    RT(cudaSetDevice(d));
    RT(cudaMemsetAsync(s.sum[d], 0, 8, s.st[d]));
    checksum_kernel<<<592, 256, 0, s.st[d]>>>(s.src[d], nvec, s.sum[d]);
    RT(cudaMemcpyAsync(&want[d], s.sum[d], 8, cudaMemcpyDeviceToHost, s.st[d]));
    RT(cudaStreamSynchronize(s.st[d]));
  }
  for (int k = 0; k < s.n; ++k) {
    RT(cudaSetDevice(k));
    for (int d = 0; d < s.n; ++d) {
      unsigned long long got = 0;
      RT(cudaMemsetAsync(s.sum[k], 0, 8, s.st[k]));
      checksum_kernel<<<592, 256, 0, s.st[k]>>>((const uint4*)s.uc_va[k] + (size_t)d * nvec, nvec, s.sum[k]);
      RT(cudaMemcpyAsync(&got, s.sum[k], 8, cudaMemcpyDeviceToHost, s.st[k]));
      RT(cudaStreamSynchronize(s.st[k]));
      if (got != want[d]) ok = false;
    }
  }
  *ok_out = ok;
  return true;
}

int main(int argc, char** argv) {
  const size_t slice_mib = argc > 1 ? (size_t)atoll(argv[1]) : 512;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  State s;
  s.slice = slice_mib << 20;
  if (cuInit(0) != CUDA_SUCCESS || cudaGetDeviceCount(&s.n) != cudaSuccess || s.n < 1) {
    printf("{\"error\": \"no CUDA device\"}\n");
    return 0;
  }
  if (s.n > kMaxGpus) s.n = kMaxGpus;
  s.win = s.slice * (size_t)s.n;
  int mc_attr = 1, fd_attr = 1, fabric_attr = 0;
  for (int d = 0; d < s.n; ++d) {
    CUdevice dev;
    int v = 0;
    cuDeviceGet(&dev, d);
    cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    mc_attr &= v;
    cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
    fd_attr &= v;
    cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev);
    fabric_attr |= v;
  }
  auto per_device = [&]() -> bool {
    for (int d = 0; d < s.n; ++d) {
      RT(cudaSetDevice(d));
      RT(cudaFree(0));
      RT(cudaStreamCreateWithFlags(&s.st[d], cudaStreamNonBlocking));
      RT(cudaEventCreate(&s.e0[d]));
      RT(cudaEventCreate(&s.e1[d]));
      RT(cudaMalloc(&s.src[d], s.slice));
      RT(cudaMalloc(&s.sum[d], 8));
      fill_kernel<<<592, 256, 0, s.st[d]>>>(s.src[d], s.slice / 16, 1000u + (uint32_t)d);
      RT(cudaStreamSynchronize(s.st[d]));
    }
    return true;
  };
  printf("{\"n_gpus\": %d, \"slice_MiB\": %zu, \"multicast_supported_attr\": %d, \"posix_fd_handles\": %d, \"fabric_handles\": %d", s.n,
         slice_mib, mc_attr, fd_attr, fabric_attr);
  if (!per_device()) {
    printf(", \"error\": \"%s\"}\n", g_err);
    return 0;
  }
  bool want_mc = mc_attr && s.n > 1;
  if (!setup_windows(s, want_mc)) {
    printf(", \"multicast_setup_error\": \"%s\"", g_err);
    if (!want_mc) {
      printf("}\n");
      return 0;
    }
    // plain windows (the failed attempt's allocations are left behind) so that the unicast numbers still exist
    s.have_mc = false;
    s.win = s.slice * (size_t)s.n;
    if (!setup_windows(s, false)) {
      printf(", \"window_setup_error\": \"%s\"}\n", g_err);
      return 0;
    }
  }
  printf(", \"granularity\": %zu", s.gran);
  const int grids[] = {148 * 2, 148 * 4, 148 * 8, 148 * 16};
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 1 && !s.have_mc) break;
    for (int g : grids) {
      double ms = 0;
      bool ok = false;
      if (!run_mode(s, mode == 1, reps, g, &ms, &ok)) {
        printf(", \"%s_error_grid%d\": \"%s\"}\n", mode ? "multicast" : "unicast", g, g_err);
        return 0;
      }
      // bytes every GPU RECEIVES from its peers per pass = (n - 1) slices; bytes it sends: unicast (n-1) slices, multicast 1
      const double in_gbs = (double)(s.n - 1) * (double)s.slice / 1e9 / (ms * 1e-3);
      printf(", \"%s_grid%d\": {\"ms\": %.3f, \"ingress_GBs_per_gpu\": %.1f, \"identical\": %s}", mode ? "multicast" : "unicast", g, ms, in_gbs,
             ok ? "true" : "false");
      fflush(stdout);
    }
  }
  printf("}\n");
  return 0;
}
