// Particle exchange of the particle-sharded sweep over PEER MEMORY (BASELINE configs[4]: one chain, particle rows
// block-partitioned over the GPUs of one NVSwitch box): the ancestor gather reads every parent row straight from the GPU
// that owns it -- NVLink loads inside the gather kernel -- instead of staging rows through send / receive buffers.
// The owners' particle buffers are shared between the per-GPU processes with CUDA IPC handles (exported / imported
// here; the handles travel through torch.distributed on the host side, fbs_b200/sharded.py).
//
// Ordering contract (sharded.py): the buffers are used in ping-pong; every rank's all-gather of the step's log-weights,
// enqueued after its own writes of the step, is the point after which the peers may read them.
#include <cuda.h>
#include "fbs_common.cuh"

namespace fbs {

// dst[b, :] = srcs[idx[b] / rows_per_rank][idx[b] % rows_per_rank, :]   (16-byte loads when the row allows it)
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(const float* const* __restrict__ srcs,
                                                               const int32_t* __restrict__ idx, int64_t B, int64_t rowv,
                                                               int rows_per_rank, int rows_total, V* __restrict__ dst) {
  const int64_t total = B * rowv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / rowv;
    const int g = min(max(idx[b], 0), rows_total - 1);
    const int owner = g / rows_per_rank, local = g - owner * rows_per_rank;
    dst[t] = reinterpret_cast<const V*>(srcs[owner])[(int64_t)local * rowv + (t - b * rowv)];
  }
}

typedef CUresult (*PFN_memGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
static PFN_memGetAddressRange address_range_fn() {
  static PFN_memGetAddressRange fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_memGetAddressRange>(ptr);
  }
  return fn;
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_ipc_export(const void* dev_ptr, unsigned char* handle64, int64_t* offset) {
  FBS_REQUIRE(dev_ptr && handle64 && offset, "ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if (address_range_fn() == nullptr) {
    set_error("ipc_export: cuMemGetAddressRange is not available from this driver");
    return FBS_ERR_CUDA;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (address_range_fn()(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr) != CUDA_SUCCESS) {
    set_error("ipc_export: cuMemGetAddressRange failed");
    return FBS_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess) {
    set_error("ipc_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return FBS_OK;
}

int fbs_ipc_import(const unsigned char* handle64, int64_t offset, void** out_ptr) {
  FBS_REQUIRE(handle64 && out_ptr, "ipc_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("ipc_import: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  *out_ptr = reinterpret_cast<unsigned char*>(base) + offset;
  return FBS_OK;
}

int fbs_ipc_release(void* imported_ptr, int64_t offset) {
  if (imported_ptr == nullptr) return FBS_OK;
  cudaError_t e = cudaIpcCloseMemHandle(reinterpret_cast<unsigned char*>(imported_ptr) - offset);
  if (e != cudaSuccess) {
    set_error("ipc_release: cudaIpcCloseMemHandle failed: %s", cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  return FBS_OK;
}

int fbs_gather_rows_peer_f32(fbs_stream_t s, const float* const* srcs, const int32_t* idx, int64_t B, int64_t row,
                             int64_t rows_per_rank, int64_t n_ranks, float* dst) {
  if (B == 0) return FBS_OK;
  FBS_REQUIRE(srcs && idx && dst, "gather_rows_peer: null argument");
  FBS_REQUIRE(row >= 1 && rows_per_rank >= 1 && n_ranks >= 1 && rows_per_rank * n_ranks < (1ll << 31),
              "gather_rows_peer: bad sizes");
  const int rows_total = (int)(rows_per_rank * n_ranks);
  const bool v4 = row % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  const int64_t rowv = v4 ? row / 4 : row;
  int64_t blocks = (B * rowv + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (v4)
    gather_rows_peer_kernel<float4><<<(int)blocks, 256, 0, as_stream(s)>>>(srcs, idx, B, rowv, (int)rows_per_rank, rows_total,
                                                                          reinterpret_cast<float4*>(dst));
  else
    gather_rows_peer_kernel<float><<<(int)blocks, 256, 0, as_stream(s)>>>(srcs, idx, B, rowv, (int)rows_per_rank, rows_total,
                                                                         dst);
  return check_launch("gather_rows_peer_kernel");
}

}  // extern "C"
