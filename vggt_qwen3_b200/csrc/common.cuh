// Host-side helpers shared by the C-ABI translation units: error reporting and the TMA descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/vq_b200.h"

namespace vq {

void set_error(const std::string& msg);

#define VQ_CHECK_ARG(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) {                                               \
      ::vq::set_error(std::string("invalid argument: ") + (msg)); \
      return VQ_ERR_INVALID;                                     \
    }                                                            \
  } while (0)

#define VQ_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::vq::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));            \
      return VQ_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// 2-D bf16 tensor map with SWIZZLE_128B.  `inner`/`outer` are the tensor extents in elements,
// `ld` the row pitch in elements, box = {box_inner (must be 64), box_outer}.
// Copies the descriptor into *out; returns false (and sets the error string) on failure.  Descriptors are cached in a
// bounded LRU keyed by every argument; they are immutable.
bool get_tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer, CUtensorMap* out);

int num_sms();   // of the current device

// cudaFuncAttributeMaxDynamicSharedMemorySize (and optionally the max-shared carve-out) for `func` on the CURRENT device,
// set once per (device, function); thread-safe.
cudaError_t ensure_dynamic_smem(const void* func, int bytes, bool max_carveout = false);

// Library-owned scratch for the GEMM's K-sliced tail tiles, one per (device, stream) so that launches on different
// streams never share it: `bytes` of partial accumulators + `n_flags` zero-initialised counters (re-armed by the
// kernel itself).  Grown on demand (the old buffer is released after the stream has drained).
cudaError_t splitk_workspace(cudaStream_t stream, size_t bytes, size_t n_flags, void** ws, unsigned** flags);

// Programmatic dependent launch (VQ_PDL=0 disables it): kernels of the hot chain (GEMM, attention, LayerNorm) are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization so that their prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the tail of the previous kernel; see pdl_wait() in ptx.cuh.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg = cudaLaunchConfig_t();
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Launch accounting + optional CUDA-event profiling of every kernel the library launches.
// Always counts launches; when profiling is enabled (vq_profile_enable) it brackets the launch with
// two events on the launch stream and records (kind, tag, shape, algorithmic flops / bytes).
enum LaunchKind { kKindGemm = 0, kKindAttention = 1, kKindRow = 2, kNumKinds = 3 };
struct LaunchScope {
  LaunchScope(int kind, const char* tag, int m, int n, int k, double flops, double bytes, cudaStream_t stream);
  ~LaunchScope();
  int slot_;
  cudaStream_t stream_;
};

}  // namespace vq
