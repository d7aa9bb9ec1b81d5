// Kernel launch with the programmatic-dependent-launch attribute (sm_90+): the kernel may be scheduled while its
// predecessor in the stream drains; it must execute ptx::grid_dep_wait() before touching anything the predecessor
// wrote (every kernel launched through this helper does so as its first statement, or after a prologue that only
// reads constants).  DFIR_PDL=0 in the environment turns the attribute off (A/B measurements).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace dfir {

// DFIR_PDL is a bit mask: 1 tensor-core convs, 2 CUDA-core kernels of the training step / streamer, 4 weight gradient
enum : int { PDL_CONV = 1, PDL_SIMT = 2, PDL_WGRAD = 4 };
inline int pdl_mask() {
  static const int m = getenv("DFIR_PDL") == nullptr ? (PDL_CONV | PDL_WGRAD) : atoi(getenv("DFIR_PDL"));
  return m;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int kind, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & kind) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace dfir
