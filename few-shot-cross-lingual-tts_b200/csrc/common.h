// Shared host-side helpers of libfs2b200.so: thread-local error message, launch counter,
// launch checking.  Kernels never allocate, free or synchronise (include/fs2b200.h conventions).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/fs2b200.h"

namespace fs2 {

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of this library starts with pdl_sync(): it releases the
// NEXT kernel of the stream (griddepcontrol.launch_dependents) and then waits until the PREVIOUS kernel has
// completed and flushed its memory (griddepcontrol.wait) before touching global memory.  Launched through
// FS2_LAUNCH (cudaLaunchKernelEx + cudaLaunchAttributeProgrammaticStreamSerialization) a kernel's launch latency
// and block scheduling overlap its predecessor instead of following it -- the step is a chain of ~370 dependent
// launches, and below ~5 ms per step (few-shot task batches) those gaps, not the kernels, bound it.  The edges
// are captured into the step's CUDA graph as programmatic dependencies.  FS2_NO_PDL=1 launches plainly.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}
// split form for kernels with a data-independent prologue (mbarrier init, TMEM allocation, tensor-map prefetch):
// pdl_trigger() first, the prologue, then pdl_wait() before the first access to global memory.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define FS2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)fs2::launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

int set_error(const char* msg);
int set_cuda_error(const char* what, cudaError_t e);
int check_launch(const char* kernel_name);
void count_launch();
int gemm_tc_launch(const fs2_gemm& g, cudaStream_t stream);
int gemm_simt_launch(const fs2_gemm& g, cudaStream_t stream);
long long conv_tc2_workspace_bytes(int pairs);  // conv_tc2.cu: fs2_gemm::workspace for `pairs` CTA pairs
}  // namespace fs2
