// Shared host-side helpers of libfs2b200.so: thread-local error message, launch counter,
// launch checking.  Kernels never allocate, free or synchronise (include/fs2b200.h conventions).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/fs2b200.h"

namespace fs2 {
int set_error(const char* msg);
int set_cuda_error(const char* what, cudaError_t e);
int check_launch(const char* kernel_name);
void count_launch();
int gemm_tc_launch(const fs2_gemm& g, cudaStream_t stream);
int gemm_simt_launch(const fs2_gemm& g, cudaStream_t stream);
}  // namespace fs2
