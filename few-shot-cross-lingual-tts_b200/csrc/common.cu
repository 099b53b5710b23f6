#include "common.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fs2 {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local const char* g_last_kernel = "";

int set_error(const char* msg) {
  std::snprintf(g_err, sizeof g_err, "%s", msg);
  return 1;
}
int set_cuda_error(const char* what, cudaError_t e) {
  std::snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return 2;
}
int check_launch(const char* kernel_name) {
  g_last_kernel = kernel_name;  // string literal of the call site
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_cuda_error(kernel_name, e);
  }
  return 0;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = std::getenv("FS2_NO_PDL") == nullptr;
  return on;
}
}  // namespace fs2

extern "C" {
int fs2_version(void) { return FS2_ABI_VERSION; }
const char* fs2_last_error(void) { return fs2::g_err; }
int64_t fs2_launch_count(void) { return fs2::g_launches.load(); }
const char* fs2_last_kernel(void) { return fs2::g_last_kernel; }
int64_t fs2_stream_capture_id(void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo(static_cast<cudaStream_t>(stream), &st, &id) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return st == cudaStreamCaptureStatusActive ? static_cast<int64_t>(id) : 0;
}

int64_t fs2_gemm_workspace_bytes(void) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return fs2::conv_tc2_workspace_bytes(sms / 2);
}

int fs2_gemm_bf16(const fs2_gemm* g, int impl, void* stream) {
  if (!g) return fs2::set_error("fs2_gemm_bf16: null descriptor");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0) return fs2::gemm_tc_launch(*g, s);
  if (impl == 1) {
    if (g->a_colsum || g->a_colsum_seg[0] || g->a_colsum_seg[1] || g->a_colsum_seg[2] || g->a_colsum_seg[3]) return fs2::set_error("fs2_gemm_bf16: a_colsum is not implemented by the debug kernel (impl 1)");
    return fs2::gemm_simt_launch(*g, s);
  }
  return fs2::set_error("fs2_gemm_bf16: unknown impl");
}
}
