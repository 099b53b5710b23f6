// One launch that refreshes EVERY bf16 operand copy of the model's fp32 master weights for a training
// step: nn.Linear weights [N][K] are cast, nn.Conv1d weights [Co][Ci][k] are packed to [Co][k][Cpad]
// (csrc/gemm_tc.cu reads that layout as a K-major forward operand and as an MN-major input-gradient
// operand), and small fp32 vectors (the three Q/K/V biases) are gathered into one buffer.
// Replaces ~85 small launches per step (31 pack + 43 cast + 10 cat, 0.43 ms) by one pass over the
// 138 MB of parameters.  A block handles one output row: the row is read coalesced, transposed through
// shared memory when k > 1, written coalesced.
#include "common.h"
#include "util.cuh"

namespace fs2 {

// One entry per tensor; `row0` = first global row index of this entry (prefix sum of rows).
struct PrepEntry {
  const float* src;
  void* dst;       // bf16 [rows][k][cpad] (kind 0) or f32 [rows*ci] (kind 1)
  int rows, ci, k, cpad;
  int kind;        // 0: cast / pack to bf16, 1: f32 copy
  int row0;
};

constexpr int kPrepMaxRow = 4096;  // floats of one source row (Ci * k) held in shared memory

__global__ void __launch_bounds__(256) weight_prep_kernel(const PrepEntry* __restrict__ tab, int n_entries,
                                                          int total_rows) {
  pdl_sync();
  __shared__ __align__(16) float srow[kPrepMaxRow];
  for (int grow = blockIdx.x; grow < total_rows; grow += gridDim.x) {
    int lo = 0, hi = n_entries - 1;  // last entry with row0 <= grow
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].row0 <= grow) lo = mid; else hi = mid - 1;
    }
    const PrepEntry e = tab[lo];
    const int r = grow - e.row0;
    const int n_src = e.ci * e.k;
    const float* src = e.src + (long long)r * n_src;
    const bool v4 = !(n_src & 3) && !(reinterpret_cast<uintptr_t>(e.src) & 15);  // 16-byte loads of the source row
    if (e.kind == 1) {
      float* dst = static_cast<float*>(e.dst) + (long long)r * n_src;
      for (int i = threadIdx.x; i < n_src; i += blockDim.x) dst[i] = src[i];
      continue;
    }
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(e.dst) + (long long)r * e.k * e.cpad;
    const bool v8 = !(e.cpad & 7) && !(reinterpret_cast<uintptr_t>(e.dst) & 15);  // 16-byte stores of the bf16 row
    if (e.k == 1) {
      if (v4 && v8 && !(e.ci & 7)) {  // 8 floats in, one 16-byte bf16 vector out
        for (int i = threadIdx.x * 8; i < e.cpad; i += blockDim.x * 8) {
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (i < e.ci) {
            const float4 x0 = *reinterpret_cast<const float4*>(src + i), x1 = *reinterpret_cast<const float4*>(src + i + 4);
            f[0] = x0.x; f[1] = x0.y; f[2] = x0.z; f[3] = x0.w;
            f[4] = x1.x; f[5] = x1.y; f[6] = x1.z; f[7] = x1.w;
          }
          st8(dst + i, pack8(f));
        }
      } else {
        for (int i = threadIdx.x; i < e.cpad; i += blockDim.x) dst[i] = __float2bfloat16(i < e.ci ? src[i] : 0.f);
      }
      continue;
    }
    __syncthreads();  // previous iteration's readers are done with srow
    if (v4) {
      for (int i = threadIdx.x * 4; i < n_src; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(srow + i) = *reinterpret_cast<const float4*>(src + i);  // [ci][k], coalesced
    } else {
      for (int i = threadIdx.x; i < n_src; i += blockDim.x) srow[i] = src[i];
    }
    __syncthreads();
    if (v8) {  // thread = 8 consecutive channels of one tap: [k][cpad], 16-byte stores
      const int n_vec = e.k * (e.cpad >> 3);
      for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
        const int tap = i / (e.cpad >> 3), c0 = (i - tap * (e.cpad >> 3)) * 8;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = c0 + j < e.ci ? srow[(c0 + j) * e.k + tap] : 0.f;
        st8(dst + tap * e.cpad + c0, pack8(f));
      }
    } else {
      const int n_dst = e.k * e.cpad;
      for (int i = threadIdx.x; i < n_dst; i += blockDim.x) {
        const int tap = i / e.cpad, ci = i - tap * e.cpad;
        dst[i] = __float2bfloat16(ci < e.ci ? srow[ci * e.k + tap] : 0.f);
      }
    }
  }
}

}  // namespace fs2

extern "C" {

// table: DEVICE array of n_entries records {src, dst, rows, ci, k, cpad, kind, row0} laid out as
// fs2_prep_entry (include/fs2b200.h); total_rows = sum of rows.  Ci * k <= 4096 for packed entries.
int fs2_weight_prep(const void* table, int n_entries, int total_rows, void* stream) {
  if (n_entries <= 0 || total_rows <= 0) return 0;
  static_assert(sizeof(fs2::PrepEntry) == sizeof(fs2_prep_entry), "fs2_prep_entry layout");
  int grid = total_rows < 148 * 16 ? total_rows : 148 * 16;
  FS2_LAUNCH((fs2::weight_prep_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const fs2::PrepEntry*>(table), n_entries, total_rows);
  fs2::count_launch();
  return fs2::check_launch("weight_prep_kernel");
}
}
