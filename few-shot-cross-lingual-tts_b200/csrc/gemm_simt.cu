// Plain CUDA-core implementation of fs2_gemm_bf16 (same descriptor, same semantics): one thread per
// output element, fp32 accumulation, no tiling.  It exists ONLY as an on-device cross-check of the
// tcgen05 engine (tests, FS2_GEMM_IMPL=simt debugging); the product path never selects it.
#include <cuda_bf16.h>

#include "common.h"

namespace fs2 {

struct SimtP {
  fs2_gemm g;
};

__device__ __forceinline__ float ld_bf16(const void* p, long long off) {
  return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[off]);
}

// element (row r, inner i, batch zb) with zero fill outside the tensor extents
__device__ __forceinline__ float fetch(const fs2_operand& o, int zb, long long r, long long i) {
  if (r < 0 || r >= o.rows || i < 0 || i >= o.inner || zb < 0 || zb >= o.batches) return 0.f;
  return ld_bf16(o.ptr, (long long)zb * o.batch_stride + r * o.ld + i);
}

__device__ __forceinline__ void epilogue_store(const fs2_gemm& g, float acc, int z, int m, int ncol,
                                               long long off) {
  acc *= g.alpha;
  if (g.bias) acc += g.bias[ncol];
  if (g.epilogue == FS2_EPI_RELU) acc = fmaxf(acc, 0.f);
  if (g.relu_mask) {  // 1-bit mask [Z*M][N/64]; the cross-check ORs its bit in (the caller zeroes the mask first)
    unsigned long long* w = static_cast<unsigned long long*>(g.relu_mask) + ((long long)z * g.M + m) * (g.N >> 6) + (ncol >> 6);
    if (g.epilogue == FS2_EPI_RELU) {
      if (acc > 0.f) atomicOr(w, 1ull << (ncol & 63));
    } else if (g.epilogue == FS2_EPI_RELU_BWD) {
      acc = ((*w >> (ncol & 63)) & 1ull) ? acc : 0.f;
    }
  } else
  if (g.epilogue == FS2_EPI_RELU_BWD || g.epilogue == FS2_EPI_ADD_AUX) {
    const float a = ld_bf16(g.aux, (long long)z * g.aux_batch_stride + (long long)m * g.ld_aux + ncol);
    acc = g.epilogue == FS2_EPI_RELU_BWD ? (a > 0.f ? acc : 0.f) : acc + a;
  }
  if (g.d_f32) {
    float* dbase = static_cast<float*>(g.d);
    if (g.d_seg_rows > 0) {  // off = m*ldd + ...: move the row into its segment
      dbase = static_cast<float*>(g.d_seg[m / g.d_seg_rows]);
      off -= (long long)(m / g.d_seg_rows) * g.d_seg_rows * g.ldd;
    }
    float* dp = dbase + off;
    if (g.d_atomic)
      atomicAdd(dp, acc);
    else
      *dp = acc;
  } else {
    static_cast<__nv_bfloat16*>(g.d)[off] = __float2bfloat16(acc);
  }
}

__global__ void gemm_simt_normal(const SimtP sp) {
  pdl_sync();
  const fs2_gemm& g = sp.g;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Z = g.Z > 0 ? g.Z : 1;
  const long long total = (long long)Z * g.M * g.N;
  if (idx >= total) return;
  const int n = idx % g.N;
  const int m = (idx / g.N) % g.M;
  const int z = idx / ((long long)g.N * g.M);
  const int azd = g.a.zdiv > 0 ? g.a.zdiv : 1, bzd = g.b.zdiv > 0 ? g.b.zdiv : 1;
  const int za = g.a.batches > 1 ? z / azd : 0, zb = g.b.batches > 1 ? z / bzd : 0;
  const long long ia = g.a.inner_base + (long long)(z % azd) * g.a.zmod_stride;
  const long long ib = g.b.inner_base + (long long)(z % bzd) * g.b.zmod_stride;
  const int taps = g.taps > 0 ? g.taps : 1;
  float acc = 0.f;
  for (int tap = 0; tap < taps; ++tap) {
    for (int k = 0; k < g.K; ++k) {
      float a, b;
      if (!g.a.mn_major)
        a = fetch(g.a, za, (long long)m + g.tap_shift0 + tap, ia + k);
      else
        a = fetch(g.a, za, k, ia + m);
      if (!g.b.mn_major)
        b = fetch(g.b, zb, n, ib + (long long)tap * g.b_tap_kstride + k);
      else
        b = fetch(g.b, zb, k, ib + (long long)tap * g.b_tap_kstride + n);
      acc += a * b;
    }
  }
  const int dzd = g.d_zdiv > 0 ? g.d_zdiv : 1;
  const long long off = (long long)(z / dzd) * g.d_zdiv_stride + (long long)(z % dzd) * g.d_zmod_stride +
                        (long long)m * g.ldd + n;
  if (g.row_lens) {  // padded rows are defined to be zero
    const int lzd = g.lens_zdiv > 0 ? g.lens_zdiv : 1;
    if (m >= g.row_lens[z / lzd]) {
      if (g.d_f32) static_cast<float*>(g.d)[off] = 0.f;
      else static_cast<__nv_bfloat16*>(g.d)[off] = __float2bfloat16(0.f);
      return;
    }
  }
  epilogue_store(g, acc, z, m, n, off);
}

__global__ void gemm_simt_wgrad(const SimtP sp) {
  pdl_sync();
  const fs2_gemm& g = sp.g;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int taps = g.taps > 0 ? g.taps : 1;
  const long long total = (long long)g.M * g.N * taps;
  if (idx >= total) return;
  const int n = idx % g.N;
  const int tap = (idx / g.N) % taps;
  const int m = idx / ((long long)g.N * taps);
  float acc = 0.f;
  for (int zb = 0; zb < g.a.batches; ++zb)
    for (int r = 0; r < g.a.rows; ++r) {
      const float a = fetch(g.a, zb, r, g.a.inner_base + m);
      if (a == 0.f) continue;
      acc += a * fetch(g.b, zb, (long long)r + g.tap_shift0 + tap, g.b.inner_base + n);
    }
  const long long cs = g.d_col_stride > 0 ? g.d_col_stride : 1;
  const long long off = (long long)m * g.ldd + (long long)tap * g.d_tap_stride + (long long)n * cs;
  epilogue_store(g, acc, 0, m, n, off);
}

int gemm_simt_launch(const fs2_gemm& g, cudaStream_t stream) {
  SimtP sp{g};
  const int taps = g.taps > 0 ? g.taps : 1;
  long long total;
  if (g.mode == FS2_GEMM_NORMAL) {
    total = (long long)(g.Z > 0 ? g.Z : 1) * g.M * g.N;
    if (total == 0) return 0;
    FS2_LAUNCH((gemm_simt_normal), (unsigned)((total + 255) / 256), 256, 0, stream, sp);
  } else {
    total = (long long)g.M * g.N * taps;
    if (total == 0) return 0;
    FS2_LAUNCH((gemm_simt_wgrad), (unsigned)((total + 255) / 256), 256, 0, stream, sp);
  }
  count_launch();
  return check_launch("gemm_simt");
}

}  // namespace fs2
