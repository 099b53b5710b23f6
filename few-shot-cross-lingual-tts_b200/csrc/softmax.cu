// Key-padding-masked softmax over attention scores, forward and backward.
//
// Replaces transformer/Modules.py:17-22 (`attn / temperature`, `masked_fill(mask, -inf)`,
// `softmax(dim=2)`) and its autograd.  The 1/sqrt(d_k) scale is applied by the GEMM epilogue that
// produced S.  The bool mask tensor of the reference (SubLayers.py:46 `mask.repeat(n_head,1,1)`) is
// replaced by the per-utterance lengths: key j of batch b is valid iff j < lens[b].
// Batch-of-heads index here is z = b*H + h (the reference uses h*B + b; internal only).
// Query rows >= lens[b] are padding (zeroed after the sub-layer, Layers.py:25): P = 0 there.
// One warp per score row, row kept in registers (Tp <= 2048).  HBM/L2-bound.
#include "common.h"
#include "util.cuh"

namespace fs2 {

constexpr int kMaxVec = 16;  // 16 float4 per lane = 2048 columns

__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const float* __restrict__ S, const int64_t* __restrict__ lens, int H, int T, int Tp,
                   long long rows, __nv_bfloat16* __restrict__ P) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int z = row / T, t = row - (long long)z * T;
  const int len = min((int)lens[z / H], T);
  const float* s = S + row * Tp;
  __nv_bfloat16* p = P + row * Tp;
  const int nvec = Tp >> 7;  // float4 vectors per lane (Tp is a multiple of 128... see host check)
  float4 v[kMaxVec];
  if (t >= len) {
    for (int c = lane * 4; c < Tp; c += 128) *reinterpret_cast<uint2*>(p + c) = make_uint2(0u, 0u);
    return;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const int c = i * 128 + lane * 4;
      v[i] = *reinterpret_cast<const float4*>(s + c);
      if (c + 0 >= len) v[i].x = -INFINITY;
      if (c + 1 >= len) v[i].y = -INFINITY;
      if (c + 2 >= len) v[i].z = -INFINITY;
      if (c + 3 >= len) v[i].w = -INFINITY;
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      v[i].x = __expf(v[i].x - mx);
      v[i].y = __expf(v[i].y - mx);
      v[i].z = __expf(v[i].z - mx);
      v[i].w = __expf(v[i].w - mx);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float inv = 1.f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const int c = i * 128 + lane * 4;
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x * inv, v[i].y * inv);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v[i].z * inv, v[i].w * inv);
      *reinterpret_cast<uint2*>(p + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

// dS = alpha * P o (dP - sum_k dP*P)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP,
                   const int64_t* __restrict__ lens, int H, int T, int Tp, long long rows, float alpha,
                   __nv_bfloat16* __restrict__ dS) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int z = row / T, t = row - (long long)z * T;
  const int len = min((int)lens[z / H], T);
  __nv_bfloat16* ds = dS + row * Tp;
  if (t >= len) {
    for (int c = lane * 4; c < Tp; c += 128) *reinterpret_cast<uint2*>(ds + c) = make_uint2(0u, 0u);
    return;
  }
  const int nvec = Tp >> 7;
  float4 pv[kMaxVec], gv[kMaxVec];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const int c = i * 128 + lane * 4;
      const uint2 u = *reinterpret_cast<const uint2*>(P + row * Tp + c);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      pv[i] = make_float4(a.x, a.y, b.x, b.y);
      gv[i] = *reinterpret_cast<const float4*>(dP + row * Tp + c);
      // columns >= len hold P == 0 but dP there may be uninitialised (NaN): zero it explicitly
      if (c + 0 >= len) gv[i].x = 0.f;
      if (c + 1 >= len) gv[i].y = 0.f;
      if (c + 2 >= len) gv[i].z = 0.f;
      if (c + 3 >= len) gv[i].w = 0.f;
      dot += (pv[i].x * gv[i].x + pv[i].y * gv[i].y) + (pv[i].z * gv[i].z + pv[i].w * gv[i].w);
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const int c = i * 128 + lane * 4;
      __nv_bfloat162 lo = __floats2bfloat162_rn(alpha * pv[i].x * (gv[i].x - dot),
                                                alpha * pv[i].y * (gv[i].y - dot));
      __nv_bfloat162 hi = __floats2bfloat162_rn(alpha * pv[i].z * (gv[i].z - dot),
                                                alpha * pv[i].w * (gv[i].w - dot));
      *reinterpret_cast<uint2*>(ds + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

}  // namespace fs2

extern "C" {

// S,dP: f32 [Z][T][Tp];  P,dS: bf16 [Z][T][Tp];  lens: int64 [Z/H];  Tp % 128 == 0, Tp <= 2048.
// Columns T..Tp-1 of S may hold anything; they are >= len, hence masked.
int fs2_softmax_fwd(const float* S, const int64_t* lens, int Z, int H, int T, int Tp, void* P,
                    void* stream) {
  if (Tp % 128 || Tp > 128 * fs2::kMaxVec || Tp < T) return fs2::set_error("softmax: bad Tp");
  const long long rows = (long long)Z * T;
  if (rows == 0) return 0;
  FS2_LAUNCH((fs2::softmax_fwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      S, lens, H, T, Tp, rows, static_cast<__nv_bfloat16*>(P));
  fs2::count_launch();
  return fs2::check_launch("softmax_fwd_kernel");
}

int fs2_softmax_bwd(const void* P, const float* dP, const int64_t* lens, int Z, int H, int T, int Tp,
                    float alpha, void* dS, void* stream) {
  if (Tp % 128 || Tp > 128 * fs2::kMaxVec || Tp < T) return fs2::set_error("softmax: bad Tp");
  const long long rows = (long long)Z * T;
  if (rows == 0) return 0;
  FS2_LAUNCH((fs2::softmax_bwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(P), dP, lens, H, T, Tp, rows, alpha,
      static_cast<__nv_bfloat16*>(dS));
  fs2::count_launch();
  return fs2::check_launch("softmax_bwd_kernel");
}
}
