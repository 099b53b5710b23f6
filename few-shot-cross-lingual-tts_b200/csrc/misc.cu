// Small memory-bound kernels around the GEMM engine: sinusoid add, casts / weight packing,
// broadcast row-vector add, column sums (bias gradients), the 256->1 predictor head.
#include <cstdlib>
#include "common.h"
#include "util.cuh"

namespace fs2 {

// y[b,t,:] = bf16(x[b,t,:] + pe[t,:])         (transformer/Models.py:155-157, 224-226)
template <typename InT>
__global__ void __launch_bounds__(256)
posenc_add_kernel(const InT* __restrict__ x, const float* __restrict__ pe, long long n_vec, int T, int C,
                  long long x_batch_stride, __nv_bfloat16* __restrict__ y) {
  pdl_sync();
  const int vec_per_row = C / 8;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
       v += (long long)gridDim.x * blockDim.x) {
    const long long row = v / vec_per_row;
    const int c = (v - row * vec_per_row) * 8;
    const int b = row / T, t = row - (long long)b * T;
    const InT* xp = x + (long long)b * x_batch_stride + (long long)t * C + c;
    float f[8];
    if constexpr (sizeof(InT) == 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(xp);
      const float4 a1 = *reinterpret_cast<const float4*>(xp + 4);
      f[0] = a0.x; f[1] = a0.y; f[2] = a0.z; f[3] = a0.w;
      f[4] = a1.x; f[5] = a1.y; f[6] = a1.z; f[7] = a1.w;
    } else {
      unpack8(ld8(reinterpret_cast<const __nv_bfloat16*>(xp)), f);
    }
    const float4 p0 = *reinterpret_cast<const float4*>(pe + (long long)t * C + c);
    const float4 p1 = *reinterpret_cast<const float4*>(pe + (long long)t * C + c + 4);
    f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w;
    f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
    st8(y + row * C + c, pack8(f));
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ y) {
  pdl_sync();
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 a = *reinterpret_cast<const float4*>(x + i);
      __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
      *reinterpret_cast<uint2*>(y + i) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {
      for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, long long n, float* __restrict__ y) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

// Conv1d.weight [Co][Ci][k] f32  ->  wp [Co][k][Cpad] bf16 (zero padded channels)
__global__ void __launch_bounds__(256)
pack_conv_weight_kernel(const float* __restrict__ w, int Co, int Ci, int k, int Cpad,
                        __nv_bfloat16* __restrict__ wp) {
  pdl_sync();
  const long long n = (long long)Co * k * Cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = i % Cpad;
    const int tap = (i / Cpad) % k;
    const int co = i / ((long long)Cpad * k);
    wp[i] = ci < Ci ? __float2bfloat16(w[((long long)co * Ci + ci) * k + tap]) : __float2bfloat16(0.f);
  }
}

// x[b,t,:] += e[b,:]  in place (fastspeech2m.py:89,101,136: `output += emb.unsqueeze(1).expand(...)`)
__global__ void __launch_bounds__(256)
add_rowvec_kernel(const __nv_bfloat16* x, const float* __restrict__ e, long long n_vec, int T, int C,
                  __nv_bfloat16* y) {
  pdl_sync();
  const int vec_per_row = C / 8;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
       v += (long long)gridDim.x * blockDim.x) {
    const long long row = v / vec_per_row;
    const int c = (v - row * vec_per_row) * 8;
    const int b = row / T;
    float f[8];
    unpack8(ld8(x + row * C + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += e[(long long)b * C + c + j];
    st8(y + row * C + c, pack8(f));
  }
}

// out = a (f32) + b (bf16)
__global__ void __launch_bounds__(256)
add_f32_bf16_kernel(const float* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n,
                    float* __restrict__ out) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + __bfloat162float(b[i]);
}

// out[g, c] += sum_{r in group g} x[g*rows_per_group + r, c]     (bias grads: groups = 1;
// speaker-embedding grads: groups = B, rows_per_group = T).  x bf16 with row stride ld (C % 8 == 0).
// Each thread owns 8 consecutive channels (one 16-byte load per row), 4 independent rows in flight.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows_per_group, int C,
              int rows_per_block, float* __restrict__ out, const int64_t* __restrict__ lens,
              int out_group_stride, int seg_stride, float* __restrict__ out_b) {
  pdl_sync();
  extern __shared__ float s_acc[];  // [C]
  const int g = blockIdx.y;
  if (blockIdx.z) {  // second column segment of the same rows (e.g. the V block next to the Q block of dQKV)
    x += (long long)blockIdx.z * seg_stride;
    out = out_b;
  }
  const long long x_group_stride = (long long)rows_per_group * ld;
  if (lens) {  // rows at / after lens[g] are padding (zero by contract): not read
    const long long l = lens[g];
    if ((long long)blockIdx.x * rows_per_block >= l) return;
    if (l < rows_per_group) rows_per_group = (int)l;
  }
  const int vpr = C / 8, rs = 256 / vpr;
  for (int i = threadIdx.x; i < C; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int v = threadIdx.x % vpr, ro = threadIdx.x / vpr;
  if (ro < rs) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows_per_group);
    const __nv_bfloat16* base = x + (long long)g * x_group_stride + v * 8;
    int r = r0 + ro;
    for (; r + 3 * rs < r1; r += 4 * rs) {
      bf16x8 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = ld8(base + (long long)(r + u * rs) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(t[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    for (; r < r1; r += rs) {
      float f[8];
      unpack8(ld8(base + (long long)r * ld), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[v * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(out + (long long)g * out_group_stride + i, s_acc[i]);
}

// grad[co][ci][tap] += packed[co][tap][ci]
__global__ void __launch_bounds__(256)
unpack_add_conv_grad_kernel(const float* __restrict__ packed, int Co, int Ci, int k, float* __restrict__ grad) {
  pdl_sync();
  const long long n = (long long)Co * Ci * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = i % k;
    const int ci = (i / k) % Ci;
    const long long co = i / ((long long)k * Ci);
    grad[i] += packed[(co * k + tap) * Ci + ci];
  }
}

// same for fp32 input (mel-space gradients)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long ld, int rows, int C, int rows_per_block,
                  float* __restrict__ out) {
  pdl_sync();
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += x[(long long)r * ld + c];
    atomicAdd(out + c, s);
  }
}

// C % 4 == 0, C <= 1024, 16-byte aligned rows: a thread owns four adjacent columns (one 16-byte load per row), the
// block's 256 / (C / 4) row lanes run side by side with four rows in flight each (the scalar kernel above read one
// float per thread and row: 30 us for the 20 MB of the mel-projection gradient).
__global__ void __launch_bounds__(256)
colsum_f32v4_kernel(const float* __restrict__ x, long long ld, int rows, int C, int rows_per_block,
                    float* __restrict__ out) {
  pdl_sync();
  __shared__ float s_acc[1024];
  const int c4 = C >> 2, rpp = 256 / c4;  // column groups, row lanes per pass
  for (int i = threadIdx.x; i < C; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x % c4, rl = threadIdx.x / c4;
  if (rl < rpp) {
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows);
    const float* base = x + cg * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int r = r0 + rl;
    for (; r + 3 * rpp < r1; r += 4 * rpp) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const float4*>(base + (long long)(r + u * rpp) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w;
      }
    }
    for (; r < r1; r += rpp) {
      const float4 t = *reinterpret_cast<const float4*>(base + (long long)r * ld);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomicAdd(&s_acc[cg * 4], acc.x);
    atomicAdd(&s_acc[cg * 4 + 1], acc.y);
    atomicAdd(&s_acc[cg * 4 + 2], acc.z);
    atomicAdd(&s_acc[cg * 4 + 3], acc.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(out + i, s_acc[i]);
}

// Predictor head: out[row] = mask(dot(x[row,:], w) + b)      (modules.py:242,246-252)
__global__ void __launch_bounds__(256)
rowdot_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ bias, const int64_t* __restrict__ lens, long long rows, int T,
                  int C, float* __restrict__ out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    float f[8];
    unpack8(ld8(x + row * C + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j] * w[c + j];
  }
  s = warp_sum(s);
  if (lane == 0) {
    const int b = row / T, t = row - (long long)b * T;
    out[row] = (lens && t >= lens[b]) ? 0.f : s + bias[0];
  }
}

// dx[row,:] = g*w ; dw += sum_rows g*x[row,:] ; db += sum_rows g, with g = masked dout[row]
__global__ void __launch_bounds__(256)
rowdot_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                  const float* __restrict__ w, const int64_t* __restrict__ lens, long long rows, int T,
                  int C, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
  pdl_sync();
  __shared__ float red[8][1024];
  __shared__ float red_b[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float accb = 0.f;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    const int b = row / T, t = row - (long long)b * T;
    const float g = (lens && t >= lens[b]) ? 0.f : dout[row];
    accb += g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 256 + lane * 8;
      if (c < C) {
        float f[8], o[8];
        unpack8(ld8(x + row * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] += g * f[j];
          o[j] = g * w[c + j];
        }
        st8(dx + row * C + c, pack8(o));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][i * 256 + lane * 8 + j] = acc[i][j];
  if (lane == 0) red_b[warp] = accb;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][c];
    atomicAdd(dw + c, s);
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red_b[wv];
    atomicAdd(db, s);
  }
}

static unsigned grid_for(long long n, int per_block, int cap = 148 * 16) {
  long long g = (n + per_block - 1) / per_block;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace fs2

extern "C" {

int fs2_posenc_add(const void* x, int x_is_f32, int64_t x_batch_stride, const float* pe, int B, int T,
                   int C, void* y, void* stream) {
  if (C % 8) return fs2::set_error("posenc_add: C must be a multiple of 8");
  const long long n_vec = (long long)B * T * (C / 8);
  if (n_vec <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_is_f32)
    FS2_LAUNCH((fs2::posenc_add_kernel<float>), fs2::grid_for(n_vec, 256), 256, 0, s, 
        static_cast<const float*>(x), pe, n_vec, T, C, x_batch_stride, static_cast<__nv_bfloat16*>(y));
  else
    FS2_LAUNCH((fs2::posenc_add_kernel<__nv_bfloat16>), fs2::grid_for(n_vec, 256), 256, 0, s, 
        static_cast<const __nv_bfloat16*>(x), pe, n_vec, T, C, x_batch_stride,
        static_cast<__nv_bfloat16*>(y));
  fs2::count_launch();
  return fs2::check_launch("posenc_add_kernel");
}

int fs2_cast_f32_bf16(const float* x, int64_t n, void* y, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7))
    return fs2::set_error("cast_f32_bf16: misaligned pointers");
  FS2_LAUNCH((fs2::cast_f32_bf16_kernel), fs2::grid_for(n, 1024), 256, 0, static_cast<cudaStream_t>(stream), 
      x, n, static_cast<__nv_bfloat16*>(y));
  fs2::count_launch();
  return fs2::check_launch("cast_f32_bf16_kernel");
}

int fs2_cast_bf16_f32(const void* x, int64_t n, float* y, void* stream) {
  if (n <= 0) return 0;
  FS2_LAUNCH((fs2::cast_bf16_f32_kernel), fs2::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), n, y);
  fs2::count_launch();
  return fs2::check_launch("cast_bf16_f32_kernel");
}

int fs2_pack_conv_weight(const float* w, int Co, int Ci, int k, int Cpad, void* wp, void* stream) {
  const long long n = (long long)Co * k * Cpad;
  if (n <= 0) return 0;
  if (Cpad < Ci) return fs2::set_error("pack_conv_weight: Cpad < Ci");
  FS2_LAUNCH((fs2::pack_conv_weight_kernel), fs2::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, Co, Ci, k, Cpad, static_cast<__nv_bfloat16*>(wp));
  fs2::count_launch();
  return fs2::check_launch("pack_conv_weight_kernel");
}

int fs2_add_f32_bf16(const float* a, const void* b, int64_t n, float* out, void* stream) {
  if (n <= 0) return 0;
  FS2_LAUNCH((fs2::add_f32_bf16_kernel), fs2::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      a, static_cast<const __nv_bfloat16*>(b), n, out);
  fs2::count_launch();
  return fs2::check_launch("add_f32_bf16_kernel");
}

int fs2_add_rowvec_bf16(const void* x, const float* e, int B, int T, int C, void* y, void* stream) {
  if (C % 8) return fs2::set_error("add_rowvec: C must be a multiple of 8");
  const long long n_vec = (long long)B * T * (C / 8);
  if (n_vec <= 0) return 0;
  FS2_LAUNCH((fs2::add_rowvec_kernel), fs2::grid_for(n_vec, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), e, n_vec, T, C, static_cast<__nv_bfloat16*>(y));
  fs2::count_launch();
  return fs2::check_launch("add_rowvec_kernel");
}

// out f32 [groups][C] += column sums of x bf16 [groups*rows_per_group][ld]
static int colsum_launch(const void* x, int64_t ld, int groups, int rows_per_group, int C, float* out,
                         const int64_t* lens, int out_group_stride, void* stream, int seg_stride = 0,
                         float* out_b = nullptr) {
  if (groups <= 0 || rows_per_group <= 0) return 0;
  if (getenv("FS2_DBG_SKIP_COLSUM")) return 0;
  if ((ld % 8) || (C % 8) || C / 8 > 256 || (reinterpret_cast<uintptr_t>(x) & 15))
    return fs2::set_error("colsum: ld and C must be multiples of 8 (16-byte aligned rows), C <= 2048");
  int rpb = (rows_per_group * groups + 148 * 8 - 1) / (148 * 8);
  if (rpb < 32) rpb = 32;
  dim3 grid((rows_per_group + rpb - 1) / rpb, groups, out_b ? 2 : 1);
  FS2_LAUNCH((fs2::colsum_kernel), grid, 256, C * sizeof(float), static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), ld, rows_per_group, C, rpb, out, lens, out_group_stride, seg_stride,
      out_b);
  fs2::count_launch();
  return fs2::check_launch("colsum_kernel");
}

int fs2_colsum_bf16(const void* x, int64_t ld, int groups, int rows_per_group, int C, float* out,
                    void* stream) {
  return colsum_launch(x, ld, groups, rows_per_group, C, out, nullptr, C, stream);
}

int fs2_colsum_ragged_bf16(const void* x, int64_t ld, int B, int T, int C, const int64_t* lens, float* out,
                           void* stream) {
  return colsum_launch(x, ld, B, T, C, out, lens, 0, stream);
}

int fs2_colsum_ragged2_bf16(const void* x, int64_t ld, int B, int T, int C, const int64_t* lens, int seg_stride,
                            float* out0, float* out1, void* stream) {
  if ((seg_stride % 8) || !out1) return fs2::set_error("colsum_ragged2: segment stride must be a multiple of 8 columns");
  return colsum_launch(x, ld, B, T, C, out0, lens, 0, stream, seg_stride, out1);
}

int fs2_colsum3_bf16(const void* x, int64_t ld, int rows, int seg_cols, float* out0, float* out1, float* out2,
                     void* stream) {
  float* outs[3] = {out0, out1, out2};
  for (int i = 0; i < 3; ++i) {
    const __nv_bfloat16* xi = static_cast<const __nv_bfloat16*>(x) + (long long)i * seg_cols;
    if (int rc = fs2_colsum_bf16(xi, ld, 1, rows, seg_cols, outs[i], stream)) return rc;
  }
  return 0;
}

int fs2_unpack_add_conv_grad(const float* packed, int Co, int Ci, int k, float* grad, void* stream) {
  const long long n = (long long)Co * Ci * k;
  if (n <= 0) return 0;
  FS2_LAUNCH((fs2::unpack_add_conv_grad_kernel), fs2::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      packed, Co, Ci, k, grad);
  fs2::count_launch();
  return fs2::check_launch("unpack_add_conv_grad_kernel");
}

int fs2_colsum_f32(const float* x, int64_t ld, int rows, int C, float* out, void* stream) {
  if (rows <= 0) return 0;
  if (C % 4 == 0 && C >= 4 && C <= 1024 && ld % 4 == 0 && !(reinterpret_cast<uintptr_t>(x) & 15)) {
    int rpb = (rows + 148 * 4 - 1) / (148 * 4);
    if (rpb < 32) rpb = 32;
    FS2_LAUNCH((fs2::colsum_f32v4_kernel), (rows + rpb - 1) / rpb, 256, 0, static_cast<cudaStream_t>(stream),
        x, ld, rows, C, rpb, out);
    fs2::count_launch();
    return fs2::check_launch("colsum_f32v4_kernel");
  }
  const int rpb = 64;
  FS2_LAUNCH((fs2::colsum_f32_kernel), (rows + rpb - 1) / rpb, 128, 0, static_cast<cudaStream_t>(stream), 
      x, ld, rows, C, rpb, out);
  fs2::count_launch();
  return fs2::check_launch("colsum_f32_kernel");
}

int fs2_rowdot_fwd(const void* x, const float* w, const float* bias, const int64_t* lens, int B, int T,
                   int C, float* out, void* stream) {
  if (C % 8 || C > 1024) return fs2::set_error("rowdot: C must be a multiple of 8, <= 1024");
  const long long rows = (long long)B * T;
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::rowdot_fwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), w, bias, lens, rows, T, C, out);
  fs2::count_launch();
  return fs2::check_launch("rowdot_fwd_kernel");
}

int fs2_rowdot_bwd(const float* dout, const void* x, const float* w, const int64_t* lens, int B, int T,
                   int C, void* dx, float* dw, float* db, void* stream) {
  if (C % 256 || C > 1024) return fs2::set_error("rowdot_bwd: C must be 256/512/768/1024");
  const long long rows = (long long)B * T;
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::rowdot_bwd_kernel), fs2::grid_for(rows, 32, 148 * 2), 256, 0, static_cast<cudaStream_t>(stream), 
      dout, static_cast<const __nv_bfloat16*>(x), w, lens, rows, T, C, static_cast<__nv_bfloat16*>(dx),
      dw, db);
  fs2::count_launch();
  return fs2::check_launch("rowdot_bwd_kernel");
}
}

// ------------------------------------------------------------------------------------------------
// Device-side collate: dst[b][t][:] = t < len_b ? src[offsets[b] + t][:] : 0 for ragged rows
// concatenated in `src` -- the pad_1D / pad_2D of the reference's host collate (lightning/utils/tool.py:134-165,
// lightning/collates/utils.py:8-85) done after ONE host -> device copy of the un-padded bytes.
// ------------------------------------------------------------------------------------------------
namespace fs2 {
template <typename V>
__global__ void __launch_bounds__(256)
pad_ragged_kernel(const V* __restrict__ src, const long long* __restrict__ offsets, int B, int max_len, int vec_per_row,
                  V* __restrict__ dst) {
  pdl_sync();
  const long long n = (long long)B * max_len * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vec_per_row;
    const long long row = i / vec_per_row;
    const int t = row % max_len;
    const int b = row / max_len;
    const long long o0 = offsets[b], len = offsets[b + 1] - o0;
    V val = V();
    if (t < len) val = src[(o0 + t) * vec_per_row + v];
    dst[i] = val;
  }
}
}  // namespace fs2

extern "C" int fs2_pad_ragged(const void* src, const int64_t* offsets, int B, int max_len, int row_bytes, void* dst,
                              void* stream) {
  if (B <= 0 || max_len <= 0) return 0;
  if (row_bytes <= 0 || row_bytes % 4) return fs2::set_error("pad_ragged: row bytes must be a positive multiple of 4");
  const bool v16 = row_bytes % 16 == 0 && !(reinterpret_cast<uintptr_t>(src) & 15) && !(reinterpret_cast<uintptr_t>(dst) & 15);
  const long long n = (long long)B * max_len * (row_bytes / (v16 ? 16 : 4));
  const unsigned grid = fs2::grid_for(n, 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (v16)
    FS2_LAUNCH((fs2::pad_ragged_kernel<uint4>), grid, 256, 0, s, static_cast<const uint4*>(src),
                                                       reinterpret_cast<const long long*>(offsets), B, max_len,
                                                       row_bytes / 16, static_cast<uint4*>(dst));
  else
    FS2_LAUNCH((fs2::pad_ragged_kernel<uint32_t>), grid, 256, 0, s, static_cast<const uint32_t*>(src),
                                                          reinterpret_cast<const long long*>(offsets), B, max_len,
                                                          row_bytes / 4, static_cast<uint32_t*>(dst));
  fs2::count_launch();
  return fs2::check_launch("pad_ragged_kernel");
}
