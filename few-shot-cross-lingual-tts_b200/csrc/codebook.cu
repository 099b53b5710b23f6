// Phoneme-embedding front-end of the few-shot (FSCL) systems, SURVEY.md 8f row 2:
//   * PhonemeQueryExtractor, mode "average" (lightning/model/reduction.py:42-110): phoneme-level mean of the SSL
//     frames of every segment, then the class-wise mean over all segments of the same phoneme;
//   * SoftMultiAttCodebook2 (lightning/systems/language/embeddings.py:77-142): softmax-weighted sum over the SSL
//     layers (NaN -> 0), q_linear (tcgen05 GEMM, csrc/gemm_tc*.cu), multi-head attention of every query over the
//     `att_banks` keys with the `emb_banks` values (transformer/Modules.py:28-47), forward and backward.
// Sizes are small (<= a few hundred queries, 128 codes, 256 dims): one block per query, one warp per head.
#include "common.h"
#include "util.cuh"

namespace fs2 {

// table_sum[cls[i]][:] += mean_{t in segment i} x[t][:]  (two_stage) or += sum (frame level); count likewise.
// grid: (L segments, ceil(D / 1024)); block 256 threads, 4 floats per thread.
__global__ void __launch_bounds__(256)
segment_class_accum_kernel(const float* __restrict__ x, const long long* __restrict__ dur, const long long* __restrict__ cls,
                           int L, long long T, long long D, int n_classes, int two_stage, float* __restrict__ table_sum,
                           float* __restrict__ count) {
  pdl_sync();
  const int i = blockIdx.x;
  __shared__ long long s_t0;
  if (threadIdx.x == 0) {  // exclusive prefix of the (clamped) durations before segment i
    long long t0 = 0;
    for (int j = 0; j < i; ++j) t0 += dur[j] > 0 ? dur[j] : 0;
    s_t0 = t0;
  }
  __syncthreads();
  const long long d = dur[i];
  const long long c = cls[i];
  if (d <= 0 || c < 0 || c >= n_classes) return;
  const long long t0 = s_t0;
  long long t1 = t0 + d;
  if (t1 > T) t1 = T;
  if (t1 <= t0) return;
  const long long col = ((long long)blockIdx.y * 256 + threadIdx.x) * 4;
  if (col < D) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long t = t0; t < t1; ++t) {
      const float4 v = *reinterpret_cast<const float4*>(x + t * D + col);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float s = two_stage ? 1.f / (float)(t1 - t0) : 1.f;
    float* dst = table_sum + c * D + col;
    atomicAdd(dst + 0, acc.x * s);
    atomicAdd(dst + 1, acc.y * s);
    atomicAdd(dst + 2, acc.z * s);
    atomicAdd(dst + 3, acc.w * s);
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(count + c, two_stage ? 1.f : (float)(t1 - t0));
}

__global__ void __launch_bounds__(256)
class_mean_finalize_kernel(float* __restrict__ table, const float* __restrict__ count, long long n, long long D) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = count[i / D];
    table[i] = c > 0.f ? table[i] / c : 0.f;
  }
}

// out[r][d] = sum_l softmax(w_raw)[l] * nan_to_zero(ref[r][l][d])  -> bf16
__global__ void __launch_bounds__(256)
layer_weighted_sum_kernel(const float* __restrict__ ref, const float* __restrict__ w_raw, long long rows, int n_layer,
                          int D, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  __shared__ float s_w[64];
  if (threadIdx.x < 64) {
    float w = 1.f;
    if (w_raw) {
      float mx = -INFINITY;
      for (int l = 0; l < n_layer; ++l) mx = fmaxf(mx, w_raw[l]);
      float den = 0.f;
      for (int l = 0; l < n_layer; ++l) den += __expf(w_raw[l] - mx);
      w = threadIdx.x < n_layer ? __expf(w_raw[threadIdx.x] - mx) / den : 0.f;
    }
    s_w[threadIdx.x] = w;
  }
  __syncthreads();
  const long long n = rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int d = i - r * D;
    float acc = 0.f;
    for (int l = 0; l < n_layer; ++l) {
      const float w = s_w[l];
      if (w == 0.f) continue;  // softmax of -inf entries: the product with a finite value is exactly 0
      float v = ref[(r * n_layer + l) * D + d];
      if (v != v) v = 0.f;
      acc += w * v;
    }
    out[i] = __float2bfloat16(acc);
  }
}

// One block per query row, one warp per head (H <= 8, C <= 1024 codes, head dim = E / H <= 128).
__global__ void __launch_bounds__(256)
codebook_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ att, const float* __restrict__ emb,
                         int C, int E, int H, float inv_temp, float* __restrict__ out, float* __restrict__ p_out) {
  pdl_sync();
  extern __shared__ float sm[];  // [H][C] probabilities
  const int r = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= H) return;
  const int dh = E / H;
  const float* qh = q + (long long)r * E + h * dh;
  float* ph = sm + h * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) {
    const float* a = att + (long long)c * E + h * dh;
    float s = 0.f;
    for (int d = 0; d < dh; ++d) s += qh[d] * a[d];
    s *= inv_temp;
    ph[c] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float den = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float e = __expf(ph[c] - mx);
    ph[c] = e;
    den += e;
  }
  den = warp_sum(den);
  const float inv = 1.f / den;
  __syncwarp();
  for (int c = lane; c < C; c += 32) {
    const float p = ph[c] * inv;
    ph[c] = p;
    p_out[((long long)r * H + h) * C + c] = p;
  }
  __syncwarp();
  for (int d = lane; d < dh; d += 32) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += ph[c] * emb[(long long)c * E + h * dh + d];
    out[(long long)r * E + h * dh + d] = acc;
  }
}

__global__ void __launch_bounds__(256)
codebook_attn_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ q, const float* __restrict__ att,
                         const float* __restrict__ emb, const float* __restrict__ p_in, int C, int E, int H,
                         float inv_temp, float* __restrict__ dq, float* __restrict__ d_att, float* __restrict__ d_emb) {
  pdl_sync();
  extern __shared__ float sm[];  // [H][C] dS
  const int r = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= H) return;
  const int dh = E / H;
  const float* qh = q + (long long)r * E + h * dh;
  const float* gh = dout + (long long)r * E + h * dh;
  const float* p = p_in + ((long long)r * H + h) * C;
  float* ds = sm + h * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float* e = emb + (long long)c * E + h * dh;
    float dp = 0.f;
    for (int d = 0; d < dh; ++d) dp += gh[d] * e[d];
    ds[c] = dp;
    dot += p[c] * dp;
  }
  dot = warp_sum(dot);
  __syncwarp();
  for (int c = lane; c < C; c += 32) ds[c] = p[c] * (ds[c] - dot) * inv_temp;
  __syncwarp();
  for (int d = lane; d < dh; d += 32) {
    float acc = 0.f;
    const float g = gh[d], qd = qh[d];
    for (int c = 0; c < C; ++c) {
      const long long o = (long long)c * E + h * dh + d;
      acc += ds[c] * att[o];
      atomicAdd(d_att + o, ds[c] * qd);
      atomicAdd(d_emb + o, p[c] * g);
    }
    dq[(long long)r * E + h * dh + d] = acc;
  }
}

}  // namespace fs2

extern "C" {

int fs2_segment_class_accum_f32(const float* x, const int64_t* dur, const int64_t* cls, int L, int64_t T, int64_t D,
                                int n_classes, int two_stage, float* table_sum, float* count, void* stream) {
  if (L <= 0 || T <= 0) return 0;
  if (D % 4 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(table_sum) & 15))
    return fs2::set_error("segment_class_accum: D must be a multiple of 4, buffers 16-byte aligned");
  dim3 grid(L, (unsigned)((D + 1023) / 1024));
  FS2_LAUNCH((fs2::segment_class_accum_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<const long long*>(dur), reinterpret_cast<const long long*>(cls), L, T, D, n_classes,
      two_stage, table_sum, count);
  fs2::count_launch();
  return fs2::check_launch("segment_class_accum_kernel");
}

int fs2_class_mean_finalize_f32(float* table, const float* count, int n_classes, int64_t D, void* stream) {
  const long long n = (long long)n_classes * D;
  if (n <= 0) return 0;
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  FS2_LAUNCH((fs2::class_mean_finalize_kernel), (unsigned)g, 256, 0, static_cast<cudaStream_t>(stream), table, count, n, D);
  fs2::count_launch();
  return fs2::check_launch("class_mean_finalize_kernel");
}

int fs2_layer_weighted_sum_bf16(const float* ref, const float* w_raw, int64_t rows, int n_layer, int D, void* out,
                                void* stream) {
  if (rows <= 0) return 0;
  if (n_layer < 1 || n_layer > 64) return fs2::set_error("layer_weighted_sum: 1..64 layers");
  const long long n = rows * D;
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  FS2_LAUNCH((fs2::layer_weighted_sum_kernel), (unsigned)g, 256, 0, static_cast<cudaStream_t>(stream), 
      ref, w_raw, rows, n_layer, D, static_cast<__nv_bfloat16*>(out));
  fs2::count_launch();
  return fs2::check_launch("layer_weighted_sum_kernel");
}

int fs2_codebook_attn_fwd_f32(const float* q, const float* att_banks, const float* emb_banks, int rows, int C, int E,
                              int H, float inv_temp, float* out, float* p, void* stream) {
  if (rows <= 0) return 0;
  if (H < 1 || H > 8 || E % H || C < 1 || C > 1024) return fs2::set_error("codebook_attn: H <= 8, C <= 1024, E % H == 0");
  FS2_LAUNCH((fs2::codebook_attn_fwd_kernel), rows, 256, (size_t)H * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      q, att_banks, emb_banks, C, E, H, inv_temp, out, p);
  fs2::count_launch();
  return fs2::check_launch("codebook_attn_fwd_kernel");
}

int fs2_codebook_attn_bwd_f32(const float* dout, const float* q, const float* att_banks, const float* emb_banks,
                              const float* p, int rows, int C, int E, int H, float inv_temp, float* dq, float* d_att,
                              float* d_emb, void* stream) {
  if (rows <= 0) return 0;
  if (H < 1 || H > 8 || E % H || C < 1 || C > 1024) return fs2::set_error("codebook_attn: H <= 8, C <= 1024, E % H == 0");
  FS2_LAUNCH((fs2::codebook_attn_bwd_kernel), rows, 256, (size_t)H * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      dout, q, att_banks, emb_banks, p, C, E, H, inv_temp, dq, d_att, d_emb);
  fs2::count_launch();
  return fs2::check_launch("codebook_attn_bwd_kernel");
}
}
