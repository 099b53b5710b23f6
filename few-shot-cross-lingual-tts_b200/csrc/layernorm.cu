// Fused (dropout +) residual-add + LayerNorm (+ dropout) + padding-row zeroing, forward and backward.
//
// Replaces, in one pass over HBM each way:
//   transformer/SubLayers.py:54-55  `layer_norm(dropout(fc(out)) + residual)`        (drop_mode 1)
//   transformer/SubLayers.py:90-91  `layer_norm(dropout(w_2(..)) + residual)`         (drop_mode 1)
//   transformer/Layers.py:25,28     `masked_fill(mask.unsqueeze(-1), 0)`              (lens != NULL)
//   lightning/model/modules.py:222-225,234-237  `dropout(layer_norm(relu(conv)))`     (drop_mode 2)
// One warp owns one row (C = 256/512/1024 channels = 1/2/4 16-byte vectors per lane), so the
// statistics never leave registers.  HBM-bound: algorithmic bytes fwd = (2 or 3) * rows * C * 2.
#include "common.h"
#include "util.cuh"

namespace fs2 {

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default (the reference never overrides it)

struct LnArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* res;
  const float* gamma;
  const float* beta;
  const int64_t* lens;
  int rows, T;
  float p_drop;
  int drop_mode;
  int relu_x;  // backward: x is a ReLU output, fold the ReLU backward (x > 0) into dx
  uint64_t seed;
  const uint64_t* seed_dev;
  __nv_bfloat16* y;
  float* mean;
  float* rstd;
  // backward only
  const __nv_bfloat16* dy;
  __nv_bfloat16* dx;
  __nv_bfloat16* dres;
  float* dgamma;
  float* dbeta;
  float* dbias;  // optional: column sums of dx (bias gradient of the GEMM that produced x)
};

__device__ __forceinline__ bool ln_row_masked(const LnArgs& a, long long row) {
  if (!a.lens) return false;
  const int r = (int)row;  // rows = B * T < 2^31 (checked on the host): 32-bit division
  const int b = r / a.T;
  return r - b * a.T >= a.lens[b];
}

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnArgs a) {
  pdl_sync();
  constexpr int C = NV * 256;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const uint32_t thresh = dropout_thresh(a.p_drop);
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint64_t seed = mix_seed(a.seed_dev, a.seed);
  const long long stride = (long long)gridDim.x * warps_per_block;
  long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  float gam[NV][8], bet[NV][8];  // this lane's affine parameters: loop invariant
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + col);
    const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(a.beta + col);
    const float4 b1 = *reinterpret_cast<const float4*>(a.beta + col + 4);
    gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
    gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
    bet[i][0] = b0.x; bet[i][1] = b0.y; bet[i][2] = b0.z; bet[i][3] = b0.w;
    bet[i][4] = b1.x; bet[i][5] = b1.y; bet[i][6] = b1.z; bet[i][7] = b1.w;
  }
  // software pipeline: the next row's 16-byte vectors are in flight while this row is reduced
  bf16x8 nx[NV], nr[NV];
  bool nmask = row < a.rows && ln_row_masked(a, row);  // padded rows are never read
  if (row < a.rows && !nmask) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      nx[i] = ld8(a.x + row * C + i * 256 + lane * 8);
      if (a.res) nr[i] = ld8(a.res + row * C + i * 256 + lane * 8);
    }
  }
  for (; row < a.rows; row += stride) {
    const bool masked = nmask;
    bf16x8 cx[NV], cr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx[i] = nx[i];
      cr[i] = nr[i];
    }
    nmask = row + stride < a.rows && ln_row_masked(a, row + stride);
    if (row + stride < a.rows && !nmask) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        nx[i] = ld8(a.x + (row + stride) * C + i * 256 + lane * 8);
        if (a.res) nr[i] = ld8(a.res + (row + stride) * C + i * 256 + lane * 8);
      }
    }
    if (masked) {  // transformer/Layers.py:25,28: the row is zero whatever the sub-layer produced
      const bf16x8 z = {};
#pragma unroll
      for (int i = 0; i < NV; ++i) st8(a.y + row * C + i * 256 + lane * 8, z);
      if (lane == 0) {
        a.mean[row] = 0.f;
        a.rstd[row] = 0.f;
      }
      continue;
    }
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      unpack8(cx[i], v[i]);
      if (a.drop_mode == 1 && thresh) {
        const uint32_t keep = dropout_keep8(seed, (uint64_t)row * C + col, thresh);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = (keep >> j) & 1u ? v[i][j] * keep_scale : 0.f;
      }
      if (a.res) {
        float r[8];
        unpack8(cr[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += r[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q += d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + kLnEps);
    if (lane == 0) {
      a.mean[row] = mean;
      a.rstd[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      float o[8];
      {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gam[i][j] + bet[i][j];
        if (a.drop_mode == 2 && thresh) {
          const uint32_t keep = dropout_keep8(seed, (uint64_t)row * C + col, thresh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (keep >> j) & 1u ? o[j] * keep_scale : 0.f;
        }
      }
      st8(a.y + row * C + col, pack8(o));
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnArgs a) {
  pdl_sync();
  constexpr int C = NV * 256;
  __shared__ float red[8][C];  // per-warp partials, reused for dgamma then dbeta
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const uint32_t thresh = dropout_thresh(a.p_drop);
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint64_t seed = mix_seed(a.seed_dev, a.seed);
  float acc_g[NV][8], acc_b[NV][8], acc_x[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[i][j] = acc_b[i][j] = acc_x[i][j] = 0.f;

  const long long stride = (long long)gridDim.x * warps_per_block;
  long long row = (long long)blockIdx.x * warps_per_block + warp;
  float gam[NV][8];  // loop invariant
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + col);
    const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + col + 4);
    gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
    gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
  }
  bf16x8 nx[NV], nd[NV], nr[NV];
  bool nmask = row < a.rows && ln_row_masked(a, row);  // padded rows are never read
  if (row < a.rows && !nmask) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      nx[i] = ld8(a.x + row * C + i * 256 + lane * 8);
      nd[i] = ld8(a.dy + row * C + i * 256 + lane * 8);
      if (a.res) nr[i] = ld8(a.res + row * C + i * 256 + lane * 8);
    }
  }
  for (; row < a.rows; row += stride) {
    const bool masked = nmask;
    bf16x8 cx[NV], cd[NV], cr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx[i] = nx[i];
      cd[i] = nd[i];
      cr[i] = nr[i];
    }
    nmask = row + stride < a.rows && ln_row_masked(a, row + stride);
    if (row + stride < a.rows && !nmask) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        nx[i] = ld8(a.x + (row + stride) * C + i * 256 + lane * 8);
        nd[i] = ld8(a.dy + (row + stride) * C + i * 256 + lane * 8);
        if (a.res) nr[i] = ld8(a.res + (row + stride) * C + i * 256 + lane * 8);
      }
    }
    if (masked) {
      const bf16x8 z = {};
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = i * 256 + lane * 8;
        st8(a.dx + row * C + col, z);
        if (a.dres) st8(a.dres + row * C + col, z);
      }
      continue;
    }
    const float mean = a.mean[row], rstd = a.rstd[row];
    float xh[NV][8], gy[NV][8];
    uint32_t keep[NV], relu_pos[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      float xv[8], dyv[8];
      unpack8(cx[i], xv);
      unpack8(cd[i], dyv);
      uint32_t pos = 0xFFu;
      if (a.relu_x) {
        pos = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) pos |= (xv[j] > 0.f ? 1u : 0u) << j;
      }
      relu_pos[i] = pos;
      keep[i] = thresh ? dropout_keep8(seed, (uint64_t)row * C + col, thresh) : 0xFFu;
      if (a.drop_mode == 1 && thresh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = (keep[i] >> j) & 1u ? xv[j] * keep_scale : 0.f;
      }
      if (a.drop_mode == 2 && thresh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dyv[j] = (keep[i] >> j) & 1u ? dyv[j] * keep_scale : 0.f;
      }
      if (a.res) {
        float r[8];
        unpack8(cr[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] += r[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xv[j] - mean) * rstd;
        acc_g[i][j] += dyv[j] * xh[i][j];
        acc_b[i][j] += dyv[j];
        gy[i][j] = dyv[j] * gam[i][j];
        s1 += gy[i][j];
        s2 += gy[i][j] * xh[i][j];
      }
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      float dpre[8], dxo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dpre[j] = rstd * (gy[i][j] - m1 - xh[i][j] * m2);
        dxo[j] = (a.drop_mode == 1 && thresh) ? ((keep[i] >> j) & 1u ? dpre[j] * keep_scale : 0.f)
                                              : dpre[j];
        if (!((relu_pos[i] >> j) & 1u)) dxo[j] = 0.f;
        acc_x[i][j] += dxo[j];
      }
      st8(a.dx + row * C + col, pack8(dxo));
      if (a.dres) st8(a.dres + row * C + col, pack8(dpre));
    }
  }
  // block reduction of the affine-parameter gradients, then one atomic per column per block
  for (int pass = 0; pass < (a.dbias ? 3 : 2); ++pass) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        red[warp][i * 256 + lane * 8 + j] = pass == 0 ? acc_g[i][j] : (pass == 1 ? acc_b[i][j] : acc_x[i][j]);
    __syncthreads();
    float* dst = pass == 0 ? a.dgamma : (pass == 1 ? a.dbeta : a.dbias);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < warps_per_block; ++w) s += red[w][c];
      atomicAdd(dst + c, s);
    }
    __syncthreads();
  }
}

static int ln_grid(int rows, int cap) {
  int g = (rows + 7) / 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

template <bool BWD>
static int ln_dispatch(const LnArgs& a, int C, cudaStream_t s) {
  if (a.rows <= 0) return 0;
  if ((long long)a.rows >= (1ll << 31) - 1) return set_error("layernorm: B * T must be < 2^31");
  if (a.p_drop < 0.f || a.p_drop >= 1.f) return set_error("layernorm: dropout p must be in [0,1)");
  const int grid = BWD ? ln_grid(a.rows, 148 * 4) : ln_grid(a.rows, 148 * 8);
  switch (C) {
    case 256:
      if (BWD) FS2_LAUNCH((ln_bwd_kernel<1>), grid, 256, 0, s, a); else FS2_LAUNCH((ln_fwd_kernel<1>), grid, 256, 0, s, a);
      break;
    case 512:
      if (BWD) FS2_LAUNCH((ln_bwd_kernel<2>), grid, 256, 0, s, a); else FS2_LAUNCH((ln_fwd_kernel<2>), grid, 256, 0, s, a);
      break;
    case 1024:
      if (BWD) FS2_LAUNCH((ln_bwd_kernel<4>), grid, 256, 0, s, a); else FS2_LAUNCH((ln_fwd_kernel<4>), grid, 256, 0, s, a);
      break;
    default:
      return set_error("layernorm: C must be 256, 512 or 1024");
  }
  count_launch();
  return check_launch(BWD ? "ln_bwd_kernel" : "ln_fwd_kernel");
}

}  // namespace fs2

extern "C" {

int fs2_ln_fwd_bf16(const void* x, const void* res, const float* gamma, const float* beta,
                    const int64_t* lens, int B, int T, int C, float p_drop, int drop_mode,
                    uint64_t seed, const uint64_t* seed_dev, void* y, float* mean, float* rstd,
                    void* stream) {
  fs2::LnArgs a{};
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.res = static_cast<const __nv_bfloat16*>(res);
  a.gamma = gamma;
  a.beta = beta;
  a.lens = lens;
  a.rows = B * T;
  a.T = T;
  a.p_drop = p_drop;
  a.drop_mode = p_drop > 0.f ? drop_mode : 0;
  a.seed = seed;
  a.seed_dev = seed_dev;
  a.y = static_cast<__nv_bfloat16*>(y);
  a.mean = mean;
  a.rstd = rstd;
  return fs2::ln_dispatch<false>(a, C, static_cast<cudaStream_t>(stream));
}

int fs2_ln_bwd_bf16(const void* dy, const void* x, const void* res, const float* gamma,
                    const float* mean, const float* rstd, const int64_t* lens, int B, int T, int C,
                    float p_drop, int drop_mode, int relu_x, uint64_t seed, const uint64_t* seed_dev,
                    void* dx, void* dres, float* dgamma, float* dbeta, float* dbias, void* stream) {
  fs2::LnArgs a{};
  a.dy = static_cast<const __nv_bfloat16*>(dy);
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.res = static_cast<const __nv_bfloat16*>(res);
  a.gamma = gamma;
  a.mean = const_cast<float*>(mean);
  a.rstd = const_cast<float*>(rstd);
  a.lens = lens;
  a.rows = B * T;
  a.T = T;
  a.p_drop = p_drop;
  a.drop_mode = p_drop > 0.f ? drop_mode : 0;
  a.seed = seed;
  a.seed_dev = seed_dev;
  a.relu_x = relu_x;
  a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dres = static_cast<__nv_bfloat16*>(dres);
  a.dgamma = dgamma;
  a.dbeta = dbeta;
  a.dbias = dbias;
  return fs2::ln_dispatch<true>(a, C, static_cast<cudaStream_t>(stream));
}
}
