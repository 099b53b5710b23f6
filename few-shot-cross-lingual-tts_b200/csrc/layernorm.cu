// Fused (dropout +) residual-add + LayerNorm (+ dropout) + padding-row zeroing, forward and backward.
//
// Replaces, in one pass over HBM each way:
//   transformer/SubLayers.py:54-55  `layer_norm(dropout(fc(out)) + residual)`        (drop_mode 1)
//   transformer/SubLayers.py:90-91  `layer_norm(dropout(w_2(..)) + residual)`         (drop_mode 1)
//   transformer/Layers.py:25,28     `masked_fill(mask.unsqueeze(-1), 0)`              (lens != NULL)
//   lightning/model/modules.py:222-225,234-237  `dropout(layer_norm(relu(conv)))`     (drop_mode 2)
// One warp owns one row (C = 256/512/1024 channels = 1/2/4 16-byte vectors per lane), so the
// statistics never leave registers.  HBM-bound: algorithmic bytes fwd = (2 or 3) * rows * C * 2.

#include <cstdlib>
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
  uint8_t* keep;  // dropout keep bits [rows][C/8] (bit j of byte v = channel 8v + j): written by the forward,
                  // read by the backward -- the Philox stream is drawn once per step
  // backward only
  const __nv_bfloat16* dy;
  __nv_bfloat16* dx;
  __nv_bfloat16* dres;
  float* dgamma;
  float* dbeta;
  float* dbias;  // optional: column sums of dx (bias gradient of the GEMM that produced x)
};

// (utterance, frame) of the rows a warp visits: row += stride without a division per row
struct RowCursor {
  int b, t, T, qb, qt;
  __device__ __forceinline__ RowCursor(long long row, long long stride, int T_) : T(T_) {
    b = (int)(row / T_);
    t = (int)(row - (long long)b * T_);
    qb = (int)(stride / T_);
    qt = (int)(stride - (long long)qb * T_);
  }
  __device__ __forceinline__ void next() {
    b += qb;
    t += qt;
    if (t >= T) {
      t -= T;
      ++b;
    }
  }
  __device__ __forceinline__ bool masked(const int64_t* lens) const { return lens && t >= lens[b]; }
};

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnArgs a) {
  pdl_sync();
  constexpr int C = NV * 256;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const bool drop = a.drop_mode != 0;
  const uint32_t thresh = keep_thresh_h2(a.p_drop);
  const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const float in_scale = a.drop_mode == 1 ? keep_scale : 1.f;
  const uint64_t seed = mix_seed(a.seed_dev, a.seed);
  const long long stride = (long long)gridDim.x * warps_per_block;
  long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  float gam[NV][8], bet[NV][8];  // this lane's affine parameters: loop invariant (mode 2: keep scale folded in)
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
    if (a.drop_mode == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gam[i][j] *= keep_scale;
        bet[i][j] *= keep_scale;
      }
    }
  }
  if (row >= a.rows) return;
  RowCursor cur(row, stride, a.T);
  // software pipeline: the next row's 16-byte vectors are in flight while this row is reduced
  bf16x8 nx[NV], nr[NV];
  bool nmask = cur.masked(a.lens);  // padded rows are never read
  if (!nmask) {
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
    cur.next();
    const long long nrow = row + stride;
    nmask = nrow < a.rows && cur.masked(a.lens);
    if (nrow < a.rows && !nmask) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        nx[i] = ld8(a.x + nrow * C + i * 256 + lane * 8);
        if (a.res) nr[i] = ld8(a.res + nrow * C + i * 256 + lane * 8);
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
    KeepMask km[NV];
    if (drop) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        km[i].draw(seed, ((uint64_t)row * C + i * 256 + lane * 8) >> 3, thresh);
        if (a.keep) a.keep[row * (C / 8) + i * 32 + lane] = static_cast<uint8_t>(km[i].to_byte());
      }
    }
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (a.drop_mode == 1) km[i].apply(cx[i]);
      unpack8(cx[i], v[i]);
      if (a.res) {
        float r[8];
        unpack8(cr[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = fmaf(v[i][j], in_scale, r[j]);
      } else if (a.drop_mode == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] *= in_scale;
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
        q = fmaf(d, d, q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + kLnEps);
    if (lane == 0) {
      a.mean[row] = mean;
      a.rstd[row] = rstd;
    }
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(v[i][j], rstd, nmr), gam[i][j], bet[i][j]);
      bf16x8 ov = pack8(o);
      if (a.drop_mode == 2) km[i].apply(ov);
      st8(a.y + row * C + i * 256 + lane * 8, ov);
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
  const bool drop = a.drop_mode != 0;
  const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const float in_scale = a.drop_mode == 1 ? keep_scale : 1.f;
  // mode 3: x is already the pre-norm sum v = dropout(f) / (1 - p) + res (written by the fused GEMM epilogue,
  // fs2_gemm::ln_v): nothing to rebuild on the input side, the mask / scale only apply to dx
  const bool out_mask = a.drop_mode == 1 || a.drop_mode == 3;
  const float out_scale = out_mask ? keep_scale : 1.f;
  float acc_g[NV][8], acc_b[NV][8], acc_x[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[i][j] = acc_b[i][j] = acc_x[i][j] = 0.f;

  const long long stride = (long long)gridDim.x * warps_per_block;
  long long row = (long long)blockIdx.x * warps_per_block + warp;
  float gam[NV][8];  // loop invariant (mode 2: the keep scale of dy is folded in -- see gy below)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + col);
    const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + col + 4);
    gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
    gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
  }
  if (row < a.rows) {
    RowCursor cur(row, stride, a.T);
    bf16x8 nx[NV], nd[NV], nr[NV];
    uint32_t nk[NV];
    bool nmask = cur.masked(a.lens);  // padded rows are never read
    if (!nmask) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        nx[i] = ld8(a.x + row * C + i * 256 + lane * 8);
        nd[i] = ld8(a.dy + row * C + i * 256 + lane * 8);
        if (a.res) nr[i] = ld8(a.res + row * C + i * 256 + lane * 8);
        if (drop) nk[i] = a.keep[row * (C / 8) + i * 32 + lane];
      }
    }
    for (; row < a.rows; row += stride) {
      const bool masked = nmask;
      bf16x8 cx[NV], cd[NV], cr[NV];
      uint32_t ck[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        cx[i] = nx[i];
        cd[i] = nd[i];
        cr[i] = nr[i];
        ck[i] = nk[i];
      }
      cur.next();
      const long long nrow = row + stride;
      nmask = nrow < a.rows && cur.masked(a.lens);
      if (nrow < a.rows && !nmask) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          nx[i] = ld8(a.x + nrow * C + i * 256 + lane * 8);
          nd[i] = ld8(a.dy + nrow * C + i * 256 + lane * 8);
          if (a.res) nr[i] = ld8(a.res + nrow * C + i * 256 + lane * 8);
          if (drop) nk[i] = a.keep[nrow * (C / 8) + i * 32 + lane];
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
      const float nmr = -mean * rstd;
      float xh[NV][8], gy[NV][8];
      KeepMask km[NV];
      uint32_t relu_pos[NV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (drop) km[i].from_byte(ck[i]); else km[i].all();
        if (a.relu_x) {  // sign bits of the packed bf16 pairs: x > 0 <=> not negative and not zero
          float xv0[8];
          unpack8(cx[i], xv0);
          uint32_t pos = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) pos |= (xv0[j] > 0.f ? 1u : 0u) << j;
          relu_pos[i] = pos;
        }
        if (a.drop_mode == 1) km[i].apply(cx[i]);
        if (a.drop_mode == 2) km[i].apply(cd[i]);
        float xv[8], dyv[8];
        unpack8(cx[i], xv);
        unpack8(cd[i], dyv);
        if (a.res) {
          float r[8];
          unpack8(cr[i], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = fmaf(xv[j], in_scale, r[j]);
        } else if (a.drop_mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] *= in_scale;
        }
        if (a.drop_mode == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dyv[j] *= keep_scale;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = fmaf(xv[j], rstd, nmr);
          acc_g[i][j] = fmaf(dyv[j], xh[i][j], acc_g[i][j]);
          acc_b[i][j] += dyv[j];
          gy[i][j] = dyv[j] * gam[i][j];
          s1 += gy[i][j];
          s2 = fmaf(gy[i][j], xh[i][j], s2);
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
          dxo[j] = dpre[j] * out_scale;
          if (a.relu_x && !((relu_pos[i] >> j) & 1u)) dxo[j] = 0.f;
        }
        bf16x8 dxv = pack8(dxo);
        if (out_mask) km[i].apply(dxv);
        if (a.dbias) {  // column sums of what is stored (dropped elements contribute zero)
          float st[8];
          unpack8(dxv, st);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_x[i][j] += out_mask ? st[j] : dxo[j];
        }
        st8(a.dx + row * C + col, dxv);
        if (a.dres) st8(a.dres + row * C + col, pack8(dpre));
      }
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
    for (int c = threadIdx.x; c < C && dst; c += blockDim.x) {
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
  // one wave of row-strided blocks: as many as are RESIDENT (the backward holds 116 registers -> 2 blocks of 256
  // per SM; a grid of 4 per SM ran as two waves with twice the block reductions / atomics: 36.9 -> 34.8 us at C2)
  static int slots[2][3] = {{0, 0, 0}, {0, 0, 0}};
  const int vi = C == 256 ? 0 : (C == 512 ? 1 : 2);
  if ((C == 256 || C == 512 || C == 1024) && !slots[BWD][vi]) {
    int dev = 0, sms = 148, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (BWD)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occ, C == 256 ? ln_bwd_kernel<1> : (C == 512 ? ln_bwd_kernel<2> : ln_bwd_kernel<4>), 256, 0);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occ, C == 256 ? ln_fwd_kernel<1> : (C == 512 ? ln_fwd_kernel<2> : ln_fwd_kernel<4>), 256, 0);
    slots[BWD][vi] = sms * (occ > 0 ? occ : 1);
  }
  const int grid = ln_grid(a.rows, slots[BWD][vi] > 0 ? slots[BWD][vi] : 148);
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
                    uint8_t* keep_out, void* stream) {
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
  a.keep = keep_out;
  return fs2::ln_dispatch<false>(a, C, static_cast<cudaStream_t>(stream));
}

int fs2_ln_bwd_bf16(const void* dy, const void* x, const void* res, const float* gamma,
                    const float* mean, const float* rstd, const int64_t* lens, int B, int T, int C,
                    float p_drop, int drop_mode, int relu_x, const uint8_t* keep_in,
                    void* dx, void* dres, float* dgamma, float* dbeta, float* dbias, void* stream) {
  if (p_drop > 0.f && !keep_in) return fs2::set_error("ln_bwd: p_drop > 0 needs the keep bits written by fs2_ln_fwd_bf16");
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
  a.keep = const_cast<uint8_t*>(keep_in);
  a.relu_x = relu_x;
  a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dres = static_cast<__nv_bfloat16*>(dres);
  a.dgamma = dgamma;
  a.dbeta = dbeta;
  a.dbias = dbias;
  return fs2::ln_dispatch<true>(a, C, static_cast<cudaStream_t>(stream));
}
}
