// PostNet BatchNorm1d (training mode) + tanh + dropout, forward and backward, channels-last.
//
// Replaces transformer/Layers.py:129-137: for each of the five blocks
//     x = F.dropout(torch.tanh(BatchNorm1d(conv(x))), 0.5, training)      (no tanh on the last)
// and `postnet(output) + output` (lightning/model/fastspeech2m.py:145) for the last block.
// Statistics are taken over ALL B*T rows, padded frames included, exactly like nn.BatchNorm1d on
// the reference's [B, C, T] tensor; running_mean / running_var use
// momentum 0.1 and the unbiased variance; eps = 1e-5.
// Kernels: column statistics (sum, sum of squares) -> apply; backward: column reductions
// (dbeta = sum g, dgamma = sum g*yhat) -> apply.  All HBM-bound (one read of y per kernel).
#include "common.h"
#include "util.cuh"

namespace fs2 {

constexpr float kBnEps = 1e-5f;

struct BnArgs {
  const __nv_bfloat16* y;  // conv output [M][C]
  long long M;
  int C;
  float* stats;        // backward apply: [2][C] dbeta, dgamma (finalized)
  float* part;         // per-block partial sums [grid][2][C] (statistics / backward reduction)
  const float* fstats; // forward stats (read-only in apply / backward)
  const float* gamma;
  const float* beta;
  int act_tanh;
  float p_drop;
  uint64_t seed;
  const uint64_t* seed_dev;
  // forward outputs
  __nv_bfloat16* out_bf16;
  float* out_f32;
  const float* res_f32;
  // backward
  const void* dout;
  int dout_is_f32;
  __nv_bfloat16* dy;
  int rows_per_block;
  uint8_t* keep_out;       // forward (optional): dropout keep bits, one byte per 8 channels [M][C/8]
  const uint8_t* keep_in;  // backward (optional): the same bits instead of regenerating the Philox stream
};

__device__ __forceinline__ void load_vec8(const __nv_bfloat16* p, float (&f)[8]) { unpack8(ld8(p), f); }

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));  // rel. error ~2^-11: below bf16 output precision
  return y;
}

// Deterministic column reductions (used by the statistics and by the backward reduction): every thread owns
// one 8-channel column vector and a fixed strided set of rows; the `rs` row groups of a block are combined in
// row-group order through shared memory and the block writes ONE partial vector [2][C] to part[blockIdx.x];
// bn_finalize_kernel then adds the partials in block order.  No atomics anywhere: two runs give the same bits
// (the fp32 atomics this replaces made BatchNorm statistics -- and with them every downstream activation --
// differ from run to run).
__device__ __forceinline__ void block_partial_store(float* sh, const float (&s)[8], const float (&q)[8], int v, int ro,
                                                    int rs, int C, float* part_row) {
  // sh: [rs][2][C]
  if (ro < rs) {
    float* d = sh + (size_t)ro * 2 * C + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = s[j];
      d[C + j] = q[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float acc = 0.f;
    for (int g = 0; g < rs; ++g) acc += sh[(size_t)g * 2 * C + i];
    part_row[i] = acc;
  }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const BnArgs a) {
  pdl_sync();
  extern __shared__ float sh[];  // [rs][2][C]
  const int vpr = a.C / 8, rs = 256 / vpr;
  const int v = threadIdx.x % vpr, ro = threadIdx.x / vpr;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (ro < rs) {
    const long long r0 = (long long)blockIdx.x * a.rows_per_block;
    const long long r1 = min(r0 + a.rows_per_block, a.M);
    for (long long r = r0 + ro; r < r1; r += 4 * rs) {  // 4 independent 16-byte loads in flight per thread
      bf16x8 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + u * rs < r1) t[u] = ld8(a.y + (r + u * rs) * a.C + v * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * rs >= r1) break;
        float f[8];
        unpack8(t[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += f[j];
          q[j] += f[j] * f[j];
        }
      }
    }
  }
  block_partial_store(sh, s, q, v, ro, rs, a.C, a.part + (size_t)blockIdx.x * 2 * a.C);
}

// Fixed-order sum of the per-block partials: block = 32 columns (lane = column, coalesced 128-byte reads), warp w
// adds partials w, w + 8, ... and the 8 warps are combined in warp order.  Optionally (forward) updates the running
// statistics, (backward) accumulates dbeta / dgamma into the parameter gradients.
struct BnFinalArgs {
  const float* part;
  int nparts, C;
  float* out;  // [2][C]
  long long M;
  float momentum;
  float* running_mean;
  float* running_var;
  int64_t* num_batches;
  float* acc0;  // += out[0][c]   (dbeta)
  float* acc1;  // += out[1][c]   (dgamma)
};
__global__ void __launch_bounds__(256) bn_finalize_kernel(const BnFinalArgs a) {
  pdl_sync();
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;  // column of the [2C] vector
  float acc = 0.f;
  if (col < 2 * a.C)
    for (int p = w; p < a.nparts; p += 8) acc += a.part[(size_t)p * 2 * a.C + col];
  sh[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < 2 * a.C) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += sh[g][lane];
    a.out[col] = t;
    if (a.acc0 && col < a.C) a.acc0[col] += t;
    if (a.acc1 && col >= a.C) a.acc1[col - a.C] += t;
  }
  if (a.running_mean) {
    // running statistics need the sum AND the sum of squares of a channel; the block that owns sum column c adds
    // the partial square sums of c itself, in the same fixed order as the block that owns column C + c
    float accq = 0.f;
    if (col < a.C)
      for (int p = w; p < a.nparts; p += 8) accq += a.part[(size_t)p * 2 * a.C + a.C + col];
    __syncthreads();
    sh[w][lane] = accq;
    __syncthreads();
    if (w == 0 && col < a.C) {
      float q = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) q += sh[g][lane];
      const float invM = 1.f / (float)a.M;
      const float m = a.out[col] * invM;  // written by this thread above
      const float var = fmaxf(q * invM - m * m, 0.f);
      const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
      a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * m;
      a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unbiased;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches) *a.num_batches += 1;
  }
}

// Thread = one 8-channel column vector (fixed for the whole kernel) x a strided set of rows, so the
// per-channel affine parameters live in registers; 64 consecutive threads cover one 1 KiB row (C=512).
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnArgs a) {
  pdl_sync();
  const int vpr = a.C / 8, rs = 256 / vpr;
  const int v = threadIdx.x % vpr, ro = threadIdx.x / vpr;
  if (ro >= rs) return;
  const int c = v * 8;
  const float invM = 1.f / (float)a.M;
  float A[8], Bc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = a.fstats[c + j] * invM;
    const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
    const float r = rsqrtf(var + kBnEps) * a.gamma[c + j];
    A[j] = r;
    Bc[j] = a.beta[c + j] - m * r;
  }
  const uint32_t thresh = dropout_thresh(a.p_drop);
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint64_t seed = mix_seed(a.seed_dev, a.seed);
  const long long r0 = (long long)blockIdx.x * a.rows_per_block;
  const long long r1 = min(r0 + a.rows_per_block, a.M);
  for (long long r = r0 + ro; r < r1; r += 2 * rs) {
    const bool two = r + rs < r1;
    float f0[8], f1[8];
    load_vec8(a.y + r * a.C + c, f0);
    if (two) load_vec8(a.y + (r + rs) * a.C + c, f1);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float (&f)[8] = u == 0 ? f0 : f1;
      const long long rr = r + u * rs;
      const uint32_t keep = thresh ? dropout_keep8(seed, (uint64_t)rr * a.C + c, thresh) : 0xFFu;
      if (a.keep_out) a.keep_out[rr * vpr + v] = static_cast<uint8_t>(keep);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = fmaf(f[j], A[j], Bc[j]);
        if (a.act_tanh) z = fast_tanh(z);
        f[j] = (keep >> j) & 1u ? z * keep_scale : 0.f;
      }
      if (a.out_f32) {
        float* o = a.out_f32 + rr * a.C + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = f[j] + (a.res_f32 ? a.res_f32[rr * a.C + c + j] : 0.f);
      } else {
        st8(a.out_bf16 + rr * a.C + c, pack8(f));
      }
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnArgs a) {
  pdl_sync();
  extern __shared__ float s_acc[];  // [rs][2][C]: sum g, sum g*yhat per row group
  const int vpr = a.C / 8, rs = 256 / vpr;
  const int v = threadIdx.x % vpr, ro = threadIdx.x / vpr;
  float sb[8], sg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sb[j] = sg[j] = 0.f;
  if (ro < rs) {
    const int c = v * 8;
    const float invM = 1.f / (float)a.M;
    float rsd[8], sh[8], gam[8], bet[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = a.fstats[c + j] * invM;
      const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
      rsd[j] = rsqrtf(var + kBnEps);
      sh[j] = -m * rsd[j];
      gam[j] = a.gamma[c + j];
      bet[j] = a.beta[c + j];
    }
    const uint32_t thresh = dropout_thresh(a.p_drop);
    const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint64_t seed = mix_seed(a.seed_dev, a.seed);
    const long long r0 = (long long)blockIdx.x * a.rows_per_block;
    const long long r1 = min(r0 + a.rows_per_block, a.M);
    auto load_g = [&](long long r, float (&g)[8]) {
      if (a.dout_is_f32) {
        const float4* d = reinterpret_cast<const float4*>(static_cast<const float*>(a.dout) + r * a.C + c);
        const float4 d0 = d[0], d1 = d[1];
        g[0] = d0.x; g[1] = d0.y; g[2] = d0.z; g[3] = d0.w;
        g[4] = d1.x; g[5] = d1.y; g[6] = d1.z; g[7] = d1.w;
      } else {
        load_vec8(static_cast<const __nv_bfloat16*>(a.dout) + r * a.C + c, g);
      }
    };
    auto accum = [&](long long r, const float (&f)[8], const float (&g)[8], uint32_t kbits) {
      const uint32_t keep = !thresh ? 0xFFu : (a.keep_in ? kbits : dropout_keep8(seed, (uint64_t)r * a.C + c, thresh));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float yh = fmaf(f[j], rsd[j], sh[j]);
        float gg = (keep >> j) & 1u ? g[j] * keep_scale : 0.f;
        if (a.act_tanh) {
          const float t = fast_tanh(fmaf(yh, gam[j], bet[j]));
          gg *= 1.f - t * t;
        }
        sb[j] += gg;
        sg[j] += gg * yh;
      }
    };
    for (long long r = r0 + ro; r < r1; r += 2 * rs) {  // two rows per iteration: four 16-byte loads in flight
      const bool two = r + rs < r1;
      float f0[8], g0[8], f1[8], g1[8];
      uint32_t k0 = 0xFFu, k1 = 0xFFu;
      load_vec8(a.y + r * a.C + c, f0);
      load_g(r, g0);
      if (a.keep_in) k0 = a.keep_in[r * vpr + v];
      if (two) {
        load_vec8(a.y + (r + rs) * a.C + c, f1);
        load_g(r + rs, g1);
        if (a.keep_in) k1 = a.keep_in[(r + rs) * vpr + v];
      }
      accum(r, f0, g0, k0);
      if (two) accum(r + rs, f1, g1, k1);
    }
  }
  block_partial_store(s_acc, sb, sg, v, ro, rs, a.C, a.part + (size_t)blockIdx.x * 2 * a.C);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnArgs a) {
  pdl_sync();
  const int vpr = a.C / 8, rs = 256 / vpr;
  const int v = threadIdx.x % vpr, ro = threadIdx.x / vpr;
  if (ro >= rs) return;
  const int c = v * 8;
  const float invM = 1.f / (float)a.M;
  float rsd[8], sh[8], gam[8], bet[8], db[8], dg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = a.fstats[c + j] * invM;
    const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
    rsd[j] = rsqrtf(var + kBnEps);
    sh[j] = -m * rsd[j];
    gam[j] = a.gamma[c + j];
    bet[j] = a.beta[c + j];
    db[j] = a.stats[c + j] * invM;
    dg[j] = a.stats[a.C + c + j] * invM;
  }
  const uint32_t thresh = dropout_thresh(a.p_drop);
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const uint64_t seed = mix_seed(a.seed_dev, a.seed);
  const long long r0 = (long long)blockIdx.x * a.rows_per_block;
  const long long r1 = min(r0 + a.rows_per_block, a.M);
  for (long long r = r0 + ro; r < r1; r += rs) {
    float f[8], g[8], o[8];
    load_vec8(a.y + r * a.C + c, f);
    if (a.dout_is_f32) {
      const float4* d = reinterpret_cast<const float4*>(static_cast<const float*>(a.dout) + r * a.C + c);
      const float4 d0 = d[0], d1 = d[1];
      g[0] = d0.x; g[1] = d0.y; g[2] = d0.z; g[3] = d0.w;
      g[4] = d1.x; g[5] = d1.y; g[6] = d1.z; g[7] = d1.w;
    } else {
      load_vec8(static_cast<const __nv_bfloat16*>(a.dout) + r * a.C + c, g);
    }
    const uint32_t keep = !thresh ? 0xFFu : (a.keep_in ? a.keep_in[r * vpr + v]
                                                       : dropout_keep8(seed, (uint64_t)r * a.C + c, thresh));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float yh = fmaf(f[j], rsd[j], sh[j]);
      float gg = (keep >> j) & 1u ? g[j] * keep_scale : 0.f;
      if (a.act_tanh) {
        const float t = fast_tanh(fmaf(yh, gam[j], bet[j]));
        gg *= 1.f - t * t;
      }
      o[j] = gam[j] * rsd[j] * (gg - db[j] - yh * dg[j]);
    }
    st8(a.dy + r * a.C + c, pack8(o));
  }
}

static int bn_check(long long M, int C) {
  if (C % 8 || C > 2048 || C / 8 > 256) return set_error("batchnorm: C must be a multiple of 8, <= 2048");
  if (M <= 0) return set_error("batchnorm: empty batch");
  return 0;
}
static int rows_per_block_for(long long M) {
  long long rpb = (M + 148 * 8 - 1) / (148 * 8);
  if (rpb < 32) rpb = 32;
  return (int)rpb;
}
// the two reduction kernels: fewer, longer blocks (4 per SM) -> 4x fewer partial vectors to write and re-read
static int rows_per_block_reduce(long long M) {
  long long rpb = (M + 148 * 4 - 1) / (148 * 4);
  if (rpb < 32) rpb = 32;
  return (int)rpb;
}
static unsigned ew_grid(long long n_vec) {
  long long g = (n_vec + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace fs2

extern "C" {

// floats of workspace the statistics / backward reductions need for [M][C] (per-block partial sums)
int64_t fs2_bn_workspace_floats(int64_t M, int C) {
  if (M <= 0 || C <= 0) return 0;
  const int rpb = fs2::rows_per_block_reduce(M);
  return ((M + rpb - 1) / rpb) * 2 * C;
}

static int bn_finalize(const float* part, int nparts, int C, float* out, int64_t M, float momentum, float* rm,
                       float* rv, int64_t* nb, float* acc0, float* acc1, cudaStream_t s) {
  fs2::BnFinalArgs f{};
  f.part = part; f.nparts = nparts; f.C = C; f.out = out; f.M = M; f.momentum = momentum;
  f.running_mean = rm; f.running_var = rv; f.num_batches = nb; f.acc0 = acc0; f.acc1 = acc1;
  FS2_LAUNCH((fs2::bn_finalize_kernel), (2 * C + 31) / 32, 256, 0, s, f);
  fs2::count_launch();
  return fs2::check_launch("bn_finalize_kernel");
}

// stats f32 [2][C] := column sums and sums of squares of y (written, not accumulated; fixed summation order ->
// bit-reproducible).  ws: fs2_bn_workspace_floats(M, C) floats of scratch.  running_mean / running_var /
// num_batches_tracked (optional): the nn.BatchNorm1d running-statistics update (momentum, unbiased variance).
int fs2_bn_stats_bf16(const void* y, int64_t M, int C, float* ws, float* stats, float momentum, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, void* stream) {
  if (int rc = fs2::bn_check(M, C)) return rc;
  fs2::BnArgs a{};
  a.y = static_cast<const __nv_bfloat16*>(y);
  a.M = M; a.C = C; a.part = ws;
  a.rows_per_block = fs2::rows_per_block_reduce(M);
  const unsigned grid = (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block);
  const int rs = 256 / (C / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FS2_LAUNCH((fs2::bn_stats_kernel), grid, 256, (size_t)rs * 2 * C * sizeof(float), s, a);
  fs2::count_launch();
  if (int rc = fs2::check_launch("bn_stats_kernel")) return rc;
  return bn_finalize(ws, (int)grid, C, stats, M, momentum, running_mean, running_var, num_batches_tracked, nullptr,
                     nullptr, s);
}

// out = dropout(act(bn(y)));  exactly one of out_bf16 / out_f32 is non-NULL; res_f32 (optional) is
// added to the f32 output (the postnet residual).  keep_out (optional, uint8 [M][C/8]): the dropout keep bits, so
// that the backward reads 1 bit per element instead of regenerating the Philox stream in both of its passes.
int fs2_bn_apply_fwd(const void* y, const float* stats, const float* gamma, const float* beta, int64_t M,
                     int C, int act_tanh, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* out_bf16, float* out_f32, const float* res_f32, uint8_t* keep_out, void* stream) {
  if (int rc = fs2::bn_check(M, C)) return rc;
  if ((out_bf16 == nullptr) == (out_f32 == nullptr)) return fs2::set_error("bn_apply: one output");
  fs2::BnArgs a{};
  a.y = static_cast<const __nv_bfloat16*>(y);
  a.M = M; a.C = C; a.fstats = stats; a.gamma = gamma; a.beta = beta;
  a.act_tanh = act_tanh; a.p_drop = p_drop; a.seed = seed; a.seed_dev = seed_dev;
  a.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); a.out_f32 = out_f32; a.res_f32 = res_f32;
  a.keep_out = p_drop > 0.f ? keep_out : nullptr;
  a.rows_per_block = fs2::rows_per_block_for(M);
  FS2_LAUNCH((fs2::bn_apply_kernel), (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block), 256, 0, static_cast<cudaStream_t>(stream), a);
  fs2::count_launch();
  return fs2::check_launch("bn_apply_kernel");
}

// dstats f32 [2][C] (written): dstats[0][c] = dbeta, dstats[1][c] = dgamma, summed in a fixed order; dbeta_acc /
// dgamma_acc (optional, f32 [C]) additionally accumulate them (the parameter gradients).  ws as above.
// dout is bf16 [M][C] (dout_is_f32 = 0) or f32.
int fs2_bn_bwd(const void* dout, int dout_is_f32, const void* y, const float* stats, const float* gamma,
               const float* beta, int64_t M, int C, int act_tanh, float p_drop, uint64_t seed,
               const uint64_t* seed_dev, const uint8_t* keep_in, float* ws, float* dstats, float* dbeta_acc,
               float* dgamma_acc, void* dy, void* stream) {
  if (int rc = fs2::bn_check(M, C)) return rc;
  fs2::BnArgs a{};
  a.y = static_cast<const __nv_bfloat16*>(y);
  a.M = M; a.C = C; a.fstats = stats; a.gamma = gamma; a.beta = beta;
  a.act_tanh = act_tanh; a.p_drop = p_drop; a.seed = seed; a.seed_dev = seed_dev;
  a.dout = dout; a.dout_is_f32 = dout_is_f32; a.stats = dstats; a.part = ws;
  a.keep_in = p_drop > 0.f ? keep_in : nullptr;
  a.dy = static_cast<__nv_bfloat16*>(dy);
  a.rows_per_block = fs2::rows_per_block_reduce(M);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid_r = (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block);
  const int rs = 256 / (C / 8);
  FS2_LAUNCH((fs2::bn_bwd_reduce_kernel), grid_r, 256, (size_t)rs * 2 * C * sizeof(float), s, a);
  fs2::count_launch();
  if (int rc = fs2::check_launch("bn_bwd_reduce_kernel")) return rc;
  if (int rc = bn_finalize(ws, (int)grid_r, C, dstats, M, 0.f, nullptr, nullptr, nullptr, dbeta_acc, dgamma_acc, s))
    return rc;
  a.rows_per_block = fs2::rows_per_block_for(M);
  const unsigned grid = (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block);
  FS2_LAUNCH((fs2::bn_bwd_apply_kernel), grid, 256, 0, s, a);
  fs2::count_launch();
  return fs2::check_launch("bn_bwd_apply_kernel");
}
}
