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
  const uint8_t* keep_in;  // backward: the same bits (required when p_drop > 0)
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

// Row walk of one thread: rows r0 + ro + k * rs (k = 0 .. n-1) of its block's row range, one 8-channel vector each.
// Everything the loops need is a start offset and two 32-bit strides (the per-row 64-bit multiplies this replaces
// were ~9 integer instructions per ELEMENT of these otherwise HBM-bound kernels).
struct BnWalk {
  int vpr, rs, v, ro, c, n;
  long long row0;   // first row of this thread
  long long e0;     // element offset of (row0, c)
  int es;           // elements between two rows of this thread (rs * C)
  long long k0;     // keep-byte offset of (row0, v)
  int ks;           // rs * vpr
  __device__ __forceinline__ BnWalk(const BnArgs& a) {
    vpr = a.C / 8;
    rs = 256 / vpr;
    v = threadIdx.x % vpr;
    ro = threadIdx.x / vpr;
    c = v * 8;
    const long long r0 = (long long)blockIdx.x * a.rows_per_block;
    const long long r1 = min(r0 + a.rows_per_block, a.M);
    row0 = r0 + ro;
    n = (ro < rs && row0 < r1) ? (int)((r1 - row0 + rs - 1) / rs) : 0;
    e0 = row0 * a.C + c;
    es = rs * a.C;
    k0 = row0 * vpr + v;
    ks = rs * vpr;
  }
};

__global__ void __launch_bounds__(256, 3) bn_stats_kernel(const BnArgs a) {
  pdl_sync();
  extern __shared__ float sh[];  // [rs][2][C]
  const BnWalk w(a);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const __nv_bfloat16* py = a.y + w.e0;
  for (int k = 0; k < w.n; k += 4, py += 4 * w.es) {  // 4 independent 16-byte loads in flight per thread
    const int nr = w.n - k;
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nr) t[u] = ld8(py + u * w.es);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u >= nr) break;
      float f[8];
      unpack8(t[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
  }
  block_partial_store(sh, s, q, w.v, w.ro, w.rs, a.C, a.part + (size_t)blockIdx.x * 2 * a.C);
}

// Fixed-order sum of the per-block partials: block = 32 columns (lane = column, coalesced 128-byte reads) x 32 warps;
// warp w adds partials w, w + 32, ... with eight independent loads in flight (the sum of ~600 partials used to be a
// serial chain of dependent L2 round trips: 21-34 us for 2.4 MB) and the 32 warps are combined in warp order.
// Optionally (forward) updates the running statistics, (backward) accumulates dbeta / dgamma into the parameter
// gradients.
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
__device__ __forceinline__ float bn_partial_column_sum(const float* part, int nparts, int ld, int col, int w) {
  float acc = 0.f;
  int p = w;
  for (; p + 7 * 32 < nparts; p += 8 * 32) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = part[(size_t)(p + 32 * u) * ld + col];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += t[u];
  }
  float t[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) t[u] = p + 32 * u < nparts ? part[(size_t)(p + 32 * u) * ld + col] : 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += t[u];
  return acc;
}
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const BnFinalArgs a) {
  pdl_sync();
  __shared__ float sh[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;  // column of the [2C] vector
  sh[w][lane] = col < 2 * a.C ? bn_partial_column_sum(a.part, a.nparts, 2 * a.C, col, w) : 0.f;
  __syncthreads();
  float t = 0.f;
  if (w == 0 && col < 2 * a.C) {
#pragma unroll
    for (int g = 0; g < 32; ++g) t += sh[g][lane];
    a.out[col] = t;
    if (a.acc0 && col < a.C) a.acc0[col] += t;
    if (a.acc1 && col >= a.C) a.acc1[col - a.C] += t;
  }
  if (a.running_mean) {
    // running statistics need the sum AND the sum of squares of a channel; the block that owns sum column c adds
    // the partial square sums of c itself, in the same fixed order as the block that owns column C + c
    __syncthreads();
    sh[w][lane] = col < a.C ? bn_partial_column_sum(a.part, a.nparts, 2 * a.C, a.C + col, w) : 0.f;
    __syncthreads();
    if (w == 0 && col < a.C) {
      float q = 0.f;
#pragma unroll
      for (int g = 0; g < 32; ++g) q += sh[g][lane];
      const float invM = 1.f / (float)a.M;
      const float m = t * invM;
      const float var = fmaxf(q * invM - m * m, 0.f);
      const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
      a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * m;
      a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unbiased;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches) *a.num_batches += 1;
  }
}

// Dropout keep bits of the FOUR rows r, r + rs, r + 2 rs, r + 3 rs of one 8-channel vector (byte u = row u).
// p == 0.5 (the PostNet's rate) needs one random bit per element: one Philox call serves four iterations of the
// row loop (128 elements) instead of one call per 8 elements; any other rate compares 16 random bits per element.
template <int MODE>  // 0: no dropout, 1: p == 0.5 (one bit per element), 2: any other rate
struct BnKeepGen {
  uint64_t seed;
  uint32_t thresh;
  uint4 rnd;
  // iteration `it` of the walk: rows row, row + rs, ... (nr of them, at most 4)
  __device__ __forceinline__ uint32_t word(int it, long long row, int nr, const BnWalk& w, int C) {
    if (MODE == 0) return 0xFFFFFFFFu;
    if (MODE == 1) {
      if ((it & 3) == 0) rnd = philox4x32(seed, ((uint64_t)row * w.vpr + w.v) ^ 0x8000000000000000ull);
      const int k = it & 3;
      return k == 0 ? rnd.x : (k == 1 ? rnd.y : (k == 2 ? rnd.z : rnd.w));
    }
    uint32_t kw = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nr) kw |= dropout_keep8(seed, (uint64_t)(row + u * w.rs) * C + w.c, thresh) << (8 * u);
    return kw;
  }
};

// Thread = one 8-channel column vector (fixed for the whole kernel) x a strided set of rows, so the
// per-channel affine parameters live in registers; 64 consecutive threads cover one 1 KiB row (C=512).
// Four rows per iteration: four independent 16-byte loads in flight per thread.
template <int MODE>
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const BnArgs a) {
  pdl_sync();
  const BnWalk w(a);
  if (w.n == 0) return;
  const int c = w.c;
  const float invM = 1.f / (float)a.M;
  const float keep_scale = MODE ? 1.f / (1.f - a.p_drop) : 1.f;
  float A[8], Bc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = a.fstats[c + j] * invM;
    const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
    const float r = rsqrtf(var + kBnEps) * a.gamma[c + j];
    A[j] = r;
    Bc[j] = a.beta[c + j] - m * r;
    if (!a.act_tanh) {  // no activation between the affine map and the dropout: fold the keep scale in
      A[j] *= keep_scale;
      Bc[j] *= keep_scale;
    }
  }
  BnKeepGen<MODE> kg;
  kg.seed = mix_seed(a.seed_dev, a.seed);
  kg.thresh = dropout_thresh(a.p_drop);
  const __nv_bfloat16* py = a.y + w.e0;
  long long eo = w.e0;  // output element offset
  uint8_t* pk = a.keep_out ? a.keep_out + w.k0 : nullptr;
  long long row = w.row0;
  int it = 0;
  for (int k = 0; k < w.n; k += 4, ++it, py += 4 * w.es, eo += 4 * (long long)w.es, row += 4 * w.rs) {
    const int nr = w.n - k;
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nr) t[u] = ld8(py + u * w.es);
    const uint32_t kw = kg.word(it, row, nr, w, a.C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u >= nr) break;
      float f[8];
      unpack8(t[u], f);
      const uint32_t keep = (kw >> (8 * u)) & 0xFFu;
      if (MODE && pk) pk[(k + u) * w.ks] = static_cast<uint8_t>(keep);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = fmaf(f[j], A[j], Bc[j]);
        if (a.act_tanh) z = fast_tanh(z) * keep_scale;
        f[j] = (!MODE || ((keep >> j) & 1u)) ? z : 0.f;
      }
      const long long e = eo + (long long)u * w.es;
      if (a.out_f32) {
        float4* o = reinterpret_cast<float4*>(a.out_f32 + e);
        float4 o0 = make_float4(f[0], f[1], f[2], f[3]), o1 = make_float4(f[4], f[5], f[6], f[7]);
        if (a.res_f32) {
          const float4* rp = reinterpret_cast<const float4*>(a.res_f32 + e);
          const float4 q0 = rp[0], q1 = rp[1];
          o0.x += q0.x; o0.y += q0.y; o0.z += q0.z; o0.w += q0.w;
          o1.x += q1.x; o1.y += q1.y; o1.z += q1.z; o1.w += q1.w;
        }
        o[0] = o0;
        o[1] = o1;
      } else {
        st8(a.out_bf16 + e, pack8(f));
      }
    }
  }
}

// gradient vector of one row, as loaded (bf16: one 16-byte vector, f32: two) and unpacked at the point of use
template <bool F32G>
struct BnGrad {
  float4 lo, hi;
  __device__ __forceinline__ void load(const void* dout, long long e) {  // e: element offset
    if (F32G) {
      const float4* d = reinterpret_cast<const float4*>(static_cast<const float*>(dout) + e);
      lo = d[0];
      hi = d[1];
    } else {
      lo = *reinterpret_cast<const float4*>(static_cast<const __nv_bfloat16*>(dout) + e);
    }
  }
  __device__ __forceinline__ void get(float (&g)[8]) const {
    if (F32G) {
      g[0] = lo.x; g[1] = lo.y; g[2] = lo.z; g[3] = lo.w;
      g[4] = hi.x; g[5] = hi.y; g[6] = hi.z; g[7] = hi.w;
    } else {
      unpack8(*reinterpret_cast<const bf16x8*>(&lo), g);
    }
  }
};

// Backward, pass 1: dbeta = sum gg, dgamma = sum gg * yhat with gg = dropout'(tanh'(.)) * dout and
// yhat = (y - mean) * rstd.  The loop accumulates S0 = sum gg' and S1 = sum gg' * y on the RAW conv output
// (gg' = gg / keep_scale: two constants per channel in registers instead of six); the block epilogue turns them into
// keep_scale * S0 and keep_scale * rstd * (S1 - mean * S0).  The dropout keep bits are READ (1 bit / element, written
// by the forward), the tanh is recomputed from y.
template <bool F32G>
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_kernel(const BnArgs a) {
  pdl_sync();
  extern __shared__ float s_acc[];  // [rs][2][C]
  const BnWalk w(a);
  const int v = w.v, ro = w.ro, rs = w.rs;
  const float invM = 1.f / (float)a.M;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  if (w.n > 0) {
    const int c = w.c;
    float A1[8], B1[8];  // tanh argument = y * A1 + B1
    if (a.act_tanh) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = a.fstats[c + j] * invM;
        const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
        A1[j] = rsqrtf(var + kBnEps) * a.gamma[c + j];
        B1[j] = a.beta[c + j] - m * A1[j];
      }
    }
    const __nv_bfloat16* py = a.y + w.e0;
    long long eg = w.e0;
    const uint8_t* pk = a.keep_in ? a.keep_in + w.k0 : nullptr;
    for (int k = 0; k < w.n; k += 4, py += 4 * w.es, eg += 4 * (long long)w.es) {  // eight 16-byte loads in flight
      const int nr = w.n - k;
      bf16x8 ty[4];
      BnGrad<F32G> tg[4];
      uint32_t kb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        kb[u] = 0xFFu;
        if (u < nr) {
          ty[u] = ld8(py + u * w.es);
          tg[u].load(a.dout, eg + (long long)u * w.es);
          if (pk) kb[u] = pk[(k + u) * w.ks];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u >= nr) break;
        float f[8], g[8];
        unpack8(ty[u], f);
        tg[u].get(g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = (kb[u] >> j) & 1u ? g[j] : 0.f;
          if (a.act_tanh) {
            const float t = fast_tanh(fmaf(f[j], A1[j], B1[j]));
            gg = fmaf(-t * t, gg, gg);
          }
          s0[j] += gg;
          s1[j] = fmaf(gg, f[j], s1[j]);
        }
      }
    }
  }
  // row groups of the block in fixed order, then (S0, S1) -> (dbeta, dgamma) contributions
  if (ro < rs) {
    float* d = s_acc + (size_t)ro * 2 * a.C + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = s0[j];
      d[a.C + j] = s1[j];
    }
  }
  __syncthreads();
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  float* part_row = a.part + (size_t)blockIdx.x * 2 * a.C;
  for (int i = threadIdx.x; i < a.C; i += 256) {
    float t0 = 0.f, t1 = 0.f;
    for (int g = 0; g < rs; ++g) {
      t0 += s_acc[(size_t)g * 2 * a.C + i];
      t1 += s_acc[(size_t)g * 2 * a.C + a.C + i];
    }
    const float m = a.fstats[i] * invM;
    const float var = fmaxf(a.fstats[a.C + i] * invM - m * m, 0.f);
    part_row[i] = keep_scale * t0;
    part_row[a.C + i] = keep_scale * rsqrtf(var + kBnEps) * (t1 - m * t0);
  }
}

// Backward, pass 2: dy = gamma * rstd * (gg - dbeta/M - yhat * dgamma/M), written as
//   dy = A1 * gg + K2 * y + K1   (A1 = gamma*rstd; K1, K2 fold mean, rstd, dbeta/M and dgamma/M)
// so that four constants per channel stay in registers; two rows per iteration.
template <bool F32G>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_kernel(const BnArgs a) {
  pdl_sync();
  const BnWalk w(a);
  if (w.n == 0) return;
  const int c = w.c;
  const float invM = 1.f / (float)a.M;
  const float keep_scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  float A1[8], B1[8], K1[8], K2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = a.fstats[c + j] * invM;
    const float var = fmaxf(a.fstats[a.C + c + j] * invM - m * m, 0.f);
    const float rsd = rsqrtf(var + kBnEps), sh = -m * rsd;
    const float db = a.stats[c + j] * invM, dg = a.stats[a.C + c + j] * invM;
    A1[j] = rsd * a.gamma[c + j];
    B1[j] = a.beta[c + j] + sh * a.gamma[c + j];
    K1[j] = -A1[j] * (db + dg * sh);
    K2[j] = -A1[j] * dg * rsd;
  }
  const __nv_bfloat16* py = a.y + w.e0;
  __nv_bfloat16* pd = a.dy + w.e0;
  long long eg = w.e0;
  const uint8_t* pk = a.keep_in ? a.keep_in + w.k0 : nullptr;
  for (int k = 0; k < w.n; k += 2, py += 2 * w.es, pd += 2 * w.es, eg += 2 * (long long)w.es) {
    const int nr = w.n - k;
    bf16x8 ty[2];
    BnGrad<F32G> tg[2];
    uint32_t kb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      kb[u] = 0xFFu;
      if (u < nr) {
        ty[u] = ld8(py + u * w.es);
        tg[u].load(a.dout, eg + (long long)u * w.es);
        if (pk) kb[u] = pk[(k + u) * w.ks];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u >= nr) break;
      float f[8], g[8], o[8];
      unpack8(ty[u], f);
      tg[u].get(g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = (kb[u] >> j) & 1u ? g[j] * keep_scale : 0.f;
        if (a.act_tanh) {
          const float t = fast_tanh(fmaf(f[j], A1[j], B1[j]));
          gg = fmaf(-t * t, gg, gg);
        }
        o[j] = fmaf(A1[j], gg, fmaf(K2[j], f[j], K1[j]));
      }
      st8(pd + u * w.es, pack8(o));
    }
  }
}

static int bn_check(long long M, int C) {
  if (C % 8 || C > 2048 || C / 8 > 256) return set_error("batchnorm: C must be a multiple of 8, <= 2048");
  if (M <= 0) return set_error("batchnorm: empty batch");
  return 0;
}
// All streaming kernels run ONE wave of 3 blocks per SM (their __launch_bounds__): every block streams the same
// number of rows, no partial last wave, and the reductions write 3 x #SMs partial vectors.
static int bn_blocks() {
  static int n = 0;
  if (!n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    n = 3 * sms;
  }
  return n;
}
static int rows_per_block_for(long long M) {
  long long rpb = (M + bn_blocks() - 1) / bn_blocks();
  if (rpb < 32) rpb = 32;
  return (int)rpb;
}
static int rows_per_block_reduce(long long M) { return rows_per_block_for(M); }
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
  FS2_LAUNCH((fs2::bn_finalize_kernel), (2 * C + 31) / 32, 1024, 0, s, f);
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
  const unsigned grid = (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!(p_drop > 0.f)) FS2_LAUNCH((fs2::bn_apply_kernel<0>), grid, 256, 0, s, a);
  else if (p_drop == 0.5f) FS2_LAUNCH((fs2::bn_apply_kernel<1>), grid, 256, 0, s, a);
  else FS2_LAUNCH((fs2::bn_apply_kernel<2>), grid, 256, 0, s, a);
  fs2::count_launch();
  return fs2::check_launch("bn_apply_kernel");
}

// dstats f32 [2][C] (written): dstats[0][c] = dbeta, dstats[1][c] = dgamma, summed in a fixed order; dbeta_acc /
// dgamma_acc (optional, f32 [C]) additionally accumulate them (the parameter gradients).  ws as above.
// keep_in: the keep bits fs2_bn_apply_fwd wrote (required when p_drop > 0; seed / seed_dev are unused).
// dout is bf16 [M][C] (dout_is_f32 = 0) or f32.
int fs2_bn_bwd(const void* dout, int dout_is_f32, const void* y, const float* stats, const float* gamma,
               const float* beta, int64_t M, int C, int act_tanh, float p_drop, uint64_t seed,
               const uint64_t* seed_dev, const uint8_t* keep_in, float* ws, float* dstats, float* dbeta_acc,
               float* dgamma_acc, void* dy, void* stream) {
  if (int rc = fs2::bn_check(M, C)) return rc;
  if (p_drop > 0.f && !keep_in) return fs2::set_error("bn_bwd: p_drop > 0 needs the keep bits written by fs2_bn_apply_fwd");
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
  if (dout_is_f32) FS2_LAUNCH((fs2::bn_bwd_reduce_kernel<true>), grid_r, 256, (size_t)rs * 2 * C * sizeof(float), s, a);
  else FS2_LAUNCH((fs2::bn_bwd_reduce_kernel<false>), grid_r, 256, (size_t)rs * 2 * C * sizeof(float), s, a);
  fs2::count_launch();
  if (int rc = fs2::check_launch("bn_bwd_reduce_kernel")) return rc;
  if (int rc = bn_finalize(ws, (int)grid_r, C, dstats, M, 0.f, nullptr, nullptr, nullptr, dbeta_acc, dgamma_acc, s))
    return rc;
  a.rows_per_block = fs2::rows_per_block_for(M);
  const unsigned grid = (unsigned)((M + a.rows_per_block - 1) / a.rows_per_block);
  if (dout_is_f32) FS2_LAUNCH((fs2::bn_bwd_apply_kernel<true>), grid, 256, 0, s, a);
  else FS2_LAUNCH((fs2::bn_bwd_apply_kernel<false>), grid, 256, 0, s, a);
  fs2::count_launch();
  return fs2::check_launch("bn_bwd_apply_kernel");
}
}
