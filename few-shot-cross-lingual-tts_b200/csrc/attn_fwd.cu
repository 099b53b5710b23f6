// Fused key-padding-masked softmax attention, forward, for sm_100a (d_k = 128).
//
// Replaces transformer/Modules.py:14-25 + the head split / merge of transformer/SubLayers.py:39-52:
//     attn = softmax(mask(Q K^T / sqrt(d_k))) ; out = attn V
// without ever writing the [H*B, T, T] score / probability tensors to HBM.  Q, K, V are read straight
// out of the fused projection buffer [B][T][3*H*dk] by TMA (head = a column offset), the two GEMMs
// run on tcgen05 with S and O accumulators in TMEM, P goes registers -> swizzled smem -> tcgen05.
//
// One CTA = one (batch*head z, 128-query tile).  Exact two-pass softmax:
//   pass A: S_j = Q K_j^T for every 128-key tile j, row max m                (no exp, no P)
//   pass B: S_j again, P_j = exp2(S_j*c - m*c) (masked keys -> 0), l += rowsum, O += P_j V_j
// (QK^T is recomputed instead of rescaling O in TMEM: +50 % of the cheap GEMM, no correction step and
//  no dependence of the PV pipeline on the running max).  Epilogue: O / l -> bf16, and
//  lse2 = m*c + log2(l) per row for the backward kernels (+inf for padded / empty rows => P = 0).
//
// Warp roles (320 threads): warps 0-7 softmax + epilogue (thread = query row = TMEM lane; warps w and
// w+4 share the 32 rows of lane quarter w%4 and split the key / output columns in halves, so every SM
// sub-partition always has two softmax warps to interleave), warp 8 TMA producer, warp 9 MMA issuer
// (+ TMEM alloc).  Row max and row sum are combined across the two halves once per pass.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace fs2 {

namespace af {
constexpr int DK = 128, BQ = 128, BKV = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KiB: two [128 x 64] bf16 swizzle-128B blocks
constexpr int BLK = 16384;                 // one [128 rows x 64 cols] block
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;      // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;  // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
constexpr int OFF_RED = OFF_P + 2 * TILE_BYTES;  // P is double buffered  // [2 halves][128 rows] f32 exchange of max / sum
constexpr int OFF_BAR = OFF_RED + 2 * 128 * 4;
constexpr int NUM_BARS = 18;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_FULL = 13,
       P_EMPTY = 15, O_FULL = 17 };
}  // namespace af

struct AttnFwdP {
  const int64_t* lens;
  const int32_t* sched;  // optional work order (fs2_attn_schedule): entry = (z << 8) | tile
  int B, T, H, nq, nkv;
  float scale_log2;
  __nv_bfloat16* out;  // [B][T][H*dk]
  float* lse2;         // [B*H][T]
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ AttnFwdP p) {
  pdl_sync();
  using namespace af;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int qt = blockIdx.x % p.nq, z = blockIdx.x / p.nq;
  if (p.sched) {  // longest utterances first (the CTAs of a ragged batch differ 8x in work)
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    qt = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int q0 = qt * BQ;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (q0 >= len) {  // a query tile of padded frames only (CTA-uniform): out = 0, lse2 = +inf, no MMAs
    for (int i = threadIdx.x; i < BQ * 16; i += blockDim.x) {
      const int r = i >> 4, c = i & 15;
      if (q0 + r < p.T)
        *reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + q0 + r) * HD + h * DK + c * 8) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    for (int r = threadIdx.x; r < BQ; r += blockDim.x)
      if (q0 + r < p.T) p.lse2[(long long)z * p.T + q0 + r] = INFINITY;
    return;
  }
  const int n = min(p.nkv, (len + BKV - 1) / BKV);  // key tiles that hold at least one valid key

  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(K_FULL + s), 1);
      mbar_init(bar(K_EMPTY + s), 1);
      mbar_init(bar(V_FULL + s), 1);
      mbar_init(bar(V_EMPTY + s), 1);
      mbar_init(bar(S_FULL + s), 1);
      mbar_init(bar(S_EMPTY + s), 8);
      mbar_init(bar(P_FULL + s), 8);
      mbar_init(bar(P_EMPTY + s), 1);
    }
    mbar_init(bar(O_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256;

  if (warp == 8) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(Q_FULL), TILE_BYTES);
      for (int kb = 0; kb < 2; ++kb)
        tma_load_3d(sbase + OFF_Q + kb * BLK, &tmQK, bar(Q_FULL), h * DK + kb * 64, q0, b);
      for (int u = 0; u < 2 * n; ++u) {
        const int j = u % n, s = u & 1;
        mbar_wait(bar(K_EMPTY + s), ((u >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(bar(K_FULL + s), TILE_BYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_3d(sbase + OFF_K + s * TILE_BYTES + kb * BLK, &tmQK, bar(K_FULL + s),
                      HD + h * DK + kb * 64, j * BKV, b);
        if (u >= n) {
          const int vs = j & 1;
          mbar_wait(bar(V_EMPTY + vs), ((j >> 1) & 1) ^ 1u);
          mbar_arrive_expect_tx(bar(V_FULL + vs), TILE_BYTES);
          for (int dh = 0; dh < 2; ++dh)
            for (int kh = 0; kh < 2; ++kh)
              tma_load_3d(sbase + OFF_V + vs * TILE_BYTES + dh * BLK + kh * 8192, &tmV, bar(V_FULL + vs),
                          2 * HD + h * DK + dh * 64, j * BKV + kh * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S = Q K^T   (both K-major)
      const uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);  // O += P V    (V is MN-major)
      auto issue_s = [&](int u) {
        const int s = u & 1;
        mbar_wait(bar(K_FULL + s), (u >> 1) & 1);
        mbar_wait(bar(S_EMPTY + s), ((u >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sq = sbase + OFF_Q, sk = sbase + OFF_K + s * TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = (i >> 2) * BLK + (i & 3) * 32;
          umma_f16(tS0 + s * 128, make_smem_desc(sq + off, 16, 1024), make_smem_desc(sk + off, 16, 1024),
                   idesc_s, i > 0);
        }
        umma_commit(bar(K_EMPTY + s));
        umma_commit(bar(S_FULL + s));
      };
      mbar_wait(bar(Q_FULL), 0);
      issue_s(0);
      for (int u = 0; u < 2 * n; ++u) {
        if (u + 1 < 2 * n) issue_s(u + 1);
        if (u >= n) {
          const int j = u - n, vs = j & 1;
          mbar_wait(bar(P_FULL + (j & 1)), (j >> 1) & 1);
          mbar_wait(bar(V_FULL + vs), (j >> 1) & 1);
          tc_fence_after();
          const uint32_t sp = sbase + OFF_P + (j & 1) * TILE_BYTES, sv = sbase + OFF_V + vs * TILE_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // A = P: K-major over keys (two 64-key blocks); B = V: MN-major, 16 key rows per step
            const uint64_t ad = make_smem_desc(sp + (i >> 2) * BLK + (i & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(sv + i * 2048, BLK, 1024);
            umma_f16(tO, ad, bd, idesc_o, (j > 0 || i > 0) ? 1u : 0u);
          }
          umma_commit(bar(V_EMPTY + vs));
          umma_commit(bar(P_EMPTY + (j & 1)));
          if (j == n - 1) umma_commit(bar(O_FULL));
        }
      }
    }
  } else {
    // ============================== softmax + epilogue warps ==============================
    const int q4 = warp & 3, half = warp >> 2;  // TMEM lane quarter, column half
    const int row = q4 * 32 + lane;
    const int q = q0 + row;
    const bool row_valid = q < len;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    uint8_t* sp_row0 = sgen + OFF_P + half * BLK + row * 128;  // this half's 64-key block of the P tile
    float* s_red = reinterpret_cast<float*>(sgen + OFF_RED);
    const int sw = row & 7;
    float mx = -INFINITY, m2 = 0.f, l = 0.f;
    for (int u = 0; u < 2 * n; ++u) {
      const int j = u % n, s = u & 1;
      const bool pass_b = u >= n;
      if (u == n) {  // combine the row maxima of the two column halves (once per CTA)
        s_red[half * 128 + row] = mx;
        softmax_bar_sync();
        const float m = fmaxf(s_red[row], s_red[128 + row]);
        m2 = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        softmax_bar_sync();
      }
      mbar_wait(bar(S_FULL + s), (u >> 1) & 1);
      tc_fence_after();
      float pv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS0 + s * 128 + lane_base + half * 64 + c * 32, v);
        tmem_ld_wait();
        const int k0 = j * BKV + half * 64 + c * 32;
        if (!pass_b) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (k0 + i < len) ? __uint_as_float(v[i]) : -INFINITY);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m2));
            pv[c * 32 + i] = (row_valid && k0 + i < len) ? e : 0.f;
            l += pv[c * 32 + i];
          }
        }
      }
      if (pass_b) {
        const int jb = u - n;
        uint8_t* sp_row = sp_row0 + (jb & 1) * TILE_BYTES;
        mbar_wait(bar(P_EMPTY + (jb & 1)), ((jb >> 1) & 1) ^ 1u);  // PV MMA of tile jb-2 has consumed this buffer
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(pv[g * 8 + 2 * t], pv[g * 8 + 2 * t + 1]);
            w[t] = *reinterpret_cast<uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(sp_row + ((g ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
      if (pass_b) fence_proxy_async_smem();  // P (generic-proxy stores) must be visible to the MMA
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(S_EMPTY + s));
        if (pass_b) mbar_arrive(bar(P_FULL + ((u - n) & 1)));
      }
    }
    // ---- combine the row sums, then epilogue: O / l -> bf16 -> staging (the P tile is free once the
    //      last PV MMA has retired) -> coalesced global store; each half stores 64 output columns
    s_red[half * 128 + row] = l;
    softmax_bar_sync();
    l = s_red[row] + s_red[128 + row];
    mbar_wait(bar(O_FULL), 0);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    if (half == 0 && q < p.T) p.lse2[(long long)z * p.T + q] = l > 0.f ? m2 + log2f(l) : INFINITY;
    uint8_t* stg = sgen + OFF_P + warp * 4096;
    uint8_t* my = stg + lane * 128;
    const int lsw = lane & 7;
    {
      float f[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + lane_base + half * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[c * 32 + i] = __uint_as_float(v[i]) * inv;
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat162 b2 = __floats2bfloat162_rn(f[ch * 8 + 2 * t], f[ch * 8 + 2 * t + 1]);
          w[t] = *reinterpret_cast<uint32_t*>(&b2);
        }
        *reinterpret_cast<uint4*>(my + ((ch ^ lsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), ch = lane & 7;
        const int gq = q0 + q4 * 32 + r;
        if (gq < p.T) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + gq) * HD + h * DK + half * 64 + ch * 8) = val;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// Forward, second generation.  Same exact two-pass softmax, restructured around what bounds these kernels on B200
// (measured: barrier round trips and single-thread issue latency, not FLOPs -- profiles/README.md):
//   * P never touches shared memory: the softmax warps write it (bf16) over the S accumulator columns they have
//     just read (tcgen05.st) and the PV MMA takes its A operand from tensor memory;
//   * two issuer warps (QK^T / PV) so that neither waits behind the other's barriers, warp-uniform issue loops
//     with elect.sync and precomputed descriptors;
//   * 3 S stages in TMEM and a 3-stage K ring (the 64 KiB that P occupied), so the QK^T MMAs run two tiles ahead.
// ================================================================================================
namespace af2 {
constexpr int DK = 128, BQ = 128, BKV = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KiB: two [128 x 64] bf16 swizzle-128B blocks
constexpr int BLK = 16384;                 // one [128 rows x 64 cols] block
constexpr int KS = 3, VS = 2, SS = 3;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;
constexpr int OFF_V = OFF_K + KS * TILE_BYTES;
constexpr int OFF_STG = OFF_K;                    // epilogue staging: the K ring is idle by then
constexpr int OFF_RED = OFF_V + VS * TILE_BYTES;  // [2 halves][128 rows] f32 exchange of max / sum
constexpr int OFF_BAR = OFF_RED + 2 * 128 * 4;
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + KS, V_FULL = K_EMPTY + KS, V_EMPTY = V_FULL + VS,
       S_FULL = V_EMPTY + VS, S_EMPTY = S_FULL + SS, P_FULL = S_EMPTY + SS, O_FULL = P_FULL + SS, NUM_BARS = O_FULL + 1 };
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
}  // namespace af2

__global__ void __launch_bounds__(352, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ AttnFwdP p) {
  pdl_sync();
  using namespace af2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int qt = blockIdx.x % p.nq, z = blockIdx.x / p.nq;
  if (p.sched) {  // longest utterances first (the CTAs of a ragged batch differ 8x in work)
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    qt = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int q0 = qt * BQ;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (q0 >= len) {  // a query tile of padded frames only (CTA-uniform): out = 0, lse2 = +inf, no MMAs
    for (int i = threadIdx.x; i < BQ * 16; i += blockDim.x) {
      const int r = i >> 4, c = i & 15;
      if (q0 + r < p.T)
        *reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + q0 + r) * HD + h * DK + c * 8) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    for (int r = threadIdx.x; r < BQ; r += blockDim.x)
      if (q0 + r < p.T) p.lse2[(long long)z * p.T + q0 + r] = INFINITY;
    return;
  }
  const int n = min(p.nkv, (len + BKV - 1) / BKV);  // key tiles that hold at least one valid key

  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(bar(K_FULL + s), 1);
      mbar_init(bar(K_EMPTY + s), 1);
    }
    for (int s = 0; s < VS; ++s) {
      mbar_init(bar(V_FULL + s), 1);
      mbar_init(bar(V_EMPTY + s), 1);
    }
    for (int s = 0; s < SS; ++s) {
      mbar_init(bar(S_FULL + s), 1);
      mbar_init(bar(S_EMPTY + s), 1);  // pass A: plain arrive of the PV warp; pass B: commit after the PV MMAs
      mbar_init(bar(P_FULL + s), 8);
    }
    mbar_init(bar(O_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  const uint32_t tS0 = tmem_base, tO = tmem_base + SS * 128;

  if (warp == 8) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(Q_FULL), TILE_BYTES);
      for (int kb = 0; kb < 2; ++kb)
        tma_load_3d(sbase + OFF_Q + kb * BLK, &tmQK, bar(Q_FULL), h * DK + kb * 64, q0, b);
      for (int u = 0; u < 2 * n; ++u) {
        const int j = u % n, s = u % KS;
        mbar_wait(bar(K_EMPTY + s), ((u / KS) & 1) ^ 1u);
        mbar_arrive_expect_tx(bar(K_FULL + s), TILE_BYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_3d(sbase + OFF_K + s * TILE_BYTES + kb * BLK, &tmQK, bar(K_FULL + s),
                      HD + h * DK + kb * 64, j * BKV, b);
        if (u >= n) {
          const int vs = j % VS;
          mbar_wait(bar(V_EMPTY + vs), ((j / VS) & 1) ^ 1u);
          mbar_arrive_expect_tx(bar(V_FULL + vs), TILE_BYTES);
          for (int dh = 0; dh < 2; ++dh)
            for (int kh = 0; kh < 2; ++kh)
              tma_load_3d(sbase + OFF_V + vs * TILE_BYTES + dh * BLK + kh * 8192, &tmV, bar(V_FULL + vs),
                          2 * HD + h * DK + dh * 64, j * BKV + kh * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    // ============================== issuer 1: S = Q K^T ==============================
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // both K-major
    const uint64_t dQ0 = make_smem_desc(sbase + OFF_Q, 16, 1024), dK0 = make_smem_desc(sbase + OFF_K, 16, 1024);
    mbar_wait(bar(Q_FULL), 0);
    for (int u = 0; u < 2 * n; ++u) {
      const int s = u % KS, s3 = u % SS;
      mbar_wait(bar(K_FULL + s), (u / KS) & 1);
      mbar_wait(bar(S_EMPTY + s3), ((u / SS) & 1) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((s * TILE_BYTES) >> 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t off = (uint64_t)(((i >> 2) * BLK + (i & 3) * 32) >> 4);
          umma_f16(tS0 + s3 * 128, dQ0 + off, dK0 + so + off, idesc_s, i > 0);
        }
        umma_commit(bar(K_EMPTY + s));
        umma_commit(bar(S_FULL + s3));
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ============================== issuer 2: O += P V (P from tensor memory) ==============================
    const uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);  // V is MN-major
    const uint64_t dV0 = make_smem_desc(sbase + OFF_V, BLK, 1024);
    for (int u = 0; u < 2 * n; ++u) {
      const int s3 = u % SS;
      mbar_wait(bar(P_FULL + s3), (u / SS) & 1);
      if (u < n) {  // pass A: the softmax warps only took the row maxima; hand the stage back
        if (lane == 0) mbar_arrive(bar(S_EMPTY + s3));
        __syncwarp();
        continue;
      }
      const int j = u - n, vs = j % VS;
      mbar_wait(bar(V_FULL + vs), (j / VS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((vs * TILE_BYTES) >> 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // A = P: keys [16i, 16i+16) = 8 TMEM columns; keys of half hf = i >> 2 start at column 64 * hf
          const uint32_t acol = s3 * 128 + (i >> 2) * 64 + (i & 3) * 8;
          umma_f16_ts(tO, tS0 + acol, dV0 + so + (uint64_t)((i * 2048) >> 4), idesc_o, (j > 0 || i > 0) ? 1u : 0u);
        }
        umma_commit(bar(V_EMPTY + vs));
        umma_commit(bar(S_EMPTY + s3));
        if (j == n - 1) umma_commit(bar(O_FULL));
      }
      __syncwarp();
    }
  } else {
    // ============================== softmax + epilogue warps ==============================
    const int q4 = warp & 3, half = warp >> 2;  // TMEM lane quarter, column half
    const int row = q4 * 32 + lane;
    const int q = q0 + row;
    const bool row_valid = q < len;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    float* s_red = reinterpret_cast<float*>(sgen + OFF_RED);
    float mx = -INFINITY, m2 = 0.f, l = 0.f;
    for (int u = 0; u < 2 * n; ++u) {
      const int j = u % n, s3 = u % SS;
      const bool pass_b = u >= n;
      if (u == n) {  // combine the row maxima of the two column halves (once per CTA)
        s_red[half * 128 + row] = mx;
        softmax_bar_sync();
        const float m = fmaxf(s_red[row], s_red[128 + row]);
        m2 = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        softmax_bar_sync();
      }
      mbar_wait(bar(S_FULL + s3), (u / SS) & 1);
      tc_fence_after();
      const uint32_t tS = tS0 + s3 * 128 + lane_base + half * 64;
      // only the LAST key tile can hold padded keys (n = ceil(len / 128)); rows of padded queries in this tile are
      // computed like any other (their Q rows are zero-filled by the projection) and dropped in the epilogue
      const bool key_mask = (j + 1) * BKV > len;
      // both 32-column loads of this warp's slice are issued before the single wait (one TMEM round trip per tile)
      uint32_t vv[2][32];
      tmem_ld32(tS, vv[0]);
      tmem_ld32(tS + 32, vv[1]);
      tmem_ld_wait();
      if (!pass_b) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t (&v)[32] = vv[c];
          const int k0 = j * BKV + half * 64 + c * 32;
          if (key_mask) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (k0 + i < len) ? __uint_as_float(v[i]) : -INFINITY);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
      } else {
        uint32_t w[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t (&v)[32] = vv[c];
          const int k0 = j * BKV + half * 64 + c * 32;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m2));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m2));
            if (key_mask) {
              p0 = (k0 + 2 * i < len) ? p0 : 0.f;
              p1 = (k0 + 2 * i + 1 < len) ? p1 : 0.f;
            }
            l += p0 + p1;
            __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
            w[c * 16 + i] = *reinterpret_cast<uint32_t*>(&b2);
          }
        }
        // P (64 keys of this half, bf16) over the first 32 of the 64 S columns this warp has just consumed
        uint32_t w0[16], w1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          w0[i] = w[i];
          w1[i] = w[16 + i];
        }
        tmem_st16(tS, w0);
        tmem_st16(tS + 16, w1);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_FULL + s3));
    }
    // ---- combine the row sums, then epilogue: O / l -> bf16 -> staging (the K ring is idle once the last QK^T
    //      has retired, which it has long before O_FULL) -> coalesced global store; each half stores 64 columns
    s_red[half * 128 + row] = l;
    softmax_bar_sync();
    l = s_red[row] + s_red[128 + row];
    mbar_wait(bar(O_FULL), 0);
    tc_fence_after();
    const float inv = (row_valid && l > 0.f) ? 1.f / l : 0.f;  // padded query rows: out = 0, lse2 = +inf
    if (half == 0 && q < p.T) p.lse2[(long long)z * p.T + q] = (row_valid && l > 0.f) ? m2 + log2f(l) : INFINITY;
    uint8_t* stg = sgen + OFF_STG + warp * 4096;
    uint8_t* my = stg + lane * 128;
    const int lsw = lane & 7;
    {
      float f[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + lane_base + half * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[c * 32 + i] = row_valid ? __uint_as_float(v[i]) * inv : 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat162 b2 = __floats2bfloat162_rn(f[ch * 8 + 2 * t], f[ch * 8 + 2 * t + 1]);
          w[t] = *reinterpret_cast<uint32_t*>(&b2);
        }
        *reinterpret_cast<uint4*>(my + ((ch ^ lsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), ch = lane & 7;
        const int gq = q0 + q4 * 32 + r;
        if (gq < p.T) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + gq) * HD + h * DK + half * 64 + ch * 8) = val;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Work order of the attention kernels for a ragged batch: one CTA = one (utterance*head z, 128-row tile); its work
// is proportional to the utterance length, which varies 8x inside a LibriTTS-shaped batch while only ~4 working
// CTAs land on each SM.  Longest-processing-time-first: tiles of the longest utterance first, tiles that hold only
// padded frames (zero fill, no MMAs) last.  Built once per FFT stack call from the device-side lengths.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) attn_schedule_kernel(const int64_t* __restrict__ lens, int B, int T, int H,
                                                             int nq, int32_t* __restrict__ sched) {
  pdl_sync();
  __shared__ int s_len[1024], s_rank[1024], s_basev[1024], s_basep[1024];
  const int b = threadIdx.x;
  if (b < B) {
    long long l = lens[b];
    s_len[b] = (int)(l < 0 ? 0 : (l > T ? T : l));
  }
  __syncthreads();
  if (b < B) {
    int r = 0;
    for (int o = 0; o < B; ++o) r += (s_len[o] > s_len[b]) || (s_len[o] == s_len[b] && o < b);
    s_rank[r] = b;  // utterance at sorted position r
  }
  __syncthreads();
  if (b == 0) {
    int accv = 0, accp = 0;
    for (int r = 0; r < B; ++r) {
      const int nv = (s_len[s_rank[r]] + 127) / 128;
      s_basev[r] = accv;
      s_basep[r] = accp;
      accv += nv * H;
      accp += (nq - nv) * H;
    }
    for (int r = 0; r < B; ++r) s_basep[r] += accv;  // padded tiles after all valid ones
  }
  __syncthreads();
  if (b < B) {  // thread = sorted position
    const int u = s_rank[b];
    const int nv = (s_len[u] + 127) / 128;
    for (int h = 0; h < H; ++h) {
      const int z = u * H + h;
      for (int t = 0; t < nv; ++t) sched[s_basev[b] + h * nv + t] = (z << 8) | t;
      for (int t = nv; t < nq; ++t) sched[s_basep[b] + h * (nq - nv) + (t - nv)] = (z << 8) | t;
    }
  }
}

}  // namespace fs2

extern "C" {

// sched: int32 [B*H*ceil(T/128)] (written): work order for fs2_attn_fwd_bf16 / fs2_attn_bwd_bf16 on this batch.
int fs2_attn_schedule(const int64_t* lens, int B, int T, int H, int32_t* sched, void* stream) {
  using namespace fs2;
  if (B <= 0 || T <= 0) return 0;
  const int nq = (T + 127) / 128;
  if (B > 1024 || nq > 255 || (long long)B * H >= (1 << 23)) return set_error("attn_schedule: B <= 1024, T <= 32640");
  FS2_LAUNCH((attn_schedule_kernel), 1, 1024, 0, static_cast<cudaStream_t>(stream), lens, B, T, H, nq, sched);
  count_launch();
  return check_launch("attn_schedule_kernel");
}

// qkv: bf16 [B][T][3*H*128] (Q | K | V, head h = columns [h*128, (h+1)*128) of each third);
// lens: int64 [B]; sched: optional work order from fs2_attn_schedule (NULL = natural order);
// out: bf16 [B][T][H*128]; lse2: f32 [B*H][T] (log2-domain log-sum-exp of the scaled
// scores; +inf on padded rows).  scale = 1/sqrt(dk) is applied inside.
int fs2_attn_fwd_bf16(const void* qkv, const int64_t* lens, const int32_t* sched, int B, int T, int H, int dk,
                      void* out, float* lse2, void* stream) {
  using namespace fs2;
  if (dk != af::DK) return set_error("attn_fwd: d_k must be 128");
  if (B <= 0 || T <= 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         af::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, af2::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_fwd)", e);
    attr = true;
  }
  const int C3 = 3 * H * dk;
  CUtensorMap tmQK, tmV;
  if (int rc = make_tmap_bf16_3d(&tmQK, qkv, C3, T, B, C3, (long long)T * C3, 64, 128)) return rc;
  if (int rc = make_tmap_bf16_3d(&tmV, qkv, C3, T, B, C3, (long long)T * C3, 64, 64)) return rc;
  AttnFwdP p{};
  p.lens = lens;
  p.sched = sched;
  p.B = B; p.T = T; p.H = H;
  p.nq = (T + af::BQ - 1) / af::BQ;
  p.nkv = (T + af::BKV - 1) / af::BKV;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)dk);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse2 = lse2;
  static const bool old_fwd = getenv("FS2_ATTN_OLD") != nullptr;  // A/B switch: first-generation kernel
  if (old_fwd) {
    FS2_LAUNCH((attn_fwd_kernel), p.nq * B * H, 320, af::SMEM_BYTES, static_cast<cudaStream_t>(stream), tmQK, tmV, p);
  } else {
    FS2_LAUNCH((attn_fwd2_kernel), p.nq * B * H, 352, af2::SMEM_BYTES, static_cast<cudaStream_t>(stream), tmQK, tmV,
               p);
  }
  count_launch();
  return check_launch("attn_fwd_kernel");
}
}
