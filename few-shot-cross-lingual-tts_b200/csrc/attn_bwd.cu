// Fused attention backward for sm_100a (d_k = 128): two tcgen05 kernels + a row-dot prologue.
//
// Replaces autograd through transformer/Modules.py:14-25 (bmm, /sqrt(dk), masked softmax, bmm):
//   D_q   = sum_d dO[q,d] * O[q,d]                                   (attn_bwd_prep_kernel)
//   P     = exp2(S*c - lse2_q),  S = Q K^T       (recomputed, never stored)
//   dP    = dO V^T ;  dS = P o (dP - D_q) / sqrt(dk)
//   dV_j += P^T dO ; dK_j += dS^T Q                (attn_bwd_dkv_kernel: CTA = (z, 128-key tile), loops q)
//   dQ_i += dS K                                    (attn_bwd_dq_kernel : CTA = (z, 128-query tile), loops k)
// The dkv kernel works on TRANSPOSED scores (S^T = K Q^T: TMEM lane = key), so that P^T / dS^T leave the
// softmax warps already K-major for the dV / dK MMAs; per-query statistics (lse2, D) are then per-column
// broadcast loads and no row reduction is needed anywhere in the backward.  dQ gets its own kernel
// (S and dP recomputed once more: 7 instead of 5 MMAs per tile pair) instead of fp32 atomics.
// All operand tiles are [rows][64 bf16] swizzle-128B blocks; one tile serves both as a K-major operand
// (S / dP MMAs) and as an MN-major operand (dV / dK / dQ MMAs) -- same bytes, different descriptor.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"
#include "util.cuh"

namespace fs2 {

namespace ab {
constexpr int DK = 128;
constexpr int BLK128 = 128 * 128;  // [128 rows x 64 cols] bf16 block = 16 KiB
constexpr int BLK64 = 64 * 128;    // [ 64 rows x 64 cols] bf16 block =  8 KiB
}  // namespace ab

struct AttnBwdP {
  const int64_t* lens;
  const int32_t* sched;  // optional work order (fs2_attn_schedule)
  const float* lse2;  // [Z][T]
  const float* dsum;  // [Z][T]  D_q
  int B, T, H, n_outer, n_inner;
  float scale, scale_log2;
  __nv_bfloat16* dqkv;  // [B][T][3*H*dk]
  int dbg;              // FS2_ATTN_DBG ablation bits (tools only)
  float* dbq;           // optional [H*dk] f32, accumulated: column sums of dQ / dV = gradients of the w_qs / w_vs
  float* dbv;           // biases (the key bias has none: softmax is invariant to it)
};

// ------------------------------------------------------------------------------------------------
// D = rowsum(dO o O) per (b, head, q); one warp per (b, q) row, both heads
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                     const int64_t* __restrict__ lens, int B, int T, int H, float scale, float* __restrict__ dsum) {
  pdl_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)B * T) return;
  const int lane = threadIdx.x & 31;
  const int b = row / T, t = row - (long long)b * T;
  if (t >= lens[b]) return;  // padded query: its D is never read (and dO may be unwritten there)
  const int HD = H * ab::DK;
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    const int c = h * ab::DK + lane * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(o + row * HD + c);
    const uint2 g = *reinterpret_cast<const uint2*>(d_o + row * HD + c);
    const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.x));
    const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.y));
    const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.x));
    const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g.y));
    s = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y;
    s = warp_sum(s);
    if (lane == 0) dsum[((long long)b * H + h) * T + t] = s * scale;  // pre-scaled: dS = P*(dP*c - D*c)
  }
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar_sync_bwd() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// write 32 fp32 values (already final) as bf16 into this thread's 128-byte swizzled smem row, 16-byte
// chunks [chunk0, chunk0+4)
__device__ __forceinline__ void store_row_chunks(uint8_t* row_ptr, int sw, int chunk0, const float (&f)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 2 * t], f[g * 8 + 2 * t + 1]);
      w[t] = *reinterpret_cast<uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(row_ptr + (((chunk0 + g) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// 64 columns [half*64, half*64+64) of a TMEM accumulator [128 lanes x 128 cols] -> bf16 -> this warp's
// staging tile -> coalesced global rows (warp = lane quarter q4, rows row0 + q4*32 ...)
// colsum (optional, f32 [128], accumulated): column sums of the stored bf16 tile over its rows < row_limit -- the
// projection's bias gradient, taken from the staging tile on its way out (rows of padded queries / keys are zero)
__device__ __forceinline__ void store_acc_half(uint32_t tacc, uint32_t lane_base, uint8_t* stg, int q4, int lane,
                                               int half, __nv_bfloat16* gbase, long long g_ld, int row0,
                                               int row_limit, float* colsum = nullptr) {
  uint8_t* my = stg + lane * 128;
  const int lsw = lane & 7;
  float f[64];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tacc + lane_base + half * 64 + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[c * 32 + i] = __uint_as_float(v[i]);
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
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3), ch = lane & 7;
    const int grow = row0 + q4 * 32 + r;
    if (grow < row_limit) {
      const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
      *reinterpret_cast<uint4*>(gbase + (long long)grow * g_ld + half * 64 + ch * 8) = val;
      if (colsum) {
        const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          cs[2 * t] += __uint_as_float(w[t] << 16);
          cs[2 * t + 1] += __uint_as_float(w[t] & 0xFFFF0000u);
        }
      }
    }
  }
  if (colsum) {  // lanes with the same 8-column group (lane & 7) hold four row groups: fold them, 8 atomics per group
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
      cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(colsum + half * 64 + lane * 8 + j, cs[j]);
    }
  }
  __syncwarp();
}

// ================================================================================================
// dK / dV kernel
// ================================================================================================
namespace dkv {
using namespace ab;
constexpr int OFF_K = 0;                       // [128 keys x 128 d]  32 KiB
constexpr int OFF_V = OFF_K + 2 * BLK128;      // 32 KiB
constexpr int OFF_Q = OFF_V + 2 * BLK128;      // 2 stages x [64 q x 128 d] 16 KiB
constexpr int OFF_DO = OFF_Q + 2 * 2 * BLK64;  // 2 stages x 16 KiB
constexpr int OFF_PT = OFF_DO + 2 * 2 * BLK64; // 2 x P^T  [128 keys x 64 q] 16 KiB
constexpr int OFF_DST = OFF_PT + 2 * BLK128;   // 2 x dS^T 16 KiB
constexpr int OFF_STAT = OFF_DST + 2 * BLK128; // [2 parities][lse2 | dsum][64] f32
constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 64 * 4;
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
enum { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = 3, SP_FULL = 5, SP_EMPTY = 7, PDS_FULL = 9, PDS_EMPTY = 11,
       ACC_FULL = 13 };
}  // namespace dkv

__global__ void __launch_bounds__(320, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmKV,   // qkv, box 64 x 128 rows
                    const __grid_constant__ CUtensorMap tmQ,    // qkv, box 64 x 64 rows
                    const __grid_constant__ CUtensorMap tmDO,   // dO,  box 64 x 64 rows
                    const __grid_constant__ AttnBwdP p) {
  pdl_sync();
  using namespace dkv;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt = blockIdx.x % p.n_outer, z = blockIdx.x / p.n_outer;
  if (p.sched) {
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    jt = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int k0 = jt * 128;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (k0 >= len) {  // a key tile of padded frames only (CTA-uniform): dK = dV = 0, no MMAs
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
      const int r = i >> 5, c = i & 31;  // 32 x 16-byte vectors per row: 16 for dK, 16 for dV
      if (k0 + r < p.T)
        *reinterpret_cast<uint4*>(gb + (long long)(k0 + r) * 3 * HD + (1 + (c >> 4)) * HD + h * DK + (c & 15) * 8) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const int n = min(p.n_inner, (len + 63) / 64);  // 64-query tiles with at least one valid query

  if (threadIdx.x == 0) {
    mbar_init(bar(KV_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(QDO_FULL + s), 1);
      mbar_init(bar(QDO_EMPTY + s), 1);
      mbar_init(bar(SP_FULL + s), 1);
      mbar_init(bar(SP_EMPTY + s), 8);
      mbar_init(bar(PDS_FULL + s), 8);
      mbar_init(bar(PDS_EMPTY + s), 1);
    }
    mbar_init(bar(ACC_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  const uint32_t tSt = tmem_base, tdPt = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(KV_FULL), 4 * BLK128);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(sbase + OFF_K + kb * BLK128, &tmKV, bar(KV_FULL), HD + h * DK + kb * 64, k0, b);
        tma_load_3d(sbase + OFF_V + kb * BLK128, &tmKV, bar(KV_FULL), 2 * HD + h * DK + kb * 64, k0, b);
      }
      for (int i = 0; i < n; ++i) {
        const int s = i & 1;
        mbar_wait(bar(QDO_EMPTY + s), ((i >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(bar(QDO_FULL + s), 4 * BLK64);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_3d(sbase + OFF_Q + s * 2 * BLK64 + kb * BLK64, &tmQ, bar(QDO_FULL + s), h * DK + kb * 64,
                      i * 64, b);
          tma_load_3d(sbase + OFF_DO + s * 2 * BLK64 + kb * BLK64, &tmDO, bar(QDO_FULL + s), h * DK + kb * 64,
                      i * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);    // S^T / dP^T: [128 keys x 64 q]
      const uint32_t idesc_a = make_idesc_bf16(128, 128, 0, 1);   // dV / dK: B operand MN-major
      auto issue_sp = [&](int i) {
        const int s = i & 1;
        mbar_wait(bar(QDO_FULL + s), (i >> 1) & 1);
        mbar_wait(bar(SP_EMPTY + s), ((i >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sq = sbase + OFF_Q + s * 2 * BLK64, sdo = sbase + OFF_DO + s * 2 * BLK64;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t offa = (t >> 2) * BLK128 + (t & 3) * 32, offb = (t >> 2) * BLK64 + (t & 3) * 32;
          umma_f16(tSt + s * 64, make_smem_desc(sbase + OFF_K + offa, 16, 1024),
                   make_smem_desc(sq + offb, 16, 1024), idesc_s, t > 0);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t offa = (t >> 2) * BLK128 + (t & 3) * 32, offb = (t >> 2) * BLK64 + (t & 3) * 32;
          umma_f16(tdPt + s * 64, make_smem_desc(sbase + OFF_V + offa, 16, 1024),
                   make_smem_desc(sdo + offb, 16, 1024), idesc_s, t > 0);
        }
        umma_commit(bar(SP_FULL + s));
      };
      mbar_wait(bar(KV_FULL), 0);
      issue_sp(0);
      for (int i = 0; i < n; ++i) {
        const int s = i & 1;
        if (i + 1 < n) issue_sp(i + 1);
        mbar_wait(bar(PDS_FULL + s), (i >> 1) & 1);
        tc_fence_after();
        const uint32_t sq = sbase + OFF_Q + s * 2 * BLK64, sdo = sbase + OFF_DO + s * 2 * BLK64;
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // K = 64 queries, 16 per step
          umma_f16(tdV, make_smem_desc(sbase + OFF_PT + s * BLK128 + t * 32, 16, 1024),
                   make_smem_desc(sdo + t * 2048, BLK64, 1024), idesc_a, (i > 0 || t > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          umma_f16(tdK, make_smem_desc(sbase + OFF_DST + s * BLK128 + t * 32, 16, 1024),
                   make_smem_desc(sq + t * 2048, BLK64, 1024), idesc_a, (i > 0 || t > 0) ? 1u : 0u);
        }
        umma_commit(bar(QDO_EMPTY + s));
        umma_commit(bar(PDS_EMPTY + s));
        if (i == n - 1) umma_commit(bar(ACC_FULL));
      }
    }
  } else {
    // softmax warps: thread = key row; warps w / w+4 take the query columns [0,32) / [32,64) of the tile
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const int key = k0 + row;
    const bool key_valid = key < len;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    uint8_t* pt_row = sgen + OFF_PT + row * 128;
    uint8_t* dst_row = sgen + OFF_DST + row * 128;
    float* s_stat = reinterpret_cast<float*>(sgen + OFF_STAT);
    const int sw = row & 7;
    const int tid = threadIdx.x;  // 0..255 here
    const float* lse2 = p.lse2 + (long long)z * p.T;
    const float* dsum = p.dsum + (long long)z * p.T;
    for (int i = 0; i < n; ++i) {
      const int s = i & 1;
      // per-query statistics of this 64-query tile -> smem (double buffered by tile parity)
      if (tid < 128) {
        const int qq = i * 64 + (tid & 63);
        float val;
        if (tid < 64) val = qq < p.T ? __ldg(lse2 + qq) : INFINITY;  // +inf => P = 0 (padded queries too)
        else val = qq < len ? __ldg(dsum + qq) : 0.f;  // D of padded queries is not computed
        s_stat[s * 128 + tid] = val;
      }
      softmax_bar_sync_bwd();
      mbar_wait(bar(SP_FULL + s), (i >> 1) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32];
      tmem_ld32(tSt + s * 64 + lane_base + half * 32, vs);
      tmem_ld32(tdPt + s * 64 + lane_base + half * 32, vd);
      tmem_ld_wait();
      float pf[32], df[32];
      const float4* st_l = reinterpret_cast<const float4*>(s_stat + s * 128 + half * 32);
      const float4* st_d = st_l + 16;
      // invalid key rows (beyond the utterance) get P = dS = 0 through an infinite "lse" offset
      const float kill = key_valid ? 0.f : INFINITY;
#pragma unroll
      for (int t4 = 0; t4 < 8; ++t4) {
        const float4 l4 = st_l[t4], d4 = st_d[t4];
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int t = t4 * 4 + e;
          const float pr = ex2_fast(fmaf(__uint_as_float(vs[t]), p.scale_log2, -(lv[e] + kill)));
          pf[t] = pr;
          df[t] = pr * fmaf(__uint_as_float(vd[t]), p.scale, -dv[e]);
        }
      }
      mbar_wait(bar(PDS_EMPTY + s), ((i >> 1) & 1) ^ 1u);  // dV / dK MMAs of tile i-2 consumed this buffer
      store_row_chunks(pt_row + s * BLK128, sw, half * 4, pf);
      store_row_chunks(dst_row + s * BLK128, sw, half * 4, df);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(SP_EMPTY + s));
        mbar_arrive(bar(PDS_FULL + s));
      }
    }
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    uint8_t* stg = sgen + OFF_PT + warp * 4096;  // P^T / dS^T tiles (32 KiB) are free now
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    store_acc_half(tdV, lane_base, stg, q4, lane, half, gb + 2 * HD + h * DK, 3 * HD, k0, p.T,
                   p.dbv ? p.dbv + h * DK : nullptr);
    store_acc_half(tdK, lane_base, stg, q4, lane, half, gb + HD + h * DK, 3 * HD, k0, p.T);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// dQ kernel
// ================================================================================================
namespace dq {
using namespace ab;
constexpr int OFF_Q = 0;                       // [128 q x 128 d] 32 KiB
constexpr int OFF_DO = OFF_Q + 2 * BLK128;     // 32 KiB
constexpr int OFF_K = OFF_DO + 2 * BLK128;     // 2 stages x [64 keys x 128 d] 16 KiB
constexpr int OFF_V = OFF_K + 2 * 2 * BLK64;   // 2 stages x 16 KiB
constexpr int OFF_DS = OFF_V + 2 * 2 * BLK64;  // 2 x dS [128 q x 64 keys] 16 KiB (also the epilogue staging)
constexpr int OFF_STG = OFF_DS;
constexpr int OFF_BAR = OFF_DS + 2 * BLK128;
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
enum { QDO_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, SP_FULL = 5, SP_EMPTY = 7, DS_FULL = 9, DS_EMPTY = 11,
       ACC_FULL = 13 };
}  // namespace dq

__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ128,  // qkv, box 64 x 128 rows
                   const __grid_constant__ CUtensorMap tmKV64,  // qkv, box 64 x 64 rows
                   const __grid_constant__ CUtensorMap tmDO128, // dO,  box 64 x 128 rows
                   const __grid_constant__ AttnBwdP p) {
  pdl_sync();
  using namespace dq;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int it = blockIdx.x % p.n_outer, z = blockIdx.x / p.n_outer;
  if (p.sched) {
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    it = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int q0 = it * 128;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (q0 >= len) {  // a query tile of padded frames only (CTA-uniform): dQ = 0, no MMAs
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
      const int r = i >> 4, c = i & 15;
      if (q0 + r < p.T)
        *reinterpret_cast<uint4*>(gb + (long long)(q0 + r) * 3 * HD + h * DK + c * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const int n = min(p.n_inner, (len + 63) / 64);  // 64-key tiles with at least one valid key

  if (threadIdx.x == 0) {
    mbar_init(bar(QDO_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(KV_FULL + s), 1);
      mbar_init(bar(KV_EMPTY + s), 1);
      mbar_init(bar(SP_FULL + s), 1);
      mbar_init(bar(SP_EMPTY + s), 8);
      mbar_init(bar(DS_FULL + s), 8);
      mbar_init(bar(DS_EMPTY + s), 1);
    }
    mbar_init(bar(ACC_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdQ = tmem_base + 256;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(QDO_FULL), 4 * BLK128);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(sbase + OFF_Q + kb * BLK128, &tmQ128, bar(QDO_FULL), h * DK + kb * 64, q0, b);
        tma_load_3d(sbase + OFF_DO + kb * BLK128, &tmDO128, bar(QDO_FULL), h * DK + kb * 64, q0, b);
      }
      for (int j = 0; j < n; ++j) {
        const int s = j & 1;
        mbar_wait(bar(KV_EMPTY + s), ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(bar(KV_FULL + s), 4 * BLK64);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_3d(sbase + OFF_K + s * 2 * BLK64 + kb * BLK64, &tmKV64, bar(KV_FULL + s),
                      HD + h * DK + kb * 64, j * 64, b);
          tma_load_3d(sbase + OFF_V + s * 2 * BLK64 + kb * BLK64, &tmKV64, bar(KV_FULL + s),
                      2 * HD + h * DK + kb * 64, j * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // S / dP: [128 q x 64 keys]
      const uint32_t idesc_q = make_idesc_bf16(128, 128, 0, 1);  // dQ += dS K (K tile as MN-major B)
      auto issue_sp = [&](int j) {
        const int s = j & 1;
        mbar_wait(bar(KV_FULL + s), (j >> 1) & 1);
        mbar_wait(bar(SP_EMPTY + s), ((j >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sk = sbase + OFF_K + s * 2 * BLK64, sv = sbase + OFF_V + s * 2 * BLK64;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t offa = (t >> 2) * BLK128 + (t & 3) * 32, offb = (t >> 2) * BLK64 + (t & 3) * 32;
          umma_f16(tS + s * 64, make_smem_desc(sbase + OFF_Q + offa, 16, 1024),
                   make_smem_desc(sk + offb, 16, 1024), idesc_s, t > 0);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t offa = (t >> 2) * BLK128 + (t & 3) * 32, offb = (t >> 2) * BLK64 + (t & 3) * 32;
          umma_f16(tdP + s * 64, make_smem_desc(sbase + OFF_DO + offa, 16, 1024),
                   make_smem_desc(sv + offb, 16, 1024), idesc_s, t > 0);
        }
        umma_commit(bar(SP_FULL + s));
      };
      mbar_wait(bar(QDO_FULL), 0);
      issue_sp(0);
      for (int j = 0; j < n; ++j) {
        const int s = j & 1;
        if (j + 1 < n) issue_sp(j + 1);
        mbar_wait(bar(DS_FULL + s), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sk = sbase + OFF_K + s * 2 * BLK64;
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // K = 64 keys, 16 per step
          umma_f16(tdQ, make_smem_desc(sbase + OFF_DS + s * BLK128 + t * 32, 16, 1024),
                   make_smem_desc(sk + t * 2048, BLK64, 1024), idesc_q, (j > 0 || t > 0) ? 1u : 0u);
        }
        umma_commit(bar(KV_EMPTY + s));
        umma_commit(bar(DS_EMPTY + s));
        if (j == n - 1) umma_commit(bar(ACC_FULL));
      }
    }
  } else {
    // softmax warps: thread = query row; warps w / w+4 take the key columns [0,32) / [32,64) of the tile
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    uint8_t* ds_row = sgen + OFF_DS + row * 128;
    const int sw = row & 7;
    const bool q_ok = q < len;
    const float l2 = q_ok ? p.lse2[(long long)z * p.T + q] : INFINITY;  // +inf => P = 0
    const float dq_sum = q_ok ? p.dsum[(long long)z * p.T + q] : 0.f;
    for (int j = 0; j < n; ++j) {
      const int s = j & 1;
      mbar_wait(bar(SP_FULL + s), (j >> 1) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32];
      tmem_ld32(tS + s * 64 + lane_base + half * 32, vs);
      tmem_ld32(tdP + s * 64 + lane_base + half * 32, vd);
      tmem_ld_wait();
      float df[32];
      const int kb = j * 64 + half * 32;
      if (kb + 32 <= len) {  // warp-uniform fast path: every key of this slice is valid
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pr = ex2_fast(fmaf(__uint_as_float(vs[t]), p.scale_log2, -l2));
          df[t] = pr * fmaf(__uint_as_float(vd[t]), p.scale, -dq_sum);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float pr = ex2_fast(fmaf(__uint_as_float(vs[t]), p.scale_log2, -l2));
          const float ds = pr * fmaf(__uint_as_float(vd[t]), p.scale, -dq_sum);
          df[t] = (kb + t < len) ? ds : 0.f;
        }
      }
      mbar_wait(bar(DS_EMPTY + s), ((j >> 1) & 1) ^ 1u);  // the dQ MMA of tile j-2 has consumed this buffer
      store_row_chunks(ds_row + s * BLK128, sw, half * 4, df);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(SP_EMPTY + s));
        mbar_arrive(bar(DS_FULL + s));
      }
    }
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    uint8_t* stg = sgen + OFF_STG + warp * 4096;
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    store_acc_half(tdQ, lane_base, stg, q4, lane, half, gb + h * DK, 3 * HD, q0, p.T,
                   p.dbq ? p.dbq + h * DK : nullptr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================
// dK / dV kernel, second generation (same ideas as the dQ kernel below): P^T and dS^T never touch shared memory --
// the softmax warps write them (bf16) over the S^T / dP^T accumulator columns they have just read and the dV / dK
// MMAs take their A operand from tensor memory.  The 64 KiB of shared memory this frees deepen the Q / dO ring from
// 2 to 4 stages (the TMA round trip leaves the per-tile dependency cycle); the S^T / dP^T MMAs and the dV / dK MMAs
// are issued by two different warps so that neither waits behind the other's barriers; per-query statistics are
// prefetched one tile ahead; MMA issue is warp-uniform (elect.sync).
// ================================================================================================
namespace dkv2 {
using namespace ab;
constexpr int QS = 4;
constexpr int OFF_K = 0;                       // [128 keys x 128 d]  32 KiB
constexpr int OFF_V = OFF_K + 2 * BLK128;      // 32 KiB
constexpr int OFF_Q = OFF_V + 2 * BLK128;      // QS stages x [64 q x 128 d] 16 KiB
constexpr int OFF_DO = OFF_Q + QS * 2 * BLK64; // QS stages x 16 KiB
constexpr int OFF_STG = OFF_Q;                 // epilogue staging (the ring is idle by then)
constexpr int OFF_STAT = OFF_DO + QS * 2 * BLK64;  // [2 parities][lse2 | dsum][64] f32
constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 64 * 4;
enum { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = QDO_FULL + QS, SP_FULL = QDO_EMPTY + QS, SP_EMPTY = SP_FULL + 2,
       PDS_FULL = SP_EMPTY + 2, ACC_FULL = PDS_FULL + 2, NUM_BARS = ACC_FULL + 1 };
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
}  // namespace dkv2

__global__ void __launch_bounds__(352, 1)
attn_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap tmKV,   // qkv, box 64 x 128 rows
                     const __grid_constant__ CUtensorMap tmQ,    // qkv, box 64 x 64 rows
                     const __grid_constant__ CUtensorMap tmDO,   // dO,  box 64 x 64 rows
                     const __grid_constant__ AttnBwdP p) {
  pdl_sync();
  using namespace dkv2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt = blockIdx.x % p.n_outer, z = blockIdx.x / p.n_outer;
  if (p.sched) {
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    jt = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int k0 = jt * 128;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (k0 >= len) {  // a key tile of padded frames only (CTA-uniform): dK = dV = 0, no MMAs
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
      const int r = i >> 5, c = i & 31;  // 32 x 16-byte vectors per row: 16 for dK, 16 for dV
      if (k0 + r < p.T)
        *reinterpret_cast<uint4*>(gb + (long long)(k0 + r) * 3 * HD + (1 + (c >> 4)) * HD + h * DK + (c & 15) * 8) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const int n = min(p.n_inner, (len + 63) / 64);  // 64-query tiles with at least one valid query

  if (threadIdx.x == 0) {
    mbar_init(bar(KV_FULL), 1);
    for (int s = 0; s < QS; ++s) {
      mbar_init(bar(QDO_FULL + s), 1);
      mbar_init(bar(QDO_EMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(SP_FULL + s), 1);
      mbar_init(bar(SP_EMPTY + s), 1);   // committed after the dV / dK MMAs that read P^T / dS^T of this stage
      mbar_init(bar(PDS_FULL + s), 8);
    }
    mbar_init(bar(ACC_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  const uint32_t tSt = tmem_base, tdPt = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(KV_FULL), 4 * BLK128);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(sbase + OFF_K + kb * BLK128, &tmKV, bar(KV_FULL), HD + h * DK + kb * 64, k0, b);
        tma_load_3d(sbase + OFF_V + kb * BLK128, &tmKV, bar(KV_FULL), 2 * HD + h * DK + kb * 64, k0, b);
      }
      for (int i = 0; i < n; ++i) {
        const int s = i % QS;
        mbar_wait(bar(QDO_EMPTY + s), ((i / QS) & 1) ^ 1u);
        mbar_arrive_expect_tx(bar(QDO_FULL + s), 4 * BLK64);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_3d(sbase + OFF_Q + s * 2 * BLK64 + kb * BLK64, &tmQ, bar(QDO_FULL + s), h * DK + kb * 64,
                      i * 64, b);
          tma_load_3d(sbase + OFF_DO + s * 2 * BLK64 + kb * BLK64, &tmDO, bar(QDO_FULL + s), h * DK + kb * 64,
                      i * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    // ---- issuer 1: S^T = K Q^T and dP^T = V dO^T for tile i into stage i & 1
    const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    const uint64_t dK0 = make_smem_desc(sbase + OFF_K, 16, 1024), dV0 = make_smem_desc(sbase + OFF_V, 16, 1024);
    const uint64_t dQ0 = make_smem_desc(sbase + OFF_Q, 16, 1024), dDO0 = make_smem_desc(sbase + OFF_DO, 16, 1024);
    mbar_wait(bar(KV_FULL), 0);
    for (int i = 0; i < n; ++i) {
      const int s = i & 1, qs = i % QS;
      mbar_wait(bar(QDO_FULL + qs), (i / QS) & 1);
      mbar_wait(bar(SP_EMPTY + s), ((i >> 1) & 1) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((qs * 2 * BLK64) >> 4);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint64_t offa = (uint64_t)(((t >> 2) * BLK128 + (t & 3) * 32) >> 4);
          const uint64_t offb = (uint64_t)(((t >> 2) * BLK64 + (t & 3) * 32) >> 4);
          umma_f16(tSt + s * 64, dK0 + offa, dQ0 + so + offb, idesc_s, t > 0);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint64_t offa = (uint64_t)(((t >> 2) * BLK128 + (t & 3) * 32) >> 4);
          const uint64_t offb = (uint64_t)(((t >> 2) * BLK64 + (t & 3) * 32) >> 4);
          umma_f16(tdPt + s * 64, dV0 + offa, dDO0 + so + offb, idesc_s, t > 0);
        }
        umma_commit(bar(SP_FULL + s));
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ---- issuer 2: dV += P^T dO, dK += dS^T Q with the A operands in tensor memory
    const uint32_t idesc_a = make_idesc_bf16(128, 128, 0, 1);  // B operand (dO / Q tile) MN-major
    const uint64_t dQm0 = make_smem_desc(sbase + OFF_Q, BLK64, 1024), dDOm0 = make_smem_desc(sbase + OFF_DO, BLK64, 1024);
    for (int i = 0; i < n; ++i) {
      const int s = i & 1, qs = i % QS;
      mbar_wait(bar(PDS_FULL + s), (i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((qs * 2 * BLK64) >> 4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // K = 64 queries, 16 per step; queries [32h, 32h+32) sit in columns [32h, 32h+16)
          const uint32_t acol = s * 64 + (t >> 1) * 32 + (t & 1) * 8;
          umma_f16_ts(tdV, tSt + acol, dDOm0 + so + (uint64_t)((t * 2048) >> 4), idesc_a, (i > 0 || t > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t acol = s * 64 + (t >> 1) * 32 + (t & 1) * 8;
          umma_f16_ts(tdK, tdPt + acol, dQm0 + so + (uint64_t)((t * 2048) >> 4), idesc_a, (i > 0 || t > 0) ? 1u : 0u);
        }
        umma_commit(bar(QDO_EMPTY + qs));
        umma_commit(bar(SP_EMPTY + s));
        if (i == n - 1) umma_commit(bar(ACC_FULL));
      }
      __syncwarp();
    }
  } else {
    // softmax warps: thread = key row; warps w / w+4 take the query columns [0,32) / [32,64) of the tile
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const int key = k0 + row;
    const bool key_valid = key < len;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    float* s_stat = reinterpret_cast<float*>(sgen + OFF_STAT);
    const int tid = threadIdx.x;  // 0..255 here
    const float* lse2 = p.lse2 + (long long)z * p.T;
    const float* dsum = p.dsum + (long long)z * p.T;
    auto load_stat = [&](int i) -> float {  // threads 0..63: lse2 of query i*64+tid, 64..127: dsum
      const int qq = i * 64 + (tid & 63);
      if (tid < 64) return qq < p.T ? __ldg(lse2 + qq) : INFINITY;  // +inf => P = 0 (padded queries too)
      return qq < len ? __ldg(dsum + qq) : 0.f;                     // D of padded queries is not computed
    };
    float nxt = tid < 128 ? load_stat(0) : 0.f;
    // invalid key rows (beyond the utterance) get P = dS = 0 through an infinite "lse" offset
    const float kill = key_valid ? 0.f : INFINITY;
    for (int i = 0; i < n; ++i) {
      const int s = i & 1;
      if (tid < 128) s_stat[s * 128 + tid] = nxt;  // (buffer s was last read in iteration i-2: a barrier lies between)
      if (tid < 128 && i + 1 < n) nxt = load_stat(i + 1);  // in flight during this tile's math
      softmax_bar_sync_bwd();
      mbar_wait(bar(SP_FULL + s), (i >> 1) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32];
      tmem_ld32(tSt + s * 64 + lane_base + half * 32, vs);
      tmem_ld32(tdPt + s * 64 + lane_base + half * 32, vd);
      tmem_ld_wait();
      uint32_t wp[16], wd[16];
      const float4* st_l = reinterpret_cast<const float4*>(s_stat + s * 128 + half * 32);
      const float4* st_d = st_l + 16;
#pragma unroll
      for (int t4 = 0; t4 < 8; ++t4) {
        const float4 l4 = st_l[t4], d4 = st_d[t4];
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
        float pr[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int t = t4 * 4 + e;
          pr[e] = ex2_fast(fmaf(__uint_as_float(vs[t]), p.scale_log2, -(lv[e] + kill)));
          ds[e] = pr[e] * fmaf(__uint_as_float(vd[t]), p.scale, -dv[e]);
        }
        __nv_bfloat162 a = __floats2bfloat162_rn(pr[0], pr[1]), c = __floats2bfloat162_rn(pr[2], pr[3]);
        __nv_bfloat162 d0 = __floats2bfloat162_rn(ds[0], ds[1]), d1 = __floats2bfloat162_rn(ds[2], ds[3]);
        wp[2 * t4] = *reinterpret_cast<uint32_t*>(&a);
        wp[2 * t4 + 1] = *reinterpret_cast<uint32_t*>(&c);
        wd[2 * t4] = *reinterpret_cast<uint32_t*>(&d0);
        wd[2 * t4 + 1] = *reinterpret_cast<uint32_t*>(&d1);
      }
      tmem_st16(tSt + s * 64 + lane_base + half * 32, wp);   // P^T over the S^T columns this warp consumed
      tmem_st16(tdPt + s * 64 + lane_base + half * 32, wd);  // dS^T over its dP^T columns
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PDS_FULL + s));
    }
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    uint8_t* stg = sgen + OFF_STG + warp * 4096;
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    store_acc_half(tdV, lane_base, stg, q4, lane, half, gb + 2 * HD + h * DK, 3 * HD, k0, p.T,
                   p.dbv ? p.dbv + h * DK : nullptr);
    store_acc_half(tdK, lane_base, stg, q4, lane, half, gb + HD + h * DK, 3 * HD, k0, p.T);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// dQ kernel, second generation: dS never touches shared memory.  The softmax warps write dS (bf16) back into
// the TMEM columns of the S accumulator they just read (tcgen05.st) and the dQ MMA takes its A operand from
// tensor memory (tcgen05.mma with [tmem] A); that frees 32 KiB of shared memory and the st.shared / proxy-fence
// / membar sequence per tile.  The freed memory deepens the K / V ring to 4 stages (the 2-stage ring put the
// TMA round trip into the per-tile dependency cycle) and S / dP get 3 TMEM stages, so the S / dP MMAs run two
// tiles ahead of the softmax warps.
// ================================================================================================
namespace dq2 {
using namespace ab;
constexpr int KVS = 4, SPS = 3;
constexpr int OFF_Q = 0;                        // [128 q x 128 d] 32 KiB
constexpr int OFF_DO = OFF_Q + 2 * BLK128;      // 32 KiB
constexpr int OFF_K = OFF_DO + 2 * BLK128;      // KVS stages x [64 keys x 128 d] 16 KiB
constexpr int OFF_V = OFF_K + KVS * 2 * BLK64;  // KVS stages x 16 KiB
constexpr int OFF_STG = OFF_K;                  // epilogue staging (the ring is idle by then)
constexpr int OFF_BAR = OFF_V + KVS * 2 * BLK64;
enum { QDO_FULL = 0, KV_FULL = 1, KV_EMPTY = KV_FULL + KVS, SP_FULL = KV_EMPTY + KVS, SP_EMPTY = SP_FULL + SPS,
       DS_FULL = SP_EMPTY + SPS, ACC_FULL = DS_FULL + SPS, NUM_BARS = ACC_FULL + 1 };
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;
}  // namespace dq2

__global__ void __launch_bounds__(352, 1)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tmQ128,  // qkv, box 64 x 128 rows
                    const __grid_constant__ CUtensorMap tmKV64,  // qkv, box 64 x 64 rows
                    const __grid_constant__ CUtensorMap tmDO128, // dO,  box 64 x 128 rows
                    const __grid_constant__ AttnBwdP p) {
  pdl_sync();
  using namespace dq2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int it = blockIdx.x % p.n_outer, z = blockIdx.x / p.n_outer;
  if (p.sched) {
    const int e = p.sched[blockIdx.x];
    z = e >> 8;
    it = e & 255;
  }
  const int b = z / p.H, h = z % p.H;
  const int q0 = it * 128;
  const int HD = p.H * DK;
  const int len = min((int)p.lens[b], p.T);
  if (q0 >= len) {  // a query tile of padded frames only (CTA-uniform): dQ = 0, no MMAs
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
      const int r = i >> 4, c = i & 15;
      if (q0 + r < p.T)
        *reinterpret_cast<uint4*>(gb + (long long)(q0 + r) * 3 * HD + h * DK + c * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const int n = min(p.n_inner, (len + 63) / 64);  // 64-key tiles with at least one valid key

  if (threadIdx.x == 0) {
    mbar_init(bar(QDO_FULL), 1);
    for (int s = 0; s < KVS; ++s) {
      mbar_init(bar(KV_FULL + s), 1);
      mbar_init(bar(KV_EMPTY + s), 1);
    }
    for (int s = 0; s < SPS; ++s) {
      mbar_init(bar(SP_FULL + s), 1);
      mbar_init(bar(SP_EMPTY + s), 1);
      mbar_init(bar(DS_FULL + s), 8);
    }
    mbar_init(bar(ACC_FULL), 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(sbase + OFF_TMEM, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);
  // columns: S stages [0,192), dP stages [192,384), dQ [384,512)
  const uint32_t tS = tmem_base, tdP = tmem_base + SPS * 64, tdQ = tmem_base + 2 * SPS * 64;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(QDO_FULL), 4 * BLK128);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(sbase + OFF_Q + kb * BLK128, &tmQ128, bar(QDO_FULL), h * DK + kb * 64, q0, b);
        tma_load_3d(sbase + OFF_DO + kb * BLK128, &tmDO128, bar(QDO_FULL), h * DK + kb * 64, q0, b);
      }
      for (int j = 0; j < n; ++j) {
        const int s = j % KVS;
        mbar_wait(bar(KV_EMPTY + s), ((j / KVS) & 1) ^ 1u);
        if (p.dbg & 32) {
          mbar_arrive(bar(KV_FULL + s));
          continue;
        }
        mbar_arrive_expect_tx(bar(KV_FULL + s), 4 * BLK64);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_3d(sbase + OFF_K + s * 2 * BLK64 + kb * BLK64, &tmKV64, bar(KV_FULL + s),
                      HD + h * DK + kb * 64, j * 64, b);
          tma_load_3d(sbase + OFF_V + s * 2 * BLK64 + kb * BLK64, &tmKV64, bar(KV_FULL + s),
                      2 * HD + h * DK + kb * 64, j * 64, b);
        }
      }
    }
  } else if (warp == 9) {
    // ---- issuer 1: S = Q K^T and dP = dO V^T for tile j into stage j % SPS (warp-uniform loop, one elected lane)
    const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // S / dP: [128 q x 64 keys]
    // descriptors: constant high word; the low word is (address >> 4) + (LBO >> 4 << 16) -> add offsets >> 4
    const uint64_t dQ0 = make_smem_desc(sbase + OFF_Q, 16, 1024), dDO0 = make_smem_desc(sbase + OFF_DO, 16, 1024);
    const uint64_t dK0 = make_smem_desc(sbase + OFF_K, 16, 1024), dV0 = make_smem_desc(sbase + OFF_V, 16, 1024);
    mbar_wait(bar(QDO_FULL), 0);
    for (int j = 0; j < n; ++j) {
      const int s = j % KVS, s3 = j % SPS;
      mbar_wait(bar(KV_FULL + s), (j / KVS) & 1);
      mbar_wait(bar(SP_EMPTY + s3), ((j / SPS) & 1) ^ 1u);  // the dQ MMA of tile j-3 has consumed dS in this stage
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((s * 2 * BLK64) >> 4);
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint64_t offa = (uint64_t)(((t >> 2) * BLK128 + (t & 3) * 32) >> 4);
            const uint64_t offb = (uint64_t)(((t >> 2) * BLK64 + (t & 3) * 32) >> 4);
            umma_f16(tS + s3 * 64, dQ0 + offa, dK0 + so + offb, idesc_s, t > 0);
          }
        }
        if (!(p.dbg & 16)) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint64_t offa = (uint64_t)(((t >> 2) * BLK128 + (t & 3) * 32) >> 4);
            const uint64_t offb = (uint64_t)(((t >> 2) * BLK64 + (t & 3) * 32) >> 4);
            umma_f16(tdP + s3 * 64, dDO0 + offa, dV0 + so + offb, idesc_s, t > 0);
          }
        }
        umma_commit(bar(SP_FULL + s3));
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ---- issuer 2: dQ += dS K with dS read from tensor memory
    const uint32_t idesc_q = make_idesc_bf16(128, 128, 0, 1);  // K tile as MN-major B
    const uint64_t dKm0 = make_smem_desc(sbase + OFF_K, BLK64, 1024);
    for (int j = 0; j < n; ++j) {
      const int s = j % KVS, s3 = j % SPS;
      mbar_wait(bar(DS_FULL + s3), (j / SPS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((s * 2 * BLK64) >> 4);
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {  // K = 64 keys, 16 per step; dS of keys [32h, 32h+32) sits in S columns [32h, 32h+16)
            umma_f16_ts(tdQ, tS + s3 * 64 + (t >> 1) * 32 + (t & 1) * 8, dKm0 + so + (uint64_t)((t * 2048) >> 4),
                        idesc_q, (j > 0 || t > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar(KV_EMPTY + s));
        umma_commit(bar(SP_EMPTY + s3));
        if (j == n - 1) umma_commit(bar(ACC_FULL));
      }
      __syncwarp();
    }
  } else {
    // softmax warps: thread = query row; warps w / w+4 take the key columns [0,32) / [32,64) of the tile
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const bool q_ok = q < len;
    const float l2 = q_ok ? p.lse2[(long long)z * p.T + q] : INFINITY;  // +inf => P = 0
    const float dq_sum = q_ok ? p.dsum[(long long)z * p.T + q] : 0.f;
    for (int j = 0; j < n; ++j) {
      const int s3 = j % SPS;
      mbar_wait(bar(SP_FULL + s3), (j / SPS) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32];
      if (!(p.dbg & 64)) {
        tmem_ld32(tS + s3 * 64 + lane_base + half * 32, vs);
        if (!(p.dbg & 2)) tmem_ld32(tdP + s3 * 64 + lane_base + half * 32, vd);
        tmem_ld_wait();
      }
      uint32_t w[16];
      const int kb = j * 64 + half * 32;
      if (p.dbg & 1) {
#pragma unroll
        for (int t = 0; t < 16; ++t) w[t] = vs[t] ^ vd[t];
      } else
      if (kb + 32 <= len) {  // warp-uniform fast path: every key of this slice is valid
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float p0 = ex2_fast(fmaf(__uint_as_float(vs[2 * t]), p.scale_log2, -l2));
          const float p1 = ex2_fast(fmaf(__uint_as_float(vs[2 * t + 1]), p.scale_log2, -l2));
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0 * fmaf(__uint_as_float(vd[2 * t]), p.scale, -dq_sum),
                                                    p1 * fmaf(__uint_as_float(vd[2 * t + 1]), p.scale, -dq_sum));
          w[t] = *reinterpret_cast<uint32_t*>(&b2);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float p0 = ex2_fast(fmaf(__uint_as_float(vs[2 * t]), p.scale_log2, -l2));
          const float p1 = ex2_fast(fmaf(__uint_as_float(vs[2 * t + 1]), p.scale_log2, -l2));
          const float d0 = p0 * fmaf(__uint_as_float(vd[2 * t]), p.scale, -dq_sum);
          const float d1 = p1 * fmaf(__uint_as_float(vd[2 * t + 1]), p.scale, -dq_sum);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(kb + 2 * t < len ? d0 : 0.f, kb + 2 * t + 1 < len ? d1 : 0.f);
          w[t] = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
      if (!(p.dbg & 64)) {
        tmem_st16(tS + s3 * 64 + lane_base + half * 32, w);  // dS over the S columns this warp has just consumed
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(DS_FULL + s3));
    }
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    uint8_t* stg = sgen + OFF_STG + warp * 4096;
    __nv_bfloat16* gb = p.dqkv + (long long)b * p.T * 3 * HD;
    store_acc_half(tdQ, lane_base, stg, q4, lane, half, gb + h * DK, 3 * HD, q0, p.T,
                   p.dbq ? p.dbq + h * DK : nullptr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fs2

extern "C" {

// qkv: bf16 [B][T][3*H*128]; o, d_o: bf16 [B][T][H*128]; lse2: f32 [B*H][T] from the forward;
// dsum: f32 [B*H][T] workspace; dqkv: bf16 [B][T][3*H*128] (every element is written).
int fs2_attn_bwd_bf16(const void* qkv, const void* o, const void* d_o, const float* lse2, const int64_t* lens,
                      const int32_t* sched, int B, int T, int H, int dk, float* dsum, void* dqkv, float* dbias_q,
                      float* dbias_v, void* stream) {
  using namespace fs2;
  if (dk != ab::DK) return set_error("attn_bwd: d_k must be 128");
  if (B <= 0 || T <= 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         dkv::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dq2::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv2::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(attn_bwd)", e);
    attr = true;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C3 = 3 * H * dk, HD = H * dk;
  const long long rows = (long long)B * T;
  FS2_LAUNCH((attn_bwd_prep_kernel), (unsigned)((rows + 7) / 8), 256, 0, s, static_cast<const __nv_bfloat16*>(o),
                                                                 static_cast<const __nv_bfloat16*>(d_o), lens, B, T,
                                                                 H, 1.f / sqrtf((float)dk), dsum);
  count_launch();
  if (int rc = check_launch("attn_bwd_prep_kernel")) return rc;
  CUtensorMap tm128, tm64, tmdo128, tmdo64;
  if (int rc = make_tmap_bf16_3d(&tm128, qkv, C3, T, B, C3, (long long)T * C3, 64, 128)) return rc;
  if (int rc = make_tmap_bf16_3d(&tm64, qkv, C3, T, B, C3, (long long)T * C3, 64, 64)) return rc;
  if (int rc = make_tmap_bf16_3d(&tmdo128, d_o, HD, T, B, HD, (long long)T * HD, 64, 128)) return rc;
  if (int rc = make_tmap_bf16_3d(&tmdo64, d_o, HD, T, B, HD, (long long)T * HD, 64, 64)) return rc;
  AttnBwdP p{};
  p.lens = lens; p.lse2 = lse2; p.dsum = dsum; p.sched = sched;
  p.B = B; p.T = T; p.H = H;
  p.scale = 1.f / sqrtf((float)dk);
  p.scale_log2 = 1.4426950408889634f * p.scale;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.dbq = dbias_q;
  p.dbv = dbias_v;
  static const int dbg = getenv("FS2_ATTN_DBG") ? atoi(getenv("FS2_ATTN_DBG")) : 0;
  p.dbg = dbg;
  p.n_outer = (T + 127) / 128;
  p.n_inner = (T + 63) / 64;
  const unsigned grid = (unsigned)(p.n_outer * B * H);
  static const bool old_dq = getenv("FS2_ATTN_OLD") != nullptr;  // A/B switch: first-generation kernels
  if (old_dq) {
    FS2_LAUNCH((attn_bwd_dkv_kernel), grid, 320, dkv::SMEM_BYTES, s, tm128, tm64, tmdo64, p);
  } else {
    FS2_LAUNCH((attn_bwd_dkv2_kernel), grid, 352, dkv2::SMEM_BYTES, s, tm128, tm64, tmdo64, p);
  }
  count_launch();
  if (int rc = check_launch("attn_bwd_dkv_kernel")) return rc;
  if (old_dq) {
    FS2_LAUNCH((attn_bwd_dq_kernel), grid, 320, dq::SMEM_BYTES, s, tm128, tm64, tmdo128, p);
  } else {
    FS2_LAUNCH((attn_bwd_dq2_kernel), grid, 352, dq2::SMEM_BYTES, s, tm128, tm64, tmdo128, p);
  }
  count_launch();
  return check_launch("attn_bwd_dq_kernel");
}
}
