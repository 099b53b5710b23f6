// Pieces shared by the 1-CTA (gemm_tc.cu) and 2-CTA (gemm_tc2.cu) tcgen05 GEMM kernels: kernel
// parameter block, tile decoding and the coalesced TMEM -> registers -> staging -> global epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.h"
#include "ptx.cuh"

namespace fs2 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 384;  // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
constexpr int kChunkBytes = 64 * 64 * 2;  // one 64x64 bf16 swizzle-128B box (MN-major operands)

struct GemmKP {
  // problem
  int mode, M, N, Z;
  int tiles_m, tiles_n, total_tiles;
  int a_mn, b_mn;
  int a_inner_base, a_zdiv, a_zmod_stride, a_batched;
  int b_inner_base, b_zdiv, b_zmod_stride, b_batched;
  // NORMAL
  int kb_per_tap, num_kb, tap_shift0, b_tap_kstride;
  // WGRAD
  int rb_per_batch, total_rb, kb_per_split, n_tiles_per_tap, n_per_tap;
  // epilogue
  int epilogue, d_f32, d_atomic, d_zdiv;
  float alpha;
  void* d;
  long long ldd, d_col_stride, d_tap_stride, d_zdiv_stride, d_zmod_stride;
  const float* bias;
  const __nv_bfloat16* aux;
  long long ld_aux, aux_batch_stride;
  int seg_rows;
  void* seg[4];
};

struct TileCoord {
  int z, tm, tn, nkb, kb0;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKP& p, int tile) {
  TileCoord t;
  t.tn = tile % p.tiles_n;
  int r = tile / p.tiles_n;
  t.tm = r % p.tiles_m;
  t.z = r / p.tiles_m;  // NORMAL: batch index; WGRAD: split index
  if (p.mode == FS2_GEMM_NORMAL) {
    t.nkb = p.num_kb;
    t.kb0 = 0;
  } else {
    t.kb0 = t.z * p.kb_per_split;
    int rem = p.total_rb - t.kb0;
    t.nkb = rem < p.kb_per_split ? rem : p.kb_per_split;
    if (t.nkb < 0) t.nkb = 0;
  }
  return t;
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one [32 rows x 128 B] output chunk of one warp: TMEM -> registers -> math -> this warp's
// 4 KiB staging area (16-byte chunks XOR-swizzled by row, conflict free both ways) -> global memory
// with 8 lanes covering one full 128-byte line per row (coalesced), instead of every thread dribbling
// 16-byte pieces of its own row.  The optional aux tile (ReLU-backward mask / residual add) is read
// through the same staging area with the same coalesced pattern.
// ------------------------------------------------------------------------------------------------
template <bool F32OUT, bool ATOMIC = false>
__device__ __forceinline__ void epilogue_chunk(const GemmKP& p, uint32_t taddr, uint8_t* stg, int lane,
                                               int m_w0, int n0, int nlimit, long long base_off,
                                               const __nv_bfloat16* aux_base, bool tile_ok) {
  constexpr int NC = F32OUT ? 32 : 64;  // accumulator columns per 128-byte output row segment
  float f[NC];
  {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    if constexpr (!F32OUT) {
      uint32_t v2[32];
      tmem_ld32(taddr + 32, v2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[32 + j] = __uint_as_float(v2[j]);
    } else {
      tmem_ld_wait();
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < NC; ++j) f[j] *= p.alpha;
    }
  }
  const int sw = lane & 7;
  uint8_t* my_row = stg + lane * 128;
  if (p.bias) {
    if (n0 + NC <= nlimit) {  // warp-uniform: whole chunk in range -> 16-byte broadcast loads
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        const float4 bv = __ldg(b4 + j);
        f[4 * j] += bv.x;
        f[4 * j + 1] += bv.y;
        f[4 * j + 2] += bv.z;
        f[4 * j + 3] += bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (n0 + j < nlimit) f[j] += __ldg(p.bias + n0 + j);
    }
  }
  if (p.epilogue == FS2_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if constexpr (!F32OUT) {
    if (aux_base) {  // bf16 aux tile, 64 columns = 128 B per row
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), ch = lane & 7;
        const int gm = m_w0 + r, col = n0 + ch * 8;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (gm < p.M && col < nlimit)
          val = __ldg(reinterpret_cast<const uint4*>(aux_base + (long long)gm * p.ld_aux + col));
        *reinterpret_cast<uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4)) = val;
      }
      __syncwarp();
      const bool bwd = p.epilogue == FS2_EPI_RELU_BWD;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 u = *reinterpret_cast<const uint4*>(my_row + ((ch ^ sw) << 4));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float lo = __uint_as_float(w[h] << 16), hi = __uint_as_float(w[h] & 0xFFFF0000u);
          float& f0 = f[ch * 8 + 2 * h];
          float& f1 = f[ch * 8 + 2 * h + 1];
          if (bwd) {
            f0 = lo > 0.f ? f0 : 0.f;
            f1 = hi > 0.f ? f1 : 0.f;
          } else {
            f0 += lo;
            f1 += hi;
          }
        }
      }
      __syncwarp();
    }
  }
  // own row -> staging
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 val;
    if constexpr (F32OUT) {
      val = make_uint4(__float_as_uint(f[4 * ch]), __float_as_uint(f[4 * ch + 1]),
                       __float_as_uint(f[4 * ch + 2]), __float_as_uint(f[4 * ch + 3]));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[ch * 8 + 2 * h], f[ch * 8 + 2 * h + 1]);
        w[h] = *reinterpret_cast<uint32_t*>(&b2);
      }
      val = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(my_row + ((ch ^ sw) << 4)) = val;
  }
  __syncwarp();
  // staging -> global: 8 lanes per row, 4 rows per instruction
  constexpr int EPV = F32OUT ? 4 : 8;  // elements per 16-byte vector
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3), ch = lane & 7;
    const int gm = m_w0 + r, col = n0 + ch * EPV;
    if (tile_ok && gm < p.M && col < nlimit) {
      const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
      long long off = base_off + (long long)gm * p.ldd + col;
      if constexpr (ATOMIC) {  // split-K accumulation: one 16-byte vector reduction per lane (coalesced)
        float* dbase = static_cast<float*>(p.d);
        if (p.seg_rows > 0) {
          const int sg = gm / p.seg_rows;
          dbase = static_cast<float*>(p.seg[sg]);
          off -= (long long)sg * p.seg_rows * p.ldd;
        }
        float* dp = dbase + off;
        if (col + 4 <= nlimit) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(__uint_as_float(val.x)),
                       "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                       : "memory");
        } else {
          const uint32_t w[4] = {val.x, val.y, val.z, val.w};
          for (int e = 0; e < 4 && col + e < nlimit; ++e) atomicAdd(dp + e, __uint_as_float(w[e]));
        }
      } else
      if (col + EPV <= nlimit) {
        if constexpr (F32OUT)
          *reinterpret_cast<uint4*>(static_cast<float*>(p.d) + off) = val;
        else
          *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.d) + off) = val;
      } else {  // ragged last vector of the row
        const uint32_t w[4] = {val.x, val.y, val.z, val.w};
        for (int e = 0; e < EPV && col + e < nlimit; ++e) {
          if constexpr (F32OUT) {
            static_cast<float*>(p.d)[off + e] = __uint_as_float(w[e]);
          } else {
            const uint16_t h = (e & 1) ? (uint16_t)(w[e >> 1] >> 16) : (uint16_t)(w[e >> 1] & 0xFFFFu);
            reinterpret_cast<uint16_t*>(p.d)[off + e] = h;
          }
        }
      }
    }
  }
  __syncwarp();
}

// Epilogue of one output tile for one epilogue warp (lane quarter q, column half chalf): every output
// of the engine has unit column stride, so both the plain and the split-K (vector-atomic) flavours go
// through the coalesced staging path of epilogue_chunk.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmKP& p, const TileCoord& t, uint32_t tmem_acc, uint8_t* stg,
                                              int q, int chalf, int lane) {
  int ncol0, nlimit;  // first column of this tile inside its (tap) column space, and its extent
  long long base_off;
  if (p.mode == FS2_GEMM_NORMAL) {
    ncol0 = t.tn * BN;
    nlimit = p.N;
    base_off = (long long)(t.z / p.d_zdiv) * p.d_zdiv_stride + (long long)(t.z % p.d_zdiv) * p.d_zmod_stride;
  } else {
    const int tap = t.tn / p.n_tiles_per_tap;
    ncol0 = (t.tn - tap * p.n_tiles_per_tap) * BN;
    nlimit = p.n_per_tap;
    base_off = (long long)tap * p.d_tap_stride;
  }
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int m_w0 = t.tm * BM + q * 32;
  const bool ok = t.nkb > 0;
  if (p.d_atomic) {  // split-K weight gradients: coalesced 16-byte vector reductions
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      if (ncol0 + c0 >= nlimit) break;  // warp-uniform
      epilogue_chunk<true, true>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, nullptr, ok);
    }
  } else if (p.d_f32) {
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      if (ncol0 + c0 >= nlimit) break;
      epilogue_chunk<true>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, nullptr, ok);
    }
  } else {
    const __nv_bfloat16* aux_base = p.aux ? p.aux + (long long)t.z * p.aux_batch_stride : nullptr;
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 64) {
      if (ncol0 + c0 >= nlimit) break;
      epilogue_chunk<false>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, aux_base, ok);
    }
  }
}

int gemm_fill_params(const fs2_gemm& g, GemmKP& kp);  // host: validation + everything but the tiling

}  // namespace fs2
