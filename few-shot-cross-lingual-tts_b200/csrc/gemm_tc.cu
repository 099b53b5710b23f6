// Persistent, warp-specialised bf16 GEMM / implicit-GEMM Conv1d / weight-gradient kernel for
// sm_100a: TMA (cp.async.bulk.tensor) feeds a multi-stage shared-memory ring, one thread issues
// tcgen05.mma with fp32 accumulators in TMEM (double buffered so the epilogue of tile i overlaps
// the main loop of tile i+1), four epilogue warps drain TMEM with tcgen05.ld and apply
// bias / ReLU / ReLU-backward / residual-add before storing bf16 or fp32 (optionally atomically).
//
// It is the single tensor-core engine behind every GEMM-shaped op of the FastSpeech2 step
// (include/fs2b200.h: fs2_gemm_bf16).  See DESIGN.md "GEMM engine" for the tile/roofline numbers.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

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

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;  // 8 epilogue warps x (32 rows x 128 B)
  static constexpr int STAGING_BYTES = 8 * 32 * 128;
  static constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int TOTAL = TMEM_PTR_OFF + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

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

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmKP p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + L::TMEM_PTR_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages (256 or 512 columns)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(sbase + L::TMEM_PTR_OFF, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int m0 = t.tm * BM;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), L::STAGE_BYTES);
          const uint32_t sa = sbase + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          if (p.mode == FS2_GEMM_NORMAL) {
            const int tap = kb / p.kb_per_tap;
            const int k0 = (kb - tap * p.kb_per_tap) * BK;
            const int n0 = t.tn * BN;
            const int za = p.a_batched ? t.z / p.a_zdiv : 0;
            const int ia = p.a_inner_base + (t.z % p.a_zdiv) * p.a_zmod_stride;
            const int zb = p.b_batched ? t.z / p.b_zdiv : 0;
            const int ib = p.b_inner_base + (t.z % p.b_zdiv) * p.b_zmod_stride;
            if (!p.a_mn) {
              tma_load_3d(sa, &tmA, full_bar(s), ia + k0, m0 + p.tap_shift0 + tap, za);
            } else {
#pragma unroll
              for (int h = 0; h < BM / 64; ++h)
                tma_load_3d(sa + h * kChunkBytes, &tmA, full_bar(s), ia + m0 + h * 64, k0, za);
            }
            if (!p.b_mn) {
              tma_load_3d(sb, &tmB, full_bar(s), ib + tap * p.b_tap_kstride + k0, n0, zb);
            } else {
#pragma unroll
              for (int h = 0; h < BN / 64; ++h)
                tma_load_3d(sb + h * kChunkBytes, &tmB, full_bar(s),
                            ib + tap * p.b_tap_kstride + n0 + h * 64, k0, zb);
            }
          } else {
            const int g = t.kb0 + kb;
            const int zb = g / p.rb_per_batch;
            const int r0 = (g - zb * p.rb_per_batch) * BK;
            const int tap = t.tn / p.n_tiles_per_tap;
            const int c0 = (t.tn - tap * p.n_tiles_per_tap) * BN;
#pragma unroll
            for (int h = 0; h < BM / 64; ++h)
              tma_load_3d(sa + h * kChunkBytes, &tmA, full_bar(s), p.a_inner_base + m0 + h * 64, r0,
                          zb);
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_3d(sb + h * kChunkBytes, &tmB, full_bar(s), p.b_inner_base + c0 + h * 64,
                          r0 + p.tap_shift0 + tap, zb);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (one thread) =======================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u, b_kstep = p.b_mn ? 16u * 128u : 32u;
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = sbase + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t ad = make_smem_desc(sa + j * a_kstep, a_lbo, 1024u);
            const uint64_t bd = make_smem_desc(sb + j * b_kstep, b_lbo, 1024u);
            umma_f16(tacc, ad, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));  // frees the smem slot when these MMAs retire
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
        if (++as == 2) {
          as = 0;
          aph ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue (8 warps, one TMEM lane = one row per thread; warps w and
    // w+4 share a lane quarter and split the tile's columns in halves) ===========
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int chalf = (warp - 4) >> 2;  // column half of the tile
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int m = t.tm * BM + row;
      int ncol0, nlimit;  // first column of this tile inside its (tap) column space, and its extent
      long long base_off;
      if (p.mode == FS2_GEMM_NORMAL) {
        ncol0 = t.tn * BN;
        nlimit = p.N;
        base_off = (long long)(t.z / p.d_zdiv) * p.d_zdiv_stride +
                   (long long)(t.z % p.d_zdiv) * p.d_zmod_stride;
      } else {
        const int tap = t.tn / p.n_tiles_per_tap;
        ncol0 = (t.tn - tap * p.n_tiles_per_tap) * BN;
        nlimit = p.n_per_tap;
        base_off = (long long)tap * p.d_tap_stride;
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const bool row_ok = (m < p.M) && (t.nkb > 0);
      const long long row_off = base_off + (long long)m * p.ldd;
      const __nv_bfloat16* aux_row =
          p.aux ? p.aux + (long long)t.z * p.aux_batch_stride + (long long)m * p.ld_aux : nullptr;
      if (p.d_col_stride == 1 && p.d_atomic) {
        // split-K weight gradients with unit column stride: coalesced vector reductions
        uint8_t* stg = sgen + L::STAGING_OFF + (warp - 4) * 4096;
        const int m_w0 = t.tm * BM + q * 32;
#pragma unroll 1
        for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
          if (ncol0 + c0 >= nlimit) break;  // warp-uniform
          epilogue_chunk<true, true>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, nullptr,
                                     t.nkb > 0);
        }
      } else if (p.d_col_stride == 1 && !p.d_atomic) {
        // coalesced path (every NORMAL-mode output)
        uint8_t* stg = sgen + L::STAGING_OFF + (warp - 4) * 4096;
        const __nv_bfloat16* aux_base = p.aux ? p.aux + (long long)t.z * p.aux_batch_stride : nullptr;
        const int m_w0 = t.tm * BM + q * 32;
        if (p.d_f32) {
#pragma unroll 1
          for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
            if (ncol0 + c0 >= nlimit) break;  // warp-uniform
            epilogue_chunk<true>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, nullptr,
                                 t.nkb > 0);
          }
        } else {
#pragma unroll 1
          for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 64) {
            if (ncol0 + c0 >= nlimit) break;
            epilogue_chunk<false>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, aux_base,
                                  t.nkb > 0);
          }
        }
      } else
#pragma unroll 1
      for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
        const int n0 = ncol0 + c * 32;
        if (n0 >= nlimit) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
        const bool full = (n0 + 32 <= nlimit);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n0 + j < nlimit) f[j] += __ldg(p.bias + n0 + j);
        }
        if (p.epilogue == FS2_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.epilogue == FS2_EPI_RELU_BWD || p.epilogue == FS2_EPI_ADD_AUX) {
          const bool bwd = p.epilogue == FS2_EPI_RELU_BWD;
          if (full) {
            const uint4* ap = reinterpret_cast<const uint4*>(aux_row + n0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 u = __ldg(ap + g);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const float lo = __uint_as_float(w[h] << 16);
                const float hi = __uint_as_float(w[h] & 0xFFFF0000u);
                float& f0 = f[g * 8 + h * 2];
                float& f1 = f[g * 8 + h * 2 + 1];
                if (bwd) {
                  f0 = lo > 0.f ? f0 : 0.f;
                  f1 = hi > 0.f ? f1 : 0.f;
                } else {
                  f0 += lo;
                  f1 += hi;
                }
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (n0 + j < nlimit) {
                const float a = __bfloat162float(aux_row[n0 + j]);
                f[j] = bwd ? (a > 0.f ? f[j] : 0.f) : f[j] + a;
              }
          }
        }
        // ---- store ----
        if (p.d_col_stride == 1 && full && !p.d_atomic) {
          if (p.d_f32) {
            float4* dp = reinterpret_cast<float4*>(static_cast<float*>(p.d) + row_off + n0);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              dp[g] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
          } else {
            uint4* dp = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.d) + row_off + n0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 2 * h], f[g * 8 + 2 * h + 1]);
                w[h] = *reinterpret_cast<uint32_t*>(&b2);
              }
              dp[g] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        } else {
          for (int j = 0; j < 32; ++j) {
            if (n0 + j >= nlimit) break;
            const long long off = row_off + (long long)(n0 + j) * p.d_col_stride;
            if (p.d_f32) {
              float* dp = static_cast<float*>(p.d) + off;
              if (p.d_atomic)
                atomicAdd(dp, f[j]);
              else
                *dp = f[j];
            } else {
              static_cast<__nv_bfloat16*>(p.d)[off] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) {
        as = 0;
        aph ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long long inner, long long rows, long long batches,
                      long long ld, long long batch_stride, int box_inner, int box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym)
      return set_error("cuTensorMapEncodeTiled driver entry point not available");
    enc = reinterpret_cast<EncodeTiledFn>(sym);
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7) || (batch_stride & 7))
    return set_error("TMA operand must be 16-byte aligned with ld / batch_stride multiples of 8 elements");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t bs = batches > 1 ? (cuuint64_t)batch_stride : (cuuint64_t)ld * rows;
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld batches=%lld ld=%lld bs=%lld",
             (int)r, inner, rows, batches, ld, batch_stride);
    return set_error(buf);
  }
  return 0;
}

static int make_tmap(CUtensorMap* map, const fs2_operand& o, int box_inner, int box_rows) {
  return make_tmap_bf16_3d(map, o.ptr, o.inner, o.rows, o.batches, o.ld, o.batch_stride, box_inner, box_rows);
}

static int g_num_sms = 0;

template <int BN, int STAGES>
static int launch_tc(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm_tc)", e);
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap(&tmA, g.a, 64, g.a.mn_major ? 64 : BM)) return rc;
  if (int rc = make_tmap(&tmB, g.b, 64, g.b.mn_major ? 64 : BN)) return rc;
  kp.tiles_n = (g.mode == FS2_GEMM_NORMAL) ? (g.N + BN - 1) / BN
                                           : g.taps * ((g.N + BN - 1) / BN);
  kp.n_tiles_per_tap = (g.N + BN - 1) / BN;
  kp.total_tiles = kp.tiles_m * kp.tiles_n * kp.Z;
  if (kp.total_tiles <= 0) return 0;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = kp.total_tiles < g_num_sms ? kp.total_tiles : g_num_sms;
  gemm_tc_kernel<BN, STAGES><<<grid, kThreads, L::DYN_BYTES, stream>>>(tmA, tmB, kp);
  count_launch();
  return check_launch("gemm_tc_kernel");
}

int gemm_tc_launch(const fs2_gemm& g, cudaStream_t stream) {
  GemmKP kp{};
  kp.mode = g.mode;
  kp.M = g.M;
  kp.N = g.N;
  kp.a_mn = g.a.mn_major;
  kp.b_mn = g.b.mn_major;
  kp.a_inner_base = g.a.inner_base;
  kp.a_zdiv = g.a.zdiv > 0 ? g.a.zdiv : 1;
  kp.a_zmod_stride = g.a.zmod_stride;
  kp.a_batched = g.a.batches > 1;
  kp.b_inner_base = g.b.inner_base;
  kp.b_zdiv = g.b.zdiv > 0 ? g.b.zdiv : 1;
  kp.b_zmod_stride = g.b.zmod_stride;
  kp.b_batched = g.b.batches > 1;
  kp.tap_shift0 = g.tap_shift0;
  kp.b_tap_kstride = g.b_tap_kstride;
  kp.tiles_m = (g.M + BM - 1) / BM;
  kp.epilogue = g.epilogue;
  kp.d_f32 = g.d_f32;
  kp.d_atomic = g.d_atomic;
  kp.d_zdiv = g.d_zdiv > 0 ? g.d_zdiv : 1;
  kp.alpha = g.alpha;
  kp.d = g.d;
  kp.ldd = g.ldd;
  kp.d_col_stride = g.d_col_stride > 0 ? g.d_col_stride : 1;
  kp.d_tap_stride = g.d_tap_stride;
  kp.d_zdiv_stride = g.d_zdiv_stride;
  kp.d_zmod_stride = g.d_zmod_stride;
  kp.bias = g.bias;
  kp.aux = static_cast<const __nv_bfloat16*>(g.aux);
  kp.ld_aux = g.ld_aux;
  kp.aux_batch_stride = g.aux_batch_stride;
  kp.seg_rows = g.d_seg_rows;
  for (int i = 0; i < 4; ++i) kp.seg[i] = g.d_seg[i];
  const int taps = g.taps > 0 ? g.taps : 1;
  if (g.mode == FS2_GEMM_NORMAL) {
    if (g.a.mn_major && taps != 1) return set_error("conv taps need a K-major A operand");
    kp.Z = g.Z > 0 ? g.Z : 1;
    kp.kb_per_tap = (g.K + BK - 1) / BK;
    kp.num_kb = kp.kb_per_tap * taps;
    if (kp.num_kb <= 0) return set_error("gemm: empty reduction");
  } else if (g.mode == FS2_GEMM_WGRAD) {
    if (!g.a.mn_major || !g.b.mn_major) return set_error("WGRAD needs MN-major operands");
    if (!g.d_f32) return set_error("WGRAD output must be f32");
    kp.rb_per_batch = (g.a.rows + BK - 1) / BK;
    kp.total_rb = kp.rb_per_batch * g.a.batches;
    int splits = g.splits > 0 ? g.splits : 1;
    if (splits > kp.total_rb) splits = kp.total_rb;
    if (splits < 1) return set_error("wgrad: empty reduction");
    kp.kb_per_split = (kp.total_rb + splits - 1) / splits;
    splits = (kp.total_rb + kp.kb_per_split - 1) / kp.kb_per_split;  // no empty split
    if (splits > 1 && !g.d_atomic) return set_error("split-K wgrad needs d_atomic=1");
    kp.Z = splits;
    kp.n_per_tap = g.N;
  } else {
    return set_error("gemm: unknown mode");
  }
  if (g.d_atomic && !g.d_f32) return set_error("atomic accumulation needs an f32 output");
  if (g.d_seg_rows > 0 && !(g.mode == FS2_GEMM_WGRAD && kp.d_col_stride == 1 && g.d_atomic))
    return set_error("gemm: row segments are supported for unit-stride atomic WGRAD outputs only");
  if (g.d_seg_rows > 0 && (g.M + g.d_seg_rows - 1) / g.d_seg_rows > 4) return set_error("gemm: at most 4 segments");
  if (kp.d_col_stride == 1) {
    const long long al = g.d_f32 ? 4 : 8;  // vector stores / reductions: 16-byte aligned rows
    if ((reinterpret_cast<uintptr_t>(g.d) & 15) || (g.ldd % al) || (g.d_zdiv_stride % al) ||
        (g.d_zmod_stride % al) || (g.d_tap_stride % al))
      return set_error("gemm: output rows must be 16-byte aligned (ldd / strides)");
  }
  if (g.aux && (g.d_f32 || (g.N & 7))) return set_error("gemm: aux epilogues need a bf16 output and N % 8 == 0");
  if (g.aux && ((reinterpret_cast<uintptr_t>(g.aux) & 15) || (g.ld_aux & 7) || (g.aux_batch_stride & 7)))
    return set_error("gemm: aux rows must be 16-byte aligned");
  if ((g.epilogue == FS2_EPI_RELU_BWD || g.epilogue == FS2_EPI_ADD_AUX) && !g.aux)
    return set_error("epilogue needs aux");

  // Tile-N choice: 256 columns (128x256 is the shape that can reach the tensor-pipe peak from one
  // CTA) unless the 128-wide tiling wastes clearly less of a ragged N.
  const int n = g.N;
  const int pad256 = ((n + 255) / 256) * 256, pad128 = ((n + 127) / 128) * 128;
  const bool use128 = (n <= 128) || (pad128 * 100 < pad256 * 92);
  if (use128) return launch_tc<128, 6>(g, kp, stream);
  return launch_tc<256, 4>(g, kp, stream);
}

}  // namespace fs2
