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
  // ragged rows (fs2_gemm::row_lens): padded rows are zero / skipped
  const long long* row_lens;
  int lens_zdiv;
  int ragged;      // 1: compact schedule over the prefix table (sched_n <= kMaxRaggedZ entries)
  int sched_n;     // table entries: NORMAL batches Z, WGRAD reduction batches
  int unit_rows;   // rows per schedule unit: NORMAL BM (1-CTA) / 2*BM (CTA pair); WGRAD BK
  int units_max;   // units per entry when nothing is padded
  int row_extent;  // rows per entry (NORMAL: M, WGRAD: a.rows)
  int tail_zero;   // NORMAL: rows zeroed behind the last scheduled tile (0 = all, < 0 = none)
  unsigned long long* relu_mask;  // optional 1-bit ReLU mask [Z*M][N/64] (written by EPI_RELU, read by EPI_RELU_BWD)
  int dbg;         // FS2_GEMM_DBG ablation bits (tools only): 1 no global stores, 2 no epilogue math, 4 no TMA, 8 no MMA
  int pair_any;    // 2-CTA kernels, ragged NORMAL, shared (un-batched) B: the two CTAs of a pair take ANY two
                   // consecutive 128-row tiles of the compact list, also from different utterances
  float* ws;       // conv_tc2: split-K scratch for the last partial wave (fs2_gemm::workspace), NULL = off
  float* a_colsum; // wgrad_taps / gemm_tc2 WGRAD: += column sums of A over the reduction (fs2_gemm::a_colsum)
  float* a_colsum_seg[4];  // ... per output segment (seg_rows > 0), NULL entries skipped
  int a_colsum_on;         // any of the above set: the leader CTA's epilogue warps read the A tiles (empty count 5)
  // gemm_tc2<LN>: fused dropout + residual + LayerNorm + pad-zero epilogue (fs2_gemm::ln_*), ln_gamma != NULL
  const float* ln_gamma;
  const float* ln_beta;
  const __nv_bfloat16* ln_res;
  long long ld_res, res_batch_stride;
  float ln_p_drop;
  unsigned long long ln_seed;
  const unsigned long long* ln_seed_dev;
  __nv_bfloat16* ln_v;
  float* ln_mean;
  float* ln_rstd;
  unsigned char* ln_keep;
};

constexpr int kMaxRaggedZ = 256;  // prefix table lives in the ~1.9 KiB of shared memory left by the smem ring

// cum[z] = inclusive prefix sum of the schedule units (row tiles / reduction blocks) that contain at least one
// valid row of entry z.  One warp builds it with shuffles.
__device__ __forceinline__ void build_ragged_table(const GemmKP& p, int* cum, int lane) {
  int carry = 0;
  for (int z0 = 0; z0 < p.sched_n; z0 += 32) {
    const int z = z0 + lane;
    int u = 0;
    if (z < p.sched_n) {
      long long len = p.row_lens[z / p.lens_zdiv];
      len = len < 0 ? 0 : (len > p.row_extent ? p.row_extent : len);
      u = (int)((len + p.unit_rows - 1) / p.unit_rows);
      if (u > p.units_max) u = p.units_max;
    }
    int v = u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (z < p.sched_n) cum[z] = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
}

// first entry z with cum[z] > r  (r < cum[n-1])
__device__ __forceinline__ int ragged_find(const int* cum, int n, int r) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] > r) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// number of tiles of the persistent loop (1-CTA: 128-row tiles; pair kernel: pair tiles, `dense` = its dense count)
__device__ __forceinline__ int sched_total(const GemmKP& p, const int* cum, int dense) {
  if (p.ragged && p.mode == FS2_GEMM_NORMAL) return cum[p.sched_n - 1] * p.tiles_n;
  return dense;
}

// WGRAD split-K range [kb0, kb0 + nkb) of split `z` over the (compacted) list of reduction blocks
__device__ __forceinline__ void wgrad_range(const GemmKP& p, const int* cum, int z, int& kb0, int& nkb) {
  int total = p.total_rb, per = p.kb_per_split;
  if (p.ragged) {
    total = cum[p.sched_n - 1];
    per = (total + p.Z - 1) / p.Z;
  }
  kb0 = z * per;
  const int rem = total - kb0;
  nkb = rem < per ? rem : per;
  if (nkb < 0) nkb = 0;
}

// compact reduction-block index g -> (batch zb, first row r0 of the 64-row block)
struct RbCursor {
  int zb, lb, nb;  // batch, local block, blocks of this batch
};
__device__ __forceinline__ RbCursor rb_seek(const GemmKP& p, const int* cum, int g) {
  RbCursor c;
  if (!p.ragged) {
    c.zb = g / p.rb_per_batch;
    c.lb = g - c.zb * p.rb_per_batch;
    c.nb = p.rb_per_batch;
  } else {
    c.zb = ragged_find(cum, p.sched_n, g);
    const int before = c.zb ? cum[c.zb - 1] : 0;
    c.lb = g - before;
    c.nb = cum[c.zb] - before;
  }
  return c;
}
__device__ __forceinline__ void rb_next(const GemmKP& p, const int* cum, RbCursor& c) {
  if (++c.lb < c.nb) return;
  c.lb = 0;
  ++c.zb;
  if (p.ragged) {
    while (c.zb < p.sched_n && cum[c.zb] == cum[c.zb - 1]) ++c.zb;  // skip fully padded batches
    c.nb = c.zb < p.sched_n ? cum[c.zb] - cum[c.zb - 1] : 1;
  }
}

struct TileCoord {
  int z, tm, tn, nkb, kb0;
  int valid = 1;  // 0: filler half of an odd pair -- operands are loaded and multiplied, nothing is stored
};

// Tile of CTA `rank` of pair tile `ptile` in the 2-CTA kernels (NORMAL mode).
__device__ __forceinline__ TileCoord decode_pair_normal(const GemmKP& p, const int* cum, int ptile, int rank,
                                                        int pair_tiles_m) {
  TileCoord t;
  t.tn = ptile % p.tiles_n;
  const int r = ptile / p.tiles_n;
  if (p.pair_any) {  // compact list of 128-row tiles, two consecutive entries per pair
    const int total = cum[p.sched_n - 1];
    int idx = 2 * r + rank;
    if (idx >= total) {
      idx = 2 * r;
      t.valid = 0;
    }
    t.z = ragged_find(cum, p.sched_n, idx);
    t.tm = idx - (t.z ? cum[t.z - 1] : 0);
  } else {
    int pm;
    if (p.ragged) {  // r-th 256-row pair tile that holds at least one valid row
      t.z = ragged_find(cum, p.sched_n, r);
      pm = r - (t.z ? cum[t.z - 1] : 0);
    } else {
      pm = r % pair_tiles_m;
      t.z = r / pair_tiles_m;
    }
    t.tm = 2 * pm + rank;  // this CTA's 128-row tile of the pair tile
  }
  t.nkb = p.num_kb;
  t.kb0 = 0;
  return t;
}
__device__ __forceinline__ int pair_sched_total(const GemmKP& p, const int* cum, int dense) {
  if (p.mode == FS2_GEMM_NORMAL && p.pair_any) return ((cum[p.sched_n - 1] + 1) >> 1) * p.tiles_n;
  return sched_total(p, cum, dense);
}

__device__ __forceinline__ TileCoord decode_tile(const GemmKP& p, const int* cum, int tile) {
  TileCoord t;
  t.tn = tile % p.tiles_n;
  int r = tile / p.tiles_n;
  if (p.mode == FS2_GEMM_NORMAL) {
    if (p.ragged) {  // r-th row tile that holds at least one valid row
      t.z = ragged_find(cum, p.sched_n, r);
      t.tm = r - (t.z ? cum[t.z - 1] : 0);
    } else {
      t.tm = r % p.tiles_m;
      t.z = r / p.tiles_m;
    }
    t.nkb = p.num_kb;
    t.kb0 = 0;
  } else {
    t.tm = r % p.tiles_m;
    t.z = r / p.tiles_m;  // split index
    wgrad_range(p, cum, t.z, t.kb0, t.nkb);
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
                                               const __nv_bfloat16* aux_base, bool tile_ok, bool row_ok = true,
                                               unsigned long long* mask_row = nullptr) {
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
    if (mask_row) {  // 1-bit ReLU mask of this thread's row, word n0 / 64 (this thread owns the whole 64-column chunk)
      unsigned long long* mw = mask_row + (n0 >> 6);
      if (p.epilogue == FS2_EPI_RELU) {
        unsigned int lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          lo |= (f[j] > 0.f ? 1u : 0u) << j;
          hi |= (f[32 + j] > 0.f ? 1u : 0u) << j;
        }
        if (tile_ok) *mw = (static_cast<unsigned long long>(hi) << 32) | lo;
      } else if (p.epilogue == FS2_EPI_RELU_BWD) {
        const unsigned long long m = tile_ok ? *mw : 0ull;
        const unsigned int lo = static_cast<unsigned int>(m), hi = static_cast<unsigned int>(m >> 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = (lo >> j) & 1u ? f[j] : 0.f;
          f[32 + j] = (hi >> j) & 1u ? f[32 + j] : 0.f;
        }
      }
    }
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
  if (!row_ok) {  // padded row (fs2_gemm::row_lens): the output is defined to be zero
#pragma unroll
    for (int j = 0; j < NC; ++j) f[j] = 0.f;
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
  const bool ok = t.nkb > 0 && t.valid && !(p.dbg & 1);
  bool row_ok = true;
  if (p.row_lens && p.mode == FS2_GEMM_NORMAL) row_ok = (m_w0 + lane) < p.row_lens[t.z / p.lens_zdiv];
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
      epilogue_chunk<true>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, nullptr, ok, row_ok);
    }
  } else {
    const __nv_bfloat16* aux_base = p.aux ? p.aux + (long long)t.z * p.aux_batch_stride : nullptr;
    unsigned long long* mask_row = nullptr;
    bool ok_row = ok;
    if (p.relu_mask) {  // this thread's row of the 1-bit mask (rows outside the output are neither read nor written)
      const int gm = m_w0 + lane;
      mask_row = p.relu_mask + ((long long)t.z * p.M + (gm < p.M ? gm : 0)) * (p.N >> 6);
      ok_row = ok && gm < p.M;
    }
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 64) {
      if (ncol0 + c0 >= nlimit) break;
      epilogue_chunk<false>(p, taddr + c0, stg, lane, m_w0, ncol0 + c0, nlimit, base_off, aux_base, ok, row_ok,
                            mask_row ? (ok_row ? mask_row : nullptr) : nullptr);
    }
  }
}

// Ragged NORMAL outputs: rows of D[z] past the last scheduled row tile are written as zero by the epilogue
// warps (256 threads, thread index `et`) of all CTAs before their first tile's accumulator is ready.
template <int BN>
__device__ __forceinline__ void zero_fill_padded(const GemmKP& p, const int* cum, int et, int cta, int ncta) {
  if (p.tail_zero < 0) return;
  const int epv = p.d_f32 ? 4 : 8;
  for (int item = cta; item < p.sched_n * p.tiles_n; item += ncta) {
    const int z = item / p.tiles_n, tn = item - z * p.tiles_n;
    const int row0 = (cum[z] - (z ? cum[z - 1] : 0)) * p.unit_rows;
    int row1 = p.M;
    if (p.tail_zero > 0 && row0 + p.tail_zero < row1) row1 = row0 + p.tail_zero;
    if (row0 >= row1) continue;
    const int c0 = tn * BN, c1 = (c0 + BN < p.N) ? c0 + BN : p.N;
    const int vpr = (c1 - c0 + epv - 1) / epv;
    const long long base = (long long)(z / p.d_zdiv) * p.d_zdiv_stride + (long long)(z % p.d_zdiv) * p.d_zmod_stride;
    const int nvec = (row1 - row0) * vpr;  // < 2^31: one tile column of one utterance
    for (int i = et; i < nvec; i += 256) {
      const int r = i / vpr, v = i - r * vpr;
      const int col = c0 + v * epv;
      const long long off = base + (long long)(row0 + r) * p.ldd + col;
      if (col + epv <= c1) {
        if (p.d_f32) *reinterpret_cast<uint4*>(static_cast<float*>(p.d) + off) = make_uint4(0u, 0u, 0u, 0u);
        else *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.d) + off) = make_uint4(0u, 0u, 0u, 0u);
      } else {
        for (int e = 0; col + e < c1; ++e) {
          if (p.d_f32) static_cast<float*>(p.d)[off + e] = 0.f;
          else reinterpret_cast<uint16_t*>(p.d)[off + e] = 0;
        }
      }
    }
  }
}

int gemm_fill_params(const fs2_gemm& g, GemmKP& kp);  // host: validation + everything but the tiling

}  // namespace fs2
