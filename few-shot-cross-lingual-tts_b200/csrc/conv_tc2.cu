// 2-CTA (cta_group::2) implicit-GEMM Conv1d with OPERAND-HALO REUSE: the tensor-core engine of the k = 9
// FFN convolution (forward and input gradient), the k = 3 predictor and the k = 5 PostNet convolutions.
//
// Why a second conv path: the plain implicit GEMM (gemm_tc.cu) reloads the activation tile once per tap
// (rows m0 + tap - pad), i.e. 9x for k = 9, and on B200 the L2 -> SM fabric (~6300 B/clk for the whole chip
// = ~43 B/clk/SM, B300_MICROARCH.md "LTS throughput cap") is what bounds a 128 x 256 tile: 48 KiB of operands
// per 512 MMA cycles = 96 B/clk.  ncu on that kernel: tensor pipe 52 % active, DRAM 5 % (profiles/).
// Here, per 64-channel block of the reduction:
//   * ONE TMA box of (128 + taps - 1) activation rows lands in shared memory (per CTA: its own 128 rows
//     + the halo); every tap reads it through a shared-memory descriptor whose start address is shifted
//     by `tap` rows (128 B each) -- with SWIZZLE_128B the XOR pattern follows the absolute address bits,
//     so a row-shifted view of the same bytes is a valid K-major operand;
//   * per tap only the weight tile streams in, and with cta_group::2 each CTA loads HALF of its 256 rows.
// L2 traffic per CTA and 64 x tap block: 16 KiB (weights) + 17 KiB / taps (activations) ~ 18 KiB for the
// same 128 x 256 x 64 MACs per CTA: 2.6x less than the 1-CTA kernel, below the fabric limit.
//
// Scheduling, epilogue, ragged rows (fs2_gemm::row_lens) are shared with gemm_tc2.cu (gemm_common.cuh).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"
#include "ptx2sm.cuh"
#include "tmap.h"

namespace fs2 {

namespace c2 {
constexpr int BN_MAX = 256;
constexpr int MAX_TAPS = 16;
constexpr int A_BUFS = 2;
constexpr int A_BUF_BYTES = 18 * 1024;        // (128 + MAX_TAPS - 1) rows x 128 B = 18304 B, 1 KiB aligned
constexpr int B_STAGES = 8;
constexpr int B_BYTES = (BN_MAX / 2) * BK * 2; // 16 KiB stage: this CTA's half of the (up to) 256 weight rows
constexpr int B_OFF = A_BUFS * A_BUF_BYTES;
constexpr int STAGING_OFF = B_OFF + B_STAGES * B_BYTES;
constexpr int STAGING_BYTES = 8 * 32 * 128;
constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;
constexpr int NUM_BARS = 2 * A_BUFS + 2 * B_STAGES + 4;
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int CUM_OFF = TMEM_PTR_OFF + 16;
constexpr int DYN_BYTES = CUM_OFF + kMaxRaggedZ * 4 + 1024;
constexpr int kThreadsC2 = 384;
}  // namespace c2

// Row-shifted operand views: the start address of the A descriptor is base + tap * 128 B, i.e. NOT aligned to
// the 1 KiB swizzle atom.  Measured on B200 (tests/test_gemm_gpu.py::test_conv1d_channels_last, test_ragged_conv):
// the descriptor's "base offset" field must stay 0 -- tcgen05.mma applies the 128-byte-swizzle XOR to the
// absolute shared-memory address bits, exactly like TMA when it wrote the tile; filling in (start >> 7) & 7
// produces wrong results.


// ------------------------------------------------------------------------------------------------------------
// Split-K tail.  The persistent loop deals `total` pair tiles to `P` CTA pairs: the last round is usually
// partial (k = 9 input gradient, ragged C2: 157 pair tiles on 74 pairs -> rounds of 74, 74, 9; the third round
// costs a full tile time with 12 % of the machine busy).  With a workspace (fs2_gemm::workspace) the `rem` tiles
// of that round are cut into s = 2 or 4 slices of the channel-block loop (every slice keeps ALL taps, so the
// halo reuse is untouched) and dealt to rem * s pairs.  Each CTA drains its partial accumulator to a slab of the
// workspace and bumps a counter; the CTA that arrives LAST adds the s slabs in slice order (fixed order ->
// bit-reproducible) and runs the usual epilogue math on the sums.  Nobody waits for anybody: when a concurrent
// kernel of the weight-gradient stream delays some CTA pairs, the early ones simply leave.
// Only long reductions are split (>= 64 k-blocks per tile): the slab round trip costs about as much as 20 k-blocks.
// ------------------------------------------------------------------------------------------------------------
namespace c2 {
constexpr int WS_HEADER_BYTES = 4096;             // counters: [tile of the tail][rank] ints
constexpr int WS_SLAB_FLOATS = BM * BN_MAX;       // one CTA's 128 x 256 partial accumulator
constexpr int WS_MAX_TAIL_TILES = WS_HEADER_BYTES / 8;
}  // namespace c2

struct TailSched {
  int full;   // pair tiles handled by the plain persistent loop
  int s;      // slices per tail tile (1 = no split)
  int units;  // rem * s: pairs [0, units) take one slice each
};
__device__ __forceinline__ TailSched tail_sched(const GemmKP& p, int total, int P) {
  TailSched t{total, 1, 0};
  if (!p.ws || total <= P) return t;
  const int rem = total % P;
  if (rem == 0 || 2 * rem > P || rem > c2::WS_MAX_TAIL_TILES) return t;
  int s = P / rem >= 4 ? 4 : 2;
  while (s > p.kb_per_tap) s >>= 1;
  if (s < 2 || p.num_kb < (s == 4 ? 64 : 96)) return t;
  t.full = total - rem;
  t.s = s;
  t.units = rem * s;
  return t;
}

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Epilogue of one split unit (8 epilogue warps = 256 threads `et` of ONE CTA; BN = 256, bf16 D).
//   tile_r: index of the tile inside the tail, j: this unit's slice, s: slices per tile.
__device__ __forceinline__ void epilogue_split(const GemmKP& p, const TileCoord& t, uint32_t tmem_acc, int rank,
                                               int tile_r, int j, int s, int et, uint32_t tempty_remote,
                                               volatile int* last_flag) {
  using namespace c2;
  const int warp = et >> 5, lane = et & 31;
  const int q = warp & 3, chalf = warp >> 2;
  int* cnt = reinterpret_cast<int*>(p.ws) + (tile_r * 2 + rank);
  float* slabs = p.ws + WS_HEADER_BYTES / 4 + (size_t)(tile_r * s) * 2 * WS_SLAB_FLOATS + (size_t)rank * WS_SLAB_FLOATS;
  const size_t slab_stride = 2 * (size_t)WS_SLAB_FLOATS;  // slice j -> j + 1
  const bool active = t.valid != 0;  // the filler half of an odd pair takes no part (uniform over the s units)
  // ---- A: TMEM -> this unit's slab, layout [column / 4][row][4] (coalesced both here and in C)
  if (active) {
    float4* my = reinterpret_cast<float4*>(slabs + (size_t)j * slab_stride);
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const int row = q * 32 + lane;
#pragma unroll 1
    for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        __stcg(my + (size_t)(c0 / 4 + i) * BM + row,
               make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                           __uint_as_float(v[4 * i + 3])));
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_remote(tempty_remote);
  if (!active) return;
  // ---- B: count this unit in; only the last of the s units of this (tile, rank) goes on
  __threadfence();
  bar_sync_epi();
  if (et == 0) {
    const int last = atomicAdd(cnt, 1) == s - 1;
    if (last) {
      __threadfence();
      *cnt = 0;  // every unit has been counted: ready for the next launch
    }
    *last_flag = last;
  }
  bar_sync_epi();
  if (!*last_flag) return;
  // ---- C: sum the slabs in slice order, 32 columns at a time per thread (row = thread % 128)
  const int row = et & 127;
  const int gm = t.tm * BM + row;
  constexpr int groups = BN_MAX / 32;
  const bool row_in = gm < p.M;
  bool row_ok = true;
  if (p.row_lens) row_ok = gm < p.row_lens[t.z / p.lens_zdiv];
  const long long base_off =
      (long long)(t.z / p.d_zdiv) * p.d_zdiv_stride + (long long)(t.z % p.d_zdiv) * p.d_zmod_stride;
  const __nv_bfloat16* aux_row = p.aux ? p.aux + (long long)t.z * p.aux_batch_stride + (long long)gm * p.ld_aux : nullptr;
  uint32_t* mask_row = p.relu_mask
                           ? reinterpret_cast<uint32_t*>(p.relu_mask) + ((long long)t.z * p.M + gm) * (p.N >> 5)
                           : nullptr;
  const bool store_ok = t.nkb > 0 && !(p.dbg & 1);
  for (int g = et >> 7; g < groups; g += 2) {
    const int cl = g * 32;                     // first column inside the tile
    const int col0 = t.tn * BN_MAX + cl;       // ... inside D
    if (col0 >= p.N) break;
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = 0.f;
    for (int jj = 0; jj < s; jj += 2) {  // s is even: two slabs = sixteen 16-byte loads in flight per thread
      const float4* src = reinterpret_cast<const float4*>(slabs + (size_t)jj * slab_stride) + (size_t)(cl / 4) * BM + row;
      const float4* src2 = reinterpret_cast<const float4*>(slabs + (size_t)(jj + 1) * slab_stride) + (size_t)(cl / 4) * BM + row;
      float4 v[8], v2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = __ldcg(src + (size_t)i * BM);
        v2[i] = __ldcg(src2 + (size_t)i * BM);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[4 * i] = (f[4 * i] + v[i].x) + v2[i].x;
        f[4 * i + 1] = (f[4 * i + 1] + v[i].y) + v2[i].y;
        f[4 * i + 2] = (f[4 * i + 2] + v[i].z) + v2[i].z;
        f[4 * i + 3] = (f[4 * i + 3] + v[i].w) + v2[i].w;
      }
    }
    if (!row_in) continue;
    if (p.alpha != 1.f) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] *= p.alpha;
    }
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) f[i] += __ldg(p.bias + col0 + i);
    }
    if (p.epilogue == FS2_EPI_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (mask_row) {
      uint32_t* mw = mask_row + (col0 >> 5);
      if (p.epilogue == FS2_EPI_RELU) {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) m |= (f[i] > 0.f ? 1u : 0u) << i;
        if (store_ok) *mw = m;
      } else if (p.epilogue == FS2_EPI_RELU_BWD) {
        const uint32_t m = store_ok ? *mw : 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = (m >> i) & 1u ? f[i] : 0.f;
      }
    }
    if (aux_row) {
      const bool bwd = p.epilogue == FS2_EPI_RELU_BWD;
#pragma unroll
      for (int v8 = 0; v8 < 4; ++v8) {
        if (col0 + v8 * 8 >= p.N) break;  // N % 8 == 0 with aux (checked on the host)
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(aux_row + col0 + v8 * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float lo = __uint_as_float(w[h] << 16), hi = __uint_as_float(w[h] & 0xFFFF0000u);
          float& f0 = f[v8 * 8 + 2 * h];
          float& f1 = f[v8 * 8 + 2 * h + 1];
          if (bwd) {
            f0 = lo > 0.f ? f0 : 0.f;
            f1 = hi > 0.f ? f1 : 0.f;
          } else {
            f0 += lo;
            f1 += hi;
          }
        }
      }
    }
    if (!row_ok) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = 0.f;
    }
    if (!store_ok) continue;
    __nv_bfloat16* drow = static_cast<__nv_bfloat16*>(p.d) + base_off + (long long)gm * p.ldd + col0;
#pragma unroll
    for (int v8 = 0; v8 < 4; ++v8) {
      const int c = col0 + v8 * 8;
      if (c + 8 <= p.N) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          __nv_bfloat162 b2 = __floats2bfloat162_rn(f[v8 * 8 + 2 * h], f[v8 * 8 + 2 * h + 1]);
          w[h] = *reinterpret_cast<uint32_t*>(&b2);
        }
        *reinterpret_cast<uint4*>(drow + v8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        for (int e = 0; e < 8 && c + e < p.N; ++e) drow[v8 * 8 + e] = __float2bfloat16_rn(f[v8 * 8 + e]);
      }
    }
  }
}

// BN = 256: the default pair tile (256 rows x 256 columns).  BN = 128: outputs that are only 256 columns wide (the
// k = 9 input gradient) -- half-size tiles cut the loss of the last partial wave (157 pair tiles on 74 CTA pairs are
// 2.12 waves = 3 rounds; 314 half tiles are 4.24 waves = 5 half rounds).
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(c2::kThreadsC2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmKP p, const int taps) {
  pdl_trigger();
  using namespace c2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + BAR_OFF;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (A_BUFS + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * A_BUFS + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * A_BUFS + B_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + TMEM_PTR_OFF);
  const int* cum = reinterpret_cast<const int*>(sgen + CUM_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int B_TILE_BYTES = (BN / 2) * BK * 2;  // bytes this CTA loads per (tap, 64-channel block)
  const uint32_t a_box_bytes = static_cast<uint32_t>(BM + taps - 1) * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_BUFS; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_2sm(sbase + TMEM_PTR_OFF, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  pdl_wait();  // everything above is independent of the previous kernel's output
  if (warp == 3 && p.ragged) build_ragged_table(p, reinterpret_cast<int*>(sgen + CUM_OFF), lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int pair_tiles_m = (p.tiles_m + 1) >> 1;
  const int total_pair_tiles = pair_sched_total(p, cum, pair_tiles_m * p.tiles_n * p.Z);
  auto decode_pair = [&](int ptile) { return decode_pair_normal(p, cum, ptile, (int)rank, pair_tiles_m); };
  // split-K tail of the last partial round (BN = 256 only): pair u < ts.units takes slice u % ts.s of tail tile u / ts.s
  const TailSched ts = BN == 256 ? tail_sched(p, total_pair_tiles, num_pairs) : TailSched{total_pair_tiles, 1, 0};
  const bool has_tail = pair < ts.units;
  const int tail_tile = ts.full + pair / ts.s, tail_j = pair % ts.s;
  const int tail_c0 = has_tail ? (tail_j * p.kb_per_tap) / ts.s : 0;
  const int tail_c1 = has_tail ? ((tail_j + 1) * p.kb_per_tap) / ts.s : 0;
  const int n_iter = (pair < ts.full ? (ts.full - pair + num_pairs - 1) / num_pairs : 0) + (has_tail ? 1 : 0);

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      const uint32_t leader_afull0 = mapa_rank(afull_bar(0), 0);
      const uint32_t leader_bfull0 = mapa_rank(bfull_bar(0), 0);
      int sa_i = 0, sb_i = 0;
      uint32_t pha = 0, phb = 0;
      for (int it = 0, ptile = pair; it < n_iter; ++it, ptile += num_pairs) {
        const bool tail = has_tail && it == n_iter - 1;
        const TileCoord t = decode_pair(tail ? tail_tile : ptile);
        const int m0 = t.tm * BM;
        const int n0 = t.tn * BN + (int)rank * (BN / 2);
        const int za = p.a_batched ? t.z / p.a_zdiv : 0;
        const int ia = p.a_inner_base + (t.z % p.a_zdiv) * p.a_zmod_stride;
        const int zb = p.b_batched ? t.z / p.b_zdiv : 0;
        const int ib = p.b_inner_base + (t.z % p.b_zdiv) * p.b_zmod_stride;
        const int c_begin = tail ? tail_c0 : 0, c_end = tail ? tail_c1 : p.kb_per_tap;
        for (int c = c_begin; c < c_end; ++c) {
          const int k0 = c * BK;
          // activation rows [m0 + shift0, m0 + shift0 + 128 + taps - 1) x 64 channels: once for all taps
          mbar_wait(aempty_bar(sa_i), pha ^ 1u);
          if (leader) mbar_arrive_expect_tx(afull_bar(sa_i), 2 * a_box_bytes);
          tma_load_3d_2sm(sbase + sa_i * A_BUF_BYTES, &tmA, leader_afull0 + 8u * sa_i, ia + k0, m0 + p.tap_shift0,
                          za);
          if (++sa_i == A_BUFS) {
            sa_i = 0;
            pha ^= 1u;
          }
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(bempty_bar(sb_i), phb ^ 1u);
            if (leader) mbar_arrive_expect_tx(bfull_bar(sb_i), 2 * B_TILE_BYTES);
            const uint32_t lfull = leader_bfull0 + 8u * sb_i;
            const uint32_t sb = sbase + B_OFF + sb_i * B_BYTES;
            if (!p.b_mn) {
              tma_load_3d_2sm(sb, &tmB, lfull, ib + tap * p.b_tap_kstride + k0, n0, zb);
            } else {
#pragma unroll
              for (int h = 0; h < BN / 128; ++h)
                tma_load_3d_2sm(sb + h * kChunkBytes, &tmB, lfull, ib + tap * p.b_tap_kstride + n0 + h * 64, k0,
                                zb);
            }
            if (++sb_i == B_STAGES) {
              sb_i = 0;
              phb ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) ========
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, p.b_mn);
      const uint32_t b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t b_kstep = p.b_mn ? 16u * 128u : 32u;
      // descriptors: constant high word; low word = (address >> 4) | (LBO >> 4) << 16 -> offsets are added >> 4
      const uint64_t ad0 = make_smem_desc(sbase, 16u, 1024u);
      const uint64_t bd0 = make_smem_desc(sbase + B_OFF, b_lbo, 1024u);
      int sa_i = 0, sb_i = 0, as = 0;
      uint32_t pha = 0, phb = 0, aph = 0;
      for (int it = 0; it < n_iter; ++it) {
        const bool tail = has_tail && it == n_iter - 1;
        const int c_begin = tail ? tail_c0 : 0, c_end = tail ? tail_c1 : p.kb_per_tap;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int c = c_begin; c < c_end; ++c) {
          mbar_wait(afull_bar(sa_i), pha);
          tc_fence_after();
          const uint64_t ad_c = ad0 + (uint64_t)((sa_i * A_BUF_BYTES) >> 4);
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(bfull_bar(sb_i), phb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bd_s = bd0 + (uint64_t)((sb_i * B_BYTES) >> 4);
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                // A: rows [tap, tap + 128) of the halo tile = the same bytes, start shifted by tap * 128 B
                umma_f16_2sm(tacc, ad_c + (uint64_t)((tap * 128 + j * 32) >> 4), bd_s + (uint64_t)((j * b_kstep) >> 4),
                             idesc, (c > c_begin || tap > 0 || j > 0) ? 1u : 0u);
              }
              umma_commit_2sm(bempty_bar(sb_i));
              if (tap == taps - 1) {
                umma_commit_2sm(aempty_bar(sa_i));  // every tap of this channel block has read the halo tile
                if (c == c_end - 1) umma_commit_2sm(tfull_bar(as));
              }
            }
            __syncwarp();
            if (++sb_i == B_STAGES) {
              sb_i = 0;
              phb ^= 1u;
            }
          }
          if (++sa_i == A_BUFS) {
            sa_i = 0;
            pha ^= 1u;
          }
        }
        if (++as == 2) {
          as = 0;
          aph ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue (both CTAs drain their own 128 rows) =======================
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    int as = 0;
    uint32_t aph = 0;
    uint8_t* stg = sgen + STAGING_OFF + (warp - 4) * 4096;
    const uint32_t leader_tempty0 = mapa_rank(tempty_bar(0), 0);
    if (p.ragged) zero_fill_padded<BN>(p, cum, threadIdx.x - 128, blockIdx.x, gridDim.x);
    for (int it = 0, ptile = pair; it < n_iter; ++it, ptile += num_pairs) {
      const bool tail = has_tail && it == n_iter - 1;
      const TileCoord t = decode_pair(tail ? tail_tile : ptile);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (tail) {  // partial accumulator -> workspace, fixed-order reduction of this unit's column slice
        epilogue_split(p, t, tmem_base + as * BN, (int)rank, pair / ts.s, tail_j, ts.s, threadIdx.x - 128,
                       leader_tempty0 + 8u * as, reinterpret_cast<volatile int*>(sgen + TMEM_PTR_OFF + 8));
        break;
      }
      epilogue_tile<BN>(p, t, tmem_base + as * BN, stg, q, chalf, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tempty0 + 8u * as);
      if (++as == 2) {
        as = 0;
        aph ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

static int c2_num_sms = 0;

// header + one slab per CTA of every pair that can hold a slice
long long conv_tc2_workspace_bytes(int pairs) {
  return c2::WS_HEADER_BYTES + (long long)pairs * 2 * c2::WS_SLAB_FLOATS * 4;
}

// NORMAL mode, K-major A, taps in [2, 16], M > 128.  kp: output of gemm_fill_params.
template <int BN>
static int conv_tc2_launch_t(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  using namespace c2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv_tc2)", e);
    attr_set = true;
  }
  const int taps = g.taps;
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_bf16_3d(&tmA, g.a.ptr, g.a.inner, g.a.rows, g.a.batches, g.a.ld, g.a.batch_stride, 64,
                                 BM + taps - 1))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmB, g.b.ptr, g.b.inner, g.b.rows, g.b.batches, g.b.ld, g.b.batch_stride, 64,
                                 g.b.mn_major ? 64 : BN / 2))
    return rc;
  kp.n_tiles_per_tap = (g.N + BN - 1) / BN;
  kp.tiles_n = kp.n_tiles_per_tap;
  if (kp.pair_any) {
    // units stay single 128-row tiles; pairs are formed across utterances
  } else if (kp.row_lens) {  // schedule units are 256-row pair tiles
    kp.unit_rows = 2 * BM;
    kp.units_max = (kp.tiles_m + 1) / 2;
  }
  const int pair_tiles = ((kp.tiles_m + 1) / 2) * kp.tiles_n * kp.Z;
  kp.total_tiles = kp.tiles_m * kp.tiles_n * kp.Z;
  if (pair_tiles <= 0) return 0;
  if (!c2_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&c2_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int max_pairs = c2_num_sms / 2;
  const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  // split-K tail of the last partial round: bf16 outputs with a large enough, 16-byte aligned workspace
  static const bool no_split = getenv("FS2_CONV_NO_SPLIT") != nullptr;  // A/B switch (tools/)
  kp.ws = nullptr;
  if (!no_split && BN == 256 && g.workspace && !g.d_f32 && !g.d_atomic && !(reinterpret_cast<uintptr_t>(g.workspace) & 15) &&
      g.workspace_bytes >= conv_tc2_workspace_bytes(pairs) && (!g.relu_mask || (g.N & 31) == 0))
    kp.ws = static_cast<float*>(g.workspace);
  FS2_LAUNCH((conv_tc2_kernel<BN>), 2 * pairs, kThreadsC2, DYN_BYTES, stream, tmA, tmB, kp, taps);
  count_launch();
  return check_launch("conv_tc2_kernel");
}

int conv_tc2_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  // 128-column tiles for outputs that are one 256-column tile wide: better wave quantisation on paper, but measured
  // SLOWER on B200 (k = 9 input gradient, ragged C2: 152 -> 172 us; the activation halo tile is loaded once per
  // column tile, which doubles the L2 -> SM traffic of an already fabric-bound kernel).  Opt-in: FS2_CONV_BN128=1.
  static const bool bn128 = getenv("FS2_CONV_BN128") && atoi(getenv("FS2_CONV_BN128")) == 1;
  if (bn128 && g.N > 128 && g.N <= 256 && (long long)kp.tiles_m * kp.Z >= 148) return conv_tc2_launch_t<128>(g, kp, stream);
  if (g.N <= 128) return conv_tc2_launch_t<128>(g, kp, stream);  // narrow outputs: one 128-column tile
  return conv_tc2_launch_t<256>(g, kp, stream);
}

}  // namespace fs2
