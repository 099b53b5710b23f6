// 2-CTA Conv1d WEIGHT GRADIENT with operand reuse across taps: the k = 9 FFN convolution, the k = 5 PostNet
// convolutions and the k = 3 predictor convolutions (transformer/SubLayers.py:68-80, transformer/Layers.py:96-127,
// lightning/model/modules.py:204-232 -- their autograd backward w.r.t. the weight).
//
//   dW[co][tap][ci] += sum_{z, t} dY[z][t][co] * X[z][t + tap + shift0][ci]
//
// The plain weight-gradient tiling (gemm_tc2.cu) treats every tap as its own output tile and streams BOTH operands
// per 64-frame block: 32 KiB per CTA for 128 x 256 x 64 MACs = ~62 B/clk, above what the L2 -> SM fabric delivers
// (~43 B/clk/SM), and measures 0.65 of the tensor peak on the k = 9 gradient.  Here one output tile holds a GROUP of
// up to four taps of the same (256 co) x (128 ci) block:
//   * A = dY tile [64 frames][128 co per CTA] (MN-major), loaded once per frame block for all taps of the group;
//   * B = X tile [64 + taps_in_group - 1 frames][64 ci per CTA] (MN-major): ONE TMA box; tap `j` reads it through a
//     shared-memory descriptor whose start address is shifted by j rows (128 B) -- the same absolute-address
//     swizzle argument as the row-shifted A views of conv_tc2.cu, here along the reduction dimension;
//   * accumulators: taps_in_group x 128 fp32 columns of tensor memory (<= 512).
// Operand traffic per CTA and frame block: 16 KiB + ~8.3 KiB for taps x (128 x 128 x 64) MACs = ~31 B/clk at three
// taps: the kernel is tensor-bound again.
// Work units = (co pair tile, ci tile, tap group, split of the frame-block list); groups of different size get
// split counts in proportion to their taps (host), one unit per CTA pair in the common case.  Partial sums meet in
// the fp32 gradient with 16-byte vector reductions (the shared epilogue of gemm_common.cuh).
// Optional fused bias gradient (fs2_gemm::a_colsum): db[co] += sum_{z,t} dY[z][t][co].  The dY tiles are in shared
// memory anyway and the epilogue warps idle during the main loop, so the leader CTA's eight epilogue warps read
// both CTAs' tiles (the peer's through ld.shared::cluster) between the tile's full barrier and an extra arrival on
// its empty barrier; units of tap group 0 share the frame blocks of a (co tile, split) among the ci tiles.  This
// replaces a separate 73 MB column-sum launch per FFN layer (27 us alone, ~1.7 % of the C2 step).
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

namespace wt {
constexpr int BNT = 128;      // ci columns per tile (64 per CTA)
constexpr int MAX_TG = 4;     // taps per group
constexpr int MAX_GROUPS = 8;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB: [64 frames][128 co] as two 64 x 64 MN-major boxes
constexpr int B_BYTES = 9 * 1024;     // (64 + MAX_TG - 1) rows x 128 B = 8576 B, 1 KiB aligned
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_OFF = STAGES * STAGE_BYTES;
constexpr int STAGING_BYTES = 8 * 32 * 128;
constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 2;
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int CUM_OFF = TMEM_PTR_OFF + 16;
constexpr int DYN_BYTES = CUM_OFF + kMaxRaggedZ * 4 + 1024;
constexpr int kThreadsW = 384;

struct Sched {
  int n_groups;
  int tap0[MAX_GROUPS], ntap[MAX_GROUPS], splits[MAX_GROUPS];
  int unit0[MAX_GROUPS + 1];  // first unit of group g; unit0[n_groups] = total units
  int tiles_co, tiles_ci;     // pair tiles over co (256 rows), tiles over ci (128 columns)
  int box_rows;               // frames of the X box: 64 + (largest group) - 1
  int cs_all;                 // fused column sums: every tap group has the same split count (same frame-block ranges)
};
}  // namespace wt

struct WtUnit {
  int g, tap0, ntap, tm_pair, tci, kb0, nkb;
};

__device__ __forceinline__ WtUnit wt_decode(const GemmKP& p, const wt::Sched& sc, const int* cum, int unit) {
  WtUnit u;
  int g = 0;
  while (g + 1 < sc.n_groups && unit >= sc.unit0[g + 1]) ++g;
  u.g = g;
  u.tap0 = sc.tap0[g];
  u.ntap = sc.ntap[g];
  int r = unit - sc.unit0[g];
  const int tiles = sc.tiles_co * sc.tiles_ci;
  const int z = r / tiles;  // split of this group
  r -= z * tiles;
  u.tm_pair = r / sc.tiles_ci;
  u.tci = r - u.tm_pair * sc.tiles_ci;
  int total = p.total_rb;
  if (p.ragged) total = cum[p.sched_n - 1];
  const int Z = sc.splits[g];
  const int per = (total + Z - 1) / Z;
  u.kb0 = z * per;
  const int rem = total - u.kb0;
  u.nkb = rem < per ? rem : per;
  if (u.nkb < 0) u.nkb = 0;
  return u;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(wt::kThreadsW, 1)
wgrad_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ GemmKP p, const __grid_constant__ wt::Sched sc) {
  pdl_trigger();
  using namespace wt;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tempty_bar = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + TMEM_PTR_OFF);
  const int* cum = reinterpret_cast<const int*>(sgen + CUM_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.a_colsum ? 5 : 1);  // MMA commit (+ four reader warps of the leader CTA)
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_2sm(sbase + TMEM_PTR_OFF, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  pdl_wait();
  if (warp == 3 && p.ragged) build_ragged_table(p, reinterpret_cast<int*>(sgen + CUM_OFF), lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int total_units = sc.unit0[sc.n_groups];

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      const uint32_t leader_full0 = mapa_rank(full_bar(0), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int unit = pair; unit < total_units; unit += num_pairs) {
        const WtUnit u = wt_decode(p, sc, cum, unit);
        const int m0 = (2 * u.tm_pair + (int)rank) * BM;                // this CTA's 128 co
        const int c0 = u.tci * BNT + (int)rank * (BNT / 2);             // this CTA's 64 ci
        const uint32_t b_box_bytes = static_cast<uint32_t>(sc.box_rows) * 128u;  // the box of the LARGEST group
        RbCursor rc{};
        if (u.nkb > 0) rc = rb_seek(p, cum, u.kb0);
        for (int kb = 0; kb < u.nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * (A_BYTES + b_box_bytes));
          const uint32_t lfull = leader_full0 + 8u * s;
          const uint32_t sa = sbase + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          const int zb = rc.zb;
          const int r0 = rc.lb * BK;
          rb_next(p, cum, rc);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_3d_2sm(sa + h * kChunkBytes, &tmA, lfull, p.a_inner_base + m0 + h * 64, r0, zb);
          // frames [r0 + shift0 + tap0, + 64 + ntap - 1): rows outside the utterance are zero-filled by TMA
          tma_load_3d_2sm(sb, &tmB, lfull, p.b_inner_base + c0, r0 + p.tap_shift0 + u.tap0, zb);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) ========
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, BNT, 1, 1);
      const uint64_t ad0 = make_smem_desc(sbase, kChunkBytes, 1024u);
      const uint64_t bd0 = make_smem_desc(sbase + A_BYTES, kChunkBytes, 1024u);
      int s = 0;
      uint32_t ph = 0, aph = 0;
      for (int unit = pair; unit < total_units; unit += num_pairs) {
        const WtUnit u = wt_decode(p, sc, cum, unit);
        mbar_wait(tempty_bar, aph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < u.nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t so = (uint64_t)((s * STAGE_BYTES) >> 4);
            for (int tap = 0; tap < u.ntap; ++tap) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                // A: frames [16 j, 16 j + 16) of the dY tile; B: the same frames shifted by `tap` rows of 128 B
                umma_f16_2sm(tmem_base + tap * BNT, ad0 + so + (uint64_t)((j * 16 * 128) >> 4),
                             bd0 + so + (uint64_t)(((j * 16 + tap) * 128) >> 4), idesc, (kb > 0 || j > 0) ? 1u : 0u);
              }
            }
            umma_commit_2sm(empty_bar(s));
            if (kb == u.nkb - 1) umma_commit_2sm(tfull_bar);
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (u.nkb <= 0) {  // empty split of a ragged reduction: keep the protocol sound, the epilogue stores nothing
          if (elect_one()) umma_commit_2sm(tfull_bar);
          __syncwarp();
        }
        aph ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue (both CTAs drain their own 128 rows, one tap at a time) ===============
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    uint32_t aph = 0;
    uint8_t* stg = sgen + STAGING_OFF + (warp - 4) * 4096;
    const uint32_t leader_tempty = mapa_rank(tempty_bar, 0);
    // fused column sums of A: warps 4-7 read the leader's tile, 8-11 the peer's; within four warps: (64-co half,
    // 32-frame half); a lane owns two adjacent co
    const bool reader = p.a_colsum != nullptr && leader;
    const int rtile = (warp - 4) >> 2, rh = (warp - 4) & 1, rf = ((warp - 4) >> 1) & 1;
    const uint32_t rbase = mapa_rank(sbase, (uint32_t)rtile) + rh * kChunkBytes + rf * 32 * 128 + (lane & 3) * 4;
    const uint32_t rempty0 = mapa_rank(empty_bar(0), (uint32_t)rtile);
    int rs = 0;
    uint32_t rph = 0;
    for (int unit = pair; unit < total_units; unit += num_pairs) {
      const WtUnit u = wt_decode(p, sc, cum, unit);
      if (reader) {
        float c0 = 0.f, c1 = 0.f;
        // who sums which frame block of a (co tile, split): the units of ALL tap groups and ci tiles take turns when
        // every group has the same split ranges (sc.cs_all), else the ci tiles of group 0
        const bool cs_mine = sc.cs_all || u.g == 0;
        const int cs_mod = sc.cs_all ? sc.n_groups * sc.tiles_ci : sc.tiles_ci;
        const int cs_idx = sc.cs_all ? u.g * sc.tiles_ci + u.tci : u.tci;
        for (int kb = 0; kb < u.nkb; ++kb) {
          mbar_wait_cluster(full_bar(rs), rph);
          if (cs_mine && (u.kb0 + kb) % cs_mod == cs_idx) {
            colsum_tile_rows32(rbase + rs * STAGE_BYTES, lane, c0, c1);  // frames rf * 32 .. + 31 of this tile
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(rempty0 + 8u * rs);
          if (++rs == STAGES) {
            rs = 0;
            rph ^= 1u;
          }
        }
        if (cs_mine && u.nkb > 0) {
          const int co = (2 * u.tm_pair + rtile) * BM + rh * 64 + lane * 2;
          if (co < p.M) atomicAdd(p.a_colsum + co, c0);
          if (co + 1 < p.M) atomicAdd(p.a_colsum + co + 1, c1);
        }
      }
      mbar_wait(tfull_bar, aph);
      tc_fence_after();
      TileCoord t;
      t.z = 0;
      t.tm = 2 * u.tm_pair + (int)rank;
      t.nkb = u.nkb;
      t.kb0 = u.kb0;
      for (int tap = 0; tap < u.ntap; ++tap) {
        t.tn = (u.tap0 + tap) * p.n_tiles_per_tap + u.tci;  // epilogue_tile: tap = tn / n_tiles_per_tap
        epilogue_tile<BNT>(p, t, tmem_base + tap * BNT, stg, q, chalf, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tempty);
      aph ^= 1u;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

static int wt_num_sms = 0;

// WGRAD with taps >= 2, ci (N) a multiple of 128, co (M) >= 256, unit-stride fp32 atomic output without segments.
bool wgrad_taps_eligible(const fs2_gemm& g, const GemmKP& kp) {
  static const bool off = getenv("FS2_WGRAD_NO_TAPS") != nullptr;  // A/B switch (tools/)
  return !off && g.mode == FS2_GEMM_WGRAD && g.taps >= 2 && g.taps <= wt::MAX_TG * wt::MAX_GROUPS &&
         (g.N % wt::BNT) == 0 && g.M >= 256 && g.d_atomic && g.d_f32 && kp.d_col_stride == 1 && g.d_seg_rows == 0 &&
         g.a.mn_major && g.b.mn_major;
}

int wgrad_taps_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  using namespace wt;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(wgrad_taps)", e);
    attr_set = true;
  }
  if (!wt_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&wt_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int max_pairs = wt_num_sms / 2;
  Sched sc{};
  const int taps = g.taps;
  sc.n_groups = (taps + MAX_TG - 1) / MAX_TG;
  const int tg = (taps + sc.n_groups - 1) / sc.n_groups;  // 9 -> 3+3+3, 5 -> 3+2, 3 -> 3
  sc.n_groups = (taps + tg - 1) / tg;
  sc.tiles_co = (kp.tiles_m + 1) / 2;
  sc.tiles_ci = g.N / BNT;
  sc.box_rows = BK + tg - 1;
  const int tiles = sc.tiles_co * sc.tiles_ci;
  for (int i = 0, t0 = 0; i < sc.n_groups; ++i, t0 += tg) {
    sc.tap0[i] = t0;
    sc.ntap[i] = taps - t0 < tg ? taps - t0 : tg;
    sc.splits[i] = 1;
  }
  // split counts: one unit per CTA pair when the tiles allow it, the next split always goes to the group whose
  // units are the longest (taps / splits); never more splits than reduction blocks
  if (g.splits > 1 || kp.total_rb > 1) {
    int units = tiles * sc.n_groups;
    for (;;) {
      int best = -1;
      for (int i = 0; i < sc.n_groups; ++i) {
        if (sc.splits[i] >= kp.total_rb) continue;
        if (best < 0 || sc.ntap[i] * sc.splits[best] > sc.ntap[best] * sc.splits[i]) best = i;
      }
      if (best < 0 || units + tiles > max_pairs) break;
      ++sc.splits[best];
      units += tiles;
    }
  }
  sc.cs_all = 1;
  for (int i = 1; i < sc.n_groups; ++i)
    if (sc.splits[i] != sc.splits[0]) sc.cs_all = 0;
  sc.unit0[0] = 0;
  for (int i = 0; i < sc.n_groups; ++i) sc.unit0[i + 1] = sc.unit0[i] + tiles * sc.splits[i];
  const int total_units = sc.unit0[sc.n_groups];
  if (total_units <= 0) return 0;
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_bf16_3d(&tmA, g.a.ptr, g.a.inner, g.a.rows, g.a.batches, g.a.ld, g.a.batch_stride, 64, 64))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmB, g.b.ptr, g.b.inner, g.b.rows, g.b.batches, g.b.ld, g.b.batch_stride, 64,
                                 BK + tg - 1))
    return rc;
  kp.n_tiles_per_tap = g.N / BNT;
  kp.n_per_tap = g.N;  // (kp.a_colsum: set by the dispatcher when this kernel sums the dY tiles)
  const int pairs = total_units < max_pairs ? total_units : max_pairs;
  FS2_LAUNCH((wgrad_taps_kernel), 2 * pairs, kThreadsW, DYN_BYTES, stream, tmA, tmB, kp, sc);
  count_launch();
  return check_launch("wgrad_taps_kernel");
}

}  // namespace fs2
