// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM / implicit-GEMM Conv1d / weight-gradient
// engine: a cluster of two CTAs on neighbouring SMs computes one 256 x 256 output tile.
//
// Each CTA owns 128 of the 256 rows (its own fp32 accumulator half in its own TMEM) and loads, per
// 64-wide K block, its own A tile [128 x 64] and HALF of the B tile [128 x 64]; the leader CTA's single
// MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), for which the tensor cores of
// both SMs read both halves of B.  Per SM that is 8 KiB of operand reads + 32 KiB of TMA fills per
// 64-wide K block instead of 12 + 48 KiB for the 1-CTA 128 x 256 tile -- shared-memory bandwidth is what
// caps the 1-CTA kernel at ~50 % tensor-pipe utilisation (profiles/README.md).
//
// Synchronisation (all mbarriers live at the same shared-memory offsets in both CTAs):
//   full[s]   : leader's barrier; both CTAs' TMA loads complete_tx on it (cp.async.bulk.tensor
//               .cta_group::2 with the leader's barrier address), the leader arms it with the bytes of both
//   empty[s]  : per CTA; released by tcgen05.commit ... multicast::cluster (mask 0b11)
//   tfull[a]  : per CTA; accumulator stage ready (multicast commit)
//   tempty[a] : leader's barrier, 16 arrivals: the 8 epilogue warps of both CTAs (remote arrive via mapa)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"
#include "ptx2sm.cuh"
#include "tmap.h"

namespace fs2 {

namespace g2 {
constexpr int BN = 256;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;          // 16 KiB: this CTA's 128 rows
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KiB: this CTA's half of the 256 B rows
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_OFF = STAGES * STAGE_BYTES;
constexpr int STAGING_BYTES = 8 * 32 * 128;
constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int CUM_OFF = TMEM_PTR_OFF + 16;  // ragged-schedule prefix table
constexpr int DYN_BYTES = CUM_OFF + kMaxRaggedZ * 4 + 1024;
constexpr int kThreads2 = 384;
}  // namespace g2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::kThreads2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmKP p) {
  pdl_trigger();
  using namespace g2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
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
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  cluster_sync_all();  // barrier inits visible cluster-wide before any remote arrive / TMA signal
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

  // A "pair tile" covers the 128-row tiles tm = 2*pm and 2*pm+1; this CTA works on tm = 2*pm + rank.
  const int pair_tiles_m = (p.tiles_m + 1) >> 1;
  const int total_pair_tiles = pair_sched_total(p, cum, pair_tiles_m * p.tiles_n * p.Z);
  auto decode_pair = [&](int ptile) {
    if (p.mode == FS2_GEMM_NORMAL) return decode_pair_normal(p, cum, ptile, (int)rank, pair_tiles_m);
    TileCoord t;
    t.tn = ptile % p.tiles_n;
    const int r = ptile / p.tiles_n;
    t.tm = 2 * (r % pair_tiles_m) + (int)rank;
    t.z = r / pair_tiles_m;  // split index
    wgrad_range(p, cum, t.z, t.kb0, t.nkb);
    return t;
  };

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      const uint32_t leader_full0 = mapa_rank(full_bar(0), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int ptile = pair; ptile < total_pair_tiles; ptile += num_pairs) {
        const TileCoord t = decode_pair(ptile);
        const int m0 = t.tm * BM;
        RbCursor rc{};
        if (p.mode != FS2_GEMM_NORMAL && t.nkb > 0) rc = rb_seek(p, cum, t.kb0);
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (p.dbg & 4) {  // ablation: no operand loads
            if (leader) mbar_arrive(full_bar(s));
            if (++s == STAGES) {
              s = 0;
              ph ^= 1u;
            }
            continue;
          }
          if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * STAGE_BYTES);
          const uint32_t lfull = leader_full0 + 8u * s;
          const uint32_t sa = sbase + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          if (p.mode == FS2_GEMM_NORMAL) {
            const int tap = kb / p.kb_per_tap;
            const int k0 = (kb - tap * p.kb_per_tap) * BK;
            const int n0 = t.tn * BN + (int)rank * (BN / 2);  // this CTA's half of the B rows
            const int za = p.a_batched ? t.z / p.a_zdiv : 0;
            const int ia = p.a_inner_base + (t.z % p.a_zdiv) * p.a_zmod_stride;
            const int zb = p.b_batched ? t.z / p.b_zdiv : 0;
            const int ib = p.b_inner_base + (t.z % p.b_zdiv) * p.b_zmod_stride;
            if (!p.a_mn) {
              tma_load_3d_2sm(sa, &tmA, lfull, ia + k0, m0 + p.tap_shift0 + tap, za);
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_3d_2sm(sa + h * kChunkBytes, &tmA, lfull, ia + m0 + h * 64, k0, za);
            }
            if (!p.b_mn) {
              tma_load_3d_2sm(sb, &tmB, lfull, ib + tap * p.b_tap_kstride + k0, n0, zb);
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_3d_2sm(sb + h * kChunkBytes, &tmB, lfull, ib + tap * p.b_tap_kstride + n0 + h * 64, k0,
                                zb);
            }
          } else {
            const int zb = rc.zb;
            const int r0 = rc.lb * BK;
            rb_next(p, cum, rc);
            const int tap = t.tn / p.n_tiles_per_tap;
            const int c0 = (t.tn - tap * p.n_tiles_per_tap) * BN + (int)rank * (BN / 2);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              tma_load_3d_2sm(sa + h * kChunkBytes, &tmA, lfull, p.a_inner_base + m0 + h * 64, r0, zb);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              tma_load_3d_2sm(sb + h * kChunkBytes, &tmB, lfull, p.b_inner_base + c0 + h * 64,
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
    // ======================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) ========
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u, b_kstep = p.b_mn ? 16u * 128u : 32u;
      const uint64_t ad0 = make_smem_desc(sbase, a_lbo, 1024u);
      const uint64_t bd0 = make_smem_desc(sbase + A_BYTES, b_lbo, 1024u);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int ptile = pair; ptile < total_pair_tiles; ptile += num_pairs) {
        const TileCoord t = decode_pair(ptile);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t so = (uint64_t)((s * STAGE_BYTES) >> 4);
            if (!(p.dbg & 8)) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                umma_f16_2sm(tacc, ad0 + so + (uint64_t)((j * a_kstep) >> 4), bd0 + so + (uint64_t)((j * b_kstep) >> 4),
                             idesc, (kb > 0 || j > 0) ? 1u : 0u);
              }
            }
            umma_commit_2sm(empty_bar(s));
            if (kb == t.nkb - 1) umma_commit_2sm(tfull_bar(as));
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (t.nkb <= 0) {  // empty split (cannot happen with the host-side split computation; keep the protocol sound)
          if (elect_one()) umma_commit_2sm(tfull_bar(as));
          __syncwarp();
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
    if (p.ragged && p.mode == FS2_GEMM_NORMAL)
      zero_fill_padded<BN>(p, cum, threadIdx.x - 128, blockIdx.x, gridDim.x);
    for (int ptile = pair; ptile < total_pair_tiles; ptile += num_pairs) {
      const TileCoord t = decode_pair(ptile);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (!(p.dbg & 2)) epilogue_tile<BN>(p, t, tmem_base + as * BN, stg, q, chalf, lane);
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
  cluster_sync_all();  // the peer's smem / TMEM must stay alive until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

static int g2_num_sms = 0;

int gemm_tc2_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  using namespace g2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm_tc2)", e);
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  // A: this CTA's 128 rows (K-major) or 64x64 boxes (MN-major); B: HALF of the 256-wide tile per CTA
  if (int rc = make_tmap_bf16_3d(&tmA, g.a.ptr, g.a.inner, g.a.rows, g.a.batches, g.a.ld, g.a.batch_stride, 64,
                                 g.a.mn_major ? 64 : BM))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmB, g.b.ptr, g.b.inner, g.b.rows, g.b.batches, g.b.ld, g.b.batch_stride, 64,
                                 g.b.mn_major ? 64 : BN / 2))
    return rc;
  kp.n_tiles_per_tap = (g.N + BN - 1) / BN;
  if (kp.pair_any) {
    // units stay single 128-row tiles; pairs are formed across utterances
  } else if (kp.row_lens && g.mode == FS2_GEMM_NORMAL) {  // schedule units are 256-row pair tiles
    kp.unit_rows = 2 * BM;
    kp.units_max = (kp.tiles_m + 1) / 2;
  }
  const int taps = g.taps > 0 ? g.taps : 1;
  kp.tiles_n = (g.mode == FS2_GEMM_NORMAL) ? kp.n_tiles_per_tap : taps * kp.n_tiles_per_tap;
  const int pair_tiles = ((kp.tiles_m + 1) / 2) * kp.tiles_n * kp.Z;
  kp.total_tiles = kp.tiles_m * kp.tiles_n * kp.Z;
  if (pair_tiles <= 0) return 0;
  if (!g2_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g2_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int max_pairs = g2_num_sms / 2;
  const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  FS2_LAUNCH((gemm_tc2_kernel), 2 * pairs, kThreads2, DYN_BYTES, stream, tmA, tmB, kp);
  count_launch();
  return check_launch("gemm_tc2_kernel");
}

}  // namespace fs2
