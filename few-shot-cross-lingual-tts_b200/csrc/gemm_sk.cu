// B-stationary 2-CTA GEMM for the small-K projections of the FFT block (K <= 256): the fused Q|K|V
// projection, the attention output projection, the k=1 Conv1d input-gradient and the output-projection
// input-gradient (transformer/SubLayers.py:39-41,54-55,87-91 and their backward).
//
// These GEMMs move ~4 bytes per FLOP-unit more than the k=9 convolutions and are bound by memory and by
// per-tile overheads, not by the tensor pipe.  What the generic pair kernel (gemm_tc2.cu) spends its time
// on for them (tools/time_qkv.py ablations): every pair tile re-loads its 128 KiB weight
// tile through L2 while all 148 CTAs hammer the same few hundred KiB of weights, and the epilogue moves
// every output element through registers twice (staging -> LDS -> STG with per-lane address math).  Here:
//
//   * the weight tile of the current 256-column block stays RESIDENT in shared memory (64 KiB per CTA:
//     K/64 k-blocks of this CTA's 128 weight rows); pair tiles are enumerated column-block-major and
//     every CTA pair owns one contiguous range, so the weights are (re)loaded at most tiles_n times per
//     CTA instead of once per tile, and the only per-tile operand traffic is the pair's 2 x 64 KiB of
//     activations;
//   * the epilogue writes bf16 rows into a 128-byte-swizzled staging tile and ONE thread per warp hands
//     it to TMA (cp.async.bulk.tensor shared -> global, SASS UTMASTG): no LDS / STG / address math, the
//     tensor map clips rows that belong to the next utterance; staging is double buffered per warp
//     (cp.async.bulk.wait_group.read 1), the bias vector is staged in shared memory once per CTA.
//
// Scheduling, barriers and the ragged (padded frames skipped) tile list are those of gemm_tc2.cu; extra
// barriers: bfull (weights of the current column block have landed, leader's barrier, both CTAs'
// TMA loads complete on it) and bempty (all MMAs that read the previous weights have retired).
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

namespace sk {
constexpr int BN = 256;
constexpr int STAGES = 5;
constexpr int MAX_KB = 4;                       // K <= 256
constexpr int A_BYTES = BM * BK * 2;            // 16 KiB: this CTA's 128 rows of one k-block
constexpr int BKB_BYTES = (BN / 2) * BK * 2;    // 16 KiB: this CTA's half of the weight rows, one k-block
constexpr int B_OFF = 0;                        // resident weights: MAX_KB k-blocks
constexpr int A_OFF = B_OFF + MAX_KB * BKB_BYTES;
constexpr int STAGING_OFF = A_OFF + STAGES * A_BYTES;
constexpr int STAGING_BYTES = 8 * 2 * 4096;     // 8 epilogue warps x 2 buffers x (32 rows x 128 B)
constexpr int BIAS_OFF = STAGING_OFF + STAGING_BYTES;
constexpr int MAX_N = 1024;
constexpr int BAR_OFF = BIAS_OFF + MAX_N * 4;
constexpr int NUM_BARS = 2 * STAGES + 4 + 2;    // full / empty ring, tfull / tempty x 2, bfull, bempty
constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
constexpr int CUM_OFF = TMEM_PTR_OFF + 16;
constexpr int DYN_BYTES = CUM_OFF + kMaxRaggedZ * 4 + 1024;
constexpr int kThreadsSk = 384;
static_assert(DYN_BYTES <= 232448, "shared memory budget");
}  // namespace sk

__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct SkItem {
  int tn, z, tm, valid;
};

// item -> (column block, this CTA's 128-row tile).  Items are column-block-major: item = tn * pair_rows + r.
__device__ __forceinline__ SkItem sk_decode(const GemmKP& p, const int* cum, int item, int pair_rows, int total_rows,
                                            int rank) {
  SkItem t;
  t.tn = item / pair_rows;
  const int r = item - t.tn * pair_rows;
  int idx = 2 * r + rank;
  t.valid = 1;
  if (idx >= total_rows) {  // odd tail: the filler half multiplies the partner's rows again and stores nothing
    idx = 2 * r;
    t.valid = 0;
  }
  if (p.ragged) {
    t.z = ragged_find(cum, p.sched_n, idx);
    t.tm = idx - (t.z ? cum[t.z - 1] : 0);
  } else {
    t.z = idx / p.tiles_m;
    t.tm = idx - t.z * p.tiles_m;
  }
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(sk::kThreadsSk, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ GemmKP p) {
  pdl_trigger();
  using namespace sk;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t bempty_bar = bar_base + 8u * (2 * STAGES + 5);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + TMEM_PTR_OFF);
  const int* cum = reinterpret_cast<const int*>(sgen + CUM_OFF);
  float* s_bias = reinterpret_cast<float*>(sgen + BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
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
    mbar_init(bfull_bar, 1);
    mbar_init(bempty_bar, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_2sm(sbase + TMEM_PTR_OFF, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  pdl_wait();  // everything above is independent of the previous kernel's output
  if (warp == 3) {
    if (p.ragged) build_ragged_table(p, reinterpret_cast<int*>(sgen + CUM_OFF), lane);
    for (int i = lane; i < p.N; i += 32) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_rows = p.ragged ? cum[p.sched_n - 1] : p.tiles_m * p.Z;  // 128-row tiles with work
  const int pair_rows = (total_rows + 1) >> 1;
  const int total_items = pair_rows * p.tiles_n;
  const int chunk = (total_items + num_pairs - 1) / num_pairs;
  const int item0 = pair * chunk;
  const int item1 = min(item0 + chunk, total_items);
  const int num_kb = p.num_kb;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0 && item0 < item1) {
      const uint32_t leader_full0 = mapa_rank(full_bar(0), 0);
      const uint32_t leader_bfull = mapa_rank(bfull_bar, 0);
      int s = 0, cur_tn = -1;
      uint32_t ph = 0, bph = 0;
      for (int item = item0; item < item1; ++item) {
        const SkItem t = sk_decode(p, cum, item, pair_rows, total_rows, (int)rank);
        if (t.tn != cur_tn) {  // (re)load the resident weights of this column block
          if (cur_tn >= 0) {
            mbar_wait(bempty_bar, bph);  // every MMA that read the previous weights has retired
            bph ^= 1u;
          }
          cur_tn = t.tn;
          if (leader) mbar_arrive_expect_tx(bfull_bar, 2 * num_kb * BKB_BYTES);
          const int n0 = t.tn * BN + (int)rank * (BN / 2);
          for (int kb = 0; kb < num_kb; ++kb) {
            const uint32_t sb = sbase + B_OFF + kb * BKB_BYTES;
            if (!p.b_mn) {
              tma_load_3d_2sm(sb, &tmB, leader_bfull, p.b_inner_base + kb * BK, n0, 0);
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_3d_2sm(sb + h * kChunkBytes, &tmB, leader_bfull, p.b_inner_base + n0 + h * 64, kb * BK, 0);
            }
          }
        }
        const int m0 = t.tm * BM;
        const int za = p.a_batched ? t.z : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (p.dbg & 4) {  // ablation: no activation loads
            if (leader) mbar_arrive(full_bar(s));
          } else {
            if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * A_BYTES);
            tma_load_3d_2sm(sbase + A_OFF + s * A_BYTES, &tmA, leader_full0 + 8u * s, p.a_inner_base + kb * BK, m0,
                            za);
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
    if (leader && item0 < item1) {
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, p.b_mn);
      const uint32_t b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t b_kstep = p.b_mn ? 16u * 128u : 32u;
      const uint64_t ad0 = make_smem_desc(sbase + A_OFF, 16u, 1024u);
      const uint64_t bd0 = make_smem_desc(sbase + B_OFF, b_lbo, 1024u);
      int s = 0, as = 0, cur_tn = -1;
      uint32_t ph = 0, aph = 0, bph = 0;
      for (int item = item0; item < item1; ++item) {
        const int tn = item / pair_rows;
        if (tn != cur_tn) {
          mbar_wait(bfull_bar, bph);
          bph ^= 1u;
          cur_tn = tn;
        }
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        const bool last_of_tn = (item + 1 == item1) || ((item + 1) / pair_rows != tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad_s = ad0 + (uint64_t)((s * A_BYTES) >> 4);
            const uint64_t bd_k = bd0 + (uint64_t)((kb * BKB_BYTES) >> 4);
            if (!(p.dbg & 8)) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                umma_f16_2sm(tacc, ad_s + (uint64_t)((j * 32) >> 4), bd_k + (uint64_t)((j * b_kstep) >> 4), idesc,
                             (kb > 0 || j > 0) ? 1u : 0u);
              }
            }
            umma_commit_2sm(empty_bar(s));
            if (kb == num_kb - 1) {
              umma_commit_2sm(tfull_bar(as));
              if (last_of_tn && item + 1 < item1) umma_commit_2sm(bempty_bar);  // the weights may be replaced
            }
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
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
    const int ew = warp - 4;
    const int q = warp & 3, chalf = ew >> 2;
    int as = 0, buf = 0;
    uint32_t aph = 0;
    uint8_t* stg0 = sgen + STAGING_OFF + ew * 8192;
    const uint32_t stg0_s = sbase + STAGING_OFF + ew * 8192;
    const uint32_t leader_tempty0 = mapa_rank(tempty_bar(0), 0);
    if (p.ragged) zero_fill_padded<BN>(p, cum, threadIdx.x - 128, blockIdx.x, gridDim.x);
    const int sw = lane & 7;
    for (int item = item0; item < item1; ++item) {
      const SkItem t = sk_decode(p, cum, item, pair_rows, total_rows, (int)rank);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const int m_w0 = t.tm * BM + q * 32;
      const bool row_ok = !p.row_lens || (m_w0 + lane) < p.row_lens[t.z / p.lens_zdiv];
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
      unsigned long long* mask_row = nullptr;
      if (p.relu_mask) {
        const int gm = m_w0 + lane;
        if (gm < p.M && t.valid) mask_row = p.relu_mask + ((long long)t.z * p.M + gm) * (p.N >> 6);
      }
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 64) {
        const int n0 = t.tn * BN + c0;
        if (n0 >= p.N || (p.dbg & 2)) break;  // warp-uniform
        float f[64];
        {
          uint32_t v[32], v2[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = __uint_as_float(v[j]);
            f[32 + j] = __uint_as_float(v2[j]);
          }
        }
        if (p.alpha != 1.f) {
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] *= p.alpha;
        }
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + n0);  // warp-uniform address: broadcast
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bv = b4[j];
            f[4 * j] += bv.x;
            f[4 * j + 1] += bv.y;
            f[4 * j + 2] += bv.z;
            f[4 * j + 3] += bv.w;
          }
        }
        if (p.epilogue == FS2_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.f);
          if (mask_row) {
            unsigned int lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              lo |= (f[j] > 0.f ? 1u : 0u) << j;
              hi |= (f[32 + j] > 0.f ? 1u : 0u) << j;
            }
            mask_row[n0 >> 6] = (static_cast<unsigned long long>(hi) << 32) | lo;
          }
        } else if (p.epilogue == FS2_EPI_RELU_BWD) {
          const unsigned long long m = mask_row ? mask_row[n0 >> 6] : 0ull;
          const unsigned int lo = static_cast<unsigned int>(m), hi = static_cast<unsigned int>(m >> 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = (lo >> j) & 1u ? f[j] : 0.f;
            f[32 + j] = (hi >> j) & 1u ? f[32 + j] : 0.f;
          }
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] = 0.f;
        }
        // this warp's staging buffer `buf` may still be read by the TMA store issued two chunks ago
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint8_t* my_row = stg0 + buf * 4096 + lane * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[ch * 8 + 2 * h], f[ch * 8 + 2 * h + 1]);
            w[h] = *reinterpret_cast<uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(my_row + ((ch ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (t.valid && !(p.dbg & 1)) tma_store_3d(&tmD, stg0_s + buf * 4096, n0, m_w0, t.z);
          tma_store_commit();
        }
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tempty0 + 8u * as);
      if (++as == 2) {
        as = 0;
        aph ^= 1u;
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM must stay alive until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

static int gsk_num_sms = 0;

// Shapes this kernel takes (everything else stays on gemm_tc2 / gemm_tc): NORMAL mode, no taps, K-major
// activations, K <= 256 (resident weights), shared un-batched weights, bf16 output with one utterance per z
// (plain 3-D tensor map), N >= 256.
bool gemm_sk_eligible(const fs2_gemm& g, const GemmKP& kp) {
  static const bool off = getenv("FS2_NO_GEMM_SK") != nullptr;
  if (off) return false;
  const int taps = g.taps > 0 ? g.taps : 1;
  if (g.mode != FS2_GEMM_NORMAL || taps != 1 || g.a.mn_major || g.d_f32 || g.d_atomic) return false;
  if (kp.num_kb > sk::MAX_KB || g.N < 256 || g.N > sk::MAX_N || (g.N & 7) || (g.K & 7)) return false;
  if (g.b.batches > 1 || g.b.zmod_stride != 0 || g.a.zdiv > 1 || g.a.zmod_stride != 0) return false;
  if (kp.d_zdiv != 1 || g.d_zmod_stride != 0 || g.aux) return false;
  if (g.epilogue != FS2_EPI_NONE && g.epilogue != FS2_EPI_RELU && g.epilogue != FS2_EPI_RELU_BWD) return false;
  if (g.epilogue == FS2_EPI_RELU_BWD && !g.relu_mask) return false;
  if (g.row_lens && !kp.ragged) return false;  // > 256 utterances: dense fallback of the generic kernels
  if (kp.Z > 1 && (g.d_zdiv_stride & 7)) return false;
  if (kp.Z * kp.tiles_m < 2) return false;
  return true;
}

int gemm_sk_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream) {
  using namespace sk;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_sk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm_sk)", e);
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmD;
  if (int rc = make_tmap_bf16_3d(&tmA, g.a.ptr, g.a.inner, g.a.rows, g.a.batches, g.a.ld, g.a.batch_stride, 64, BM))
    return rc;
  if (int rc = make_tmap_bf16_3d(&tmB, g.b.ptr, g.b.inner, g.b.rows, g.b.batches, g.b.ld, g.b.batch_stride, 64,
                                 g.b.mn_major ? 64 : BN / 2))
    return rc;
  // output [Z][M][N]: 64-column x 32-row boxes (one epilogue warp's chunk); rows >= M are clipped by the TMA
  if (int rc = make_tmap_bf16_3d(&tmD, g.d, g.N, g.M, kp.Z, g.ldd, kp.Z > 1 ? g.d_zdiv_stride : (long long)g.M * g.ldd,
                                 64, 32))
    return rc;
  kp.n_tiles_per_tap = (g.N + BN - 1) / BN;
  kp.tiles_n = kp.n_tiles_per_tap;
  kp.total_tiles = kp.tiles_m * kp.tiles_n * kp.Z;
  if (kp.total_tiles <= 0) return 0;
  if (!gsk_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&gsk_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // upper bound of the work list (the ragged list is only known on the device): ceil(rows / 2) * tiles_n
  const int max_items = ((kp.tiles_m * kp.Z + 1) / 2) * kp.tiles_n;
  const int max_pairs = gsk_num_sms / 2;
  const int pairs = max_items < max_pairs ? max_items : max_pairs;
  FS2_LAUNCH((gemm_sk_kernel), 2 * pairs, kThreadsSk, DYN_BYTES, stream, tmA, tmB, tmD, kp);
  count_launch();
  return check_launch("gemm_sk_kernel");
}

}  // namespace fs2
