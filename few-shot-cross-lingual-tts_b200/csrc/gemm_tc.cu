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
#include <cstdlib>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace fs2 {

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
  static constexpr int CUM_OFF = TMEM_PTR_OFF + 16;  // ragged-schedule prefix table
  static constexpr int TOTAL = CUM_OFF + kMaxRaggedZ * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmKP p) {
  pdl_trigger();
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
  const int* cum = reinterpret_cast<const int*>(sgen + L::CUM_OFF);

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
  pdl_wait();  // everything above is independent of the previous kernel's output
  if (warp == 3 && p.ragged) build_ragged_table(p, reinterpret_cast<int*>(sgen + L::CUM_OFF), lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int total_tiles = sched_total(p, cum, p.total_tiles);

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, cum, tile);
        const int m0 = t.tm * BM;
        RbCursor rc{};
        if (p.mode != FS2_GEMM_NORMAL && t.nkb > 0) rc = rb_seek(p, cum, t.kb0);
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
            const int zb = rc.zb;
            const int r0 = rc.lb * BK;
            rb_next(p, cum, rc);
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
    // ======================= MMA issuer (warp-uniform loop, one elected lane issues) =======================
    {
      const uint32_t idesc = make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u, b_kstep = p.b_mn ? 16u * 128u : 32u;
      const uint64_t ad0 = make_smem_desc(sbase, a_lbo, 1024u);
      const uint64_t bd0 = make_smem_desc(sbase + L::A_BYTES, b_lbo, 1024u);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, cum, tile);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t so = (uint64_t)((s * L::STAGE_BYTES) >> 4);
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              umma_f16(tacc, ad0 + so + (uint64_t)((j * a_kstep) >> 4), bd0 + so + (uint64_t)((j * b_kstep) >> 4), idesc,
                       (kb > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(s));  // frees the smem slot when these MMAs retire
            if (kb == t.nkb - 1) umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++s == STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (t.nkb <= 0) {
          if (elect_one()) umma_commit(tfull_bar(as));
          __syncwarp();
        }
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
    int as = 0;
    uint32_t aph = 0;
    uint8_t* stg = sgen + L::STAGING_OFF + (warp - 4) * 4096;
    if (p.ragged && p.mode == FS2_GEMM_NORMAL)
      zero_fill_padded<BN>(p, cum, threadIdx.x - 128, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, cum, tile);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      epilogue_tile<BN>(p, t, tmem_base + as * BN, stg, q, chalf, lane);
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
  FS2_LAUNCH((gemm_tc_kernel<BN, STAGES>), grid, kThreads, L::DYN_BYTES, stream, tmA, tmB, kp);
  count_launch();
  return check_launch("gemm_tc_kernel");
}

int gemm_fill_params(const fs2_gemm& g, GemmKP& kp) {
  kp = GemmKP{};
  static const int dbg = getenv("FS2_GEMM_DBG") ? atoi(getenv("FS2_GEMM_DBG")) : 0;
  kp.dbg = dbg;
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
  if (g.row_lens) {
    kp.row_lens = reinterpret_cast<const long long*>(g.row_lens);
    kp.lens_zdiv = g.lens_zdiv > 0 ? g.lens_zdiv : 1;
    kp.tail_zero = g.tail_zero_rows;
    if (g.mode == FS2_GEMM_NORMAL) {
      if (g.d_atomic) return set_error("gemm: row_lens and atomic outputs do not combine in NORMAL mode");
      kp.sched_n = kp.Z;
      kp.unit_rows = BM;  // the pair kernel doubles it
      kp.units_max = kp.tiles_m;
      kp.row_extent = g.M;
    } else {
      kp.sched_n = g.a.batches;
      kp.unit_rows = BK;
      kp.units_max = kp.rb_per_batch;
      kp.row_extent = g.a.rows;
    }
    kp.ragged = kp.sched_n <= kMaxRaggedZ ? 1 : 0;  // larger batches: dense schedule, rows still zeroed
    // 2-CTA kernels may team up row tiles of different utterances when both see the same B operand
    static const bool no_pair_any = getenv("FS2_NO_PAIR_ANY") != nullptr;  // A/B switch (tools/)
    kp.pair_any = (!no_pair_any && kp.ragged && g.mode == FS2_GEMM_NORMAL && g.b.batches <= 1 &&
                   g.b.zmod_stride == 0) ? 1 : 0;
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
  if (g.epilogue == FS2_EPI_ADD_AUX && !g.aux) return set_error("epilogue needs aux");
  if (g.epilogue == FS2_EPI_RELU_BWD && !g.aux && !g.relu_mask) return set_error("ReLU backward needs aux or relu_mask");
  if (g.relu_mask) {
    if (g.mode != FS2_GEMM_NORMAL || g.d_f32 || (g.N & 63) || (reinterpret_cast<uintptr_t>(g.relu_mask) & 7))
      return set_error("gemm: relu_mask needs NORMAL mode, a bf16 output and N % 64 == 0");
    if (g.epilogue == FS2_EPI_RELU_BWD && g.aux) return set_error("gemm: pass either aux or relu_mask to the ReLU backward");
    if (g.epilogue != FS2_EPI_RELU && g.epilogue != FS2_EPI_RELU_BWD) return set_error("gemm: relu_mask without a ReLU epilogue");
    kp.relu_mask = static_cast<unsigned long long*>(g.relu_mask);
  }
  if (kp.d_col_stride != 1) return set_error("gemm: outputs must have unit column stride (d_col_stride = 1)");
  return 0;
}

int gemm_tc2_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream);  // gemm_tc2.cu (2-CTA tiles)
int conv_tc2_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream);  // conv_tc2.cu (2-CTA + halo reuse)
bool gemm_sk_eligible(const fs2_gemm& g, const GemmKP& kp);                // gemm_sk.cu (small K, resident weights,
int gemm_sk_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream);    //  TMA-store epilogue)
bool wgrad_taps_eligible(const fs2_gemm& g, const GemmKP& kp);             // wgrad_taps.cu (Conv1d weight gradient,
int wgrad_taps_launch(const fs2_gemm& g, GemmKP& kp, cudaStream_t stream); //  operand reuse across taps)

// fs2_gemm::a_colsum without the fused kernel: one column-sum launch over the (valid rows of the) A operand
static int a_colsum_fallback(const fs2_gemm& g, cudaStream_t stream) {
  if (!g.a.mn_major || g.a.batch_stride != g.a.rows * g.a.ld || (g.row_lens && g.lens_zdiv != 1) || (g.M % 8))
    return set_error("gemm: a_colsum needs a densely batched MN-major A (batch_stride == rows * ld), M % 8 == 0");
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(g.a.ptr) + g.a.inner_base;
  const int nseg = g.d_seg_rows > 0 ? (g.M + g.d_seg_rows - 1) / g.d_seg_rows : 1;
  for (int i = 0; i < nseg; ++i) {
    float* out = g.d_seg_rows > 0 ? g.a_colsum_seg[i] : g.a_colsum;
    if (!out) continue;
    const int c0 = i * g.d_seg_rows, cols = g.d_seg_rows > 0 ? (g.M - c0 < g.d_seg_rows ? g.M - c0 : g.d_seg_rows) : g.M;
    if ((c0 % 8) || (cols % 8)) return set_error("gemm: a_colsum segments must be multiples of 8 columns");
    const int rc = g.row_lens ? fs2_colsum_ragged_bf16(x + c0, g.a.ld, (int)g.a.batches, (int)g.a.rows, cols, g.row_lens,
                                                       out, stream)
                              : fs2_colsum_bf16(x + c0, g.a.ld, 1, (int)(g.a.batches * g.a.rows), cols, out, stream);
    if (rc) return rc;
  }
  return 0;
}

int gemm_tc_launch(const fs2_gemm& g, cudaStream_t stream) {
  GemmKP kp;
  if (int rc = gemm_fill_params(g, kp)) return rc;
  kp.a_colsum = nullptr;
  kp.a_colsum_on = 0;
  for (int i = 0; i < 4; ++i) kp.a_colsum_seg[i] = nullptr;
  // Tile-N choice: 256 columns (128x256 is the shape that can reach the tensor-pipe peak from one
  // CTA) unless the 128-wide tiling wastes clearly less of a ragged N.
  const int n = g.N;
  const int pad256 = ((n + 255) / 256) * 256, pad128 = ((n + 127) / 128) * 128;
  const bool use128 = (n <= 128) || (pad128 * 100 < pad256 * 92);
  const bool want_seg = g.a_colsum_seg[0] || g.a_colsum_seg[1] || g.a_colsum_seg[2] || g.a_colsum_seg[3];
  if (g.a_colsum || want_seg) {  // fused where one of the two kernels that read their dY tiles runs (order below)
    if (g.mode != FS2_GEMM_WGRAD) return set_error("gemm: a_colsum is a weight-gradient (WGRAD) option");
    if ((g.d_seg_rows > 0) != want_seg || (want_seg && g.a_colsum))
      return set_error("gemm: a_colsum_seg goes with a segmented output (d_seg_rows > 0), a_colsum with a plain one");
    static const bool no_fuse = getenv("FS2_NO_FUSED_COLSUM") != nullptr;  // A/B switch
    static const bool no_2cta_ = getenv("FS2_GEMM_NO_2CTA") != nullptr;
    const bool routed = !no_fuse && !g.ln_gamma && !use128 && !gemm_sk_eligible(g, kp);
    const bool fused_taps = routed && wgrad_taps_eligible(g, kp);
    const bool fused_tc2 = routed && !fused_taps && !no_2cta_ && (g.M >= 256 || kp.pair_any != 0);
    if (fused_taps || fused_tc2) {
      kp.a_colsum = g.a_colsum;
      for (int i = 0; i < 4; ++i) kp.a_colsum_seg[i] = g.a_colsum_seg[i];
      kp.a_colsum_on = 1;
    } else if (int rc = a_colsum_fallback(g, stream)) {
      return rc;
    }
  }
  // Conv1d (taps > 1, K-major activations): one activation tile with halo serves every tap, 2-CTA tiles -- also for
  // narrow outputs (the 512 -> 80 PostNet convolution and the input gradient of the 80 -> 512 one: 128-column
  // pair tiles; the 1-CTA kernel re-read the activation tile once per tap: 94 / 72 us for 26 GFLOP)
  static const bool no_halo = getenv("FS2_CONV_NO_HALO") != nullptr;
  // (with a ragged schedule and shared weights the pair kernels team up ANY two row tiles, so even
  //  utterances of a single 128-row tile fill both CTAs)
  const bool pair_any = kp.pair_any != 0;
  if (!no_halo && g.mode == FS2_GEMM_NORMAL && g.taps > 1 && g.taps <= 16 && !g.a.mn_major &&
      (g.M > 128 || pair_any) && (!use128 || n <= 128))
    return conv_tc2_launch(g, kp, stream);
  if (g.ln_gamma) return gemm_tc2_launch(g, kp, stream);  // fused LayerNorm epilogue: validated there
  if (use128) return launch_tc<128, 6>(g, kp, stream);
  if (gemm_sk_eligible(g, kp)) return gemm_sk_launch(g, kp, stream);
  if (wgrad_taps_eligible(g, kp)) return wgrad_taps_launch(g, kp, stream);
  // 256-row x 256-column tiles on CTA pairs (cta_group::2) when every pair has two real row tiles:
  // halves the shared-memory traffic per FLOP, which is what limits the 1-CTA 128x256 tile.
  static const bool no_2cta = getenv("FS2_GEMM_NO_2CTA") != nullptr;
  // measured (tools/bench_gemm.py, B200): +14 % on the k=9 weight gradient, +7 % on K=1024 GEMMs, +2 % on
  // 8192^3, but -5 % on the implicit-GEMM conv forward -> taps > 1 stay on the 1-CTA kernel.
  const int taps_ = g.taps > 0 ? g.taps : 1;
  if (!no_2cta && (g.M >= 256 || pair_any) && (g.mode == FS2_GEMM_WGRAD || taps_ == 1))
    return gemm_tc2_launch(g, kp, stream);
  return launch_tc<256, 4>(g, kp, stream);
}

}  // namespace fs2
