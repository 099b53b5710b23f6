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
#include "util.cuh"

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


// ------------------------------------------------------------------------------------------------------------
// Fused LayerNorm epilogue (fs2_gemm::ln_*): transformer/SubLayers.py:88-91 `layer_norm(dropout(w_2(h)) + residual)`
// and the padding-row zeroing of transformer/Layers.py:28 inside the GEMM that computes w_2 (N = 256: a CTA's 128
// accumulator rows are complete LayerNorm rows).  The K = 1024 main loop of that GEMM is bound by operand ingest
// (~1100 clk per k-block), so the 8 epilogue warps have ~18k cycles per tile: enough for the dropout mask (the same
// Philox stream, counters and 13-bit compare as layernorm.cu, so both paths draw identical masks), the residual add,
// the row statistics and the normalisation.  The fp32 pre-norm sum v goes back into the accumulator columns
// (tcgen05.st) between the passes instead of living in registers.  Thread = one row x 128 columns (warps w / w + 4
// split the columns); the two halves exchange their partial sums through shared memory (parity double-buffered).
// Outputs: D = y (bf16), ln_v (bf16 v for the backward), ln_mean / ln_rstd, ln_keep (1 bit per element).
// ------------------------------------------------------------------------------------------------------------
namespace g2 {
constexpr int LN_STAGES = 5;                       // the sixth stage's shared memory holds gamma / beta / exchange
constexpr int LN_GB_OFF = LN_STAGES * STAGE_BYTES;  // gamma[256], beta[256] f32
constexpr int LN_RED_OFF = LN_GB_OFF + 2 * 256 * 4;  // [parity 2][sum | sum of squares][half 2][128] f32
}  // namespace g2

__device__ __forceinline__ void bar_sync_epi2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  uint32_t a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = v[i];
    b[i] = v[16 + i];
  }
  tmem_st16(taddr, a);
  tmem_st16(taddr + 16, b);
}
// 64 bf16 columns (32 packed words) of this thread's row -> the warp's staging tile -> global rows (coalesced)
__device__ __forceinline__ void ln_flush64(uint8_t* stg, int lane, const uint32_t (&w)[32], __nv_bfloat16* base,
                                           long long ld, int m_w0, int M) {
  const int sw = lane & 7;
  uint8_t* my_row = stg + lane * 128;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<uint4*>(my_row + ((ch ^ sw) << 4)) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3), ch = lane & 7;
    if (m_w0 + r < M)
      *reinterpret_cast<uint4*>(base + (long long)r * ld + ch * 8) =
          *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
  }
  __syncwarp();
}

__device__ __forceinline__ void epilogue_ln_tile(const GemmKP& p, const TileCoord& t, uint32_t tmem_acc, uint8_t* stg,
                                                 int q, int chalf, int lane, const float* s_gb, float* s_red,
                                                 int parity) {
  constexpr int C = 256;
  if (!t.valid) return;  // filler half of an odd pair (uniform over the CTA): nothing to store, no barriers
  const int row = q * 32 + lane;
  const int m_w0 = t.tm * BM + q * 32;
  const int gm = m_w0 + lane;
  const bool in_rows = gm < p.M;
  bool row_ok = in_rows;
  if (p.row_lens) row_ok = in_rows && gm < p.row_lens[t.z / p.lens_zdiv];
  const long long R = (long long)t.z * p.M + (in_rows ? gm : 0);  // dense row index
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + chalf * 128;
  const int colb = chalf * 128;
  const bool drop = p.ln_p_drop > 0.f;
  const float scale = drop ? 1.f / (1.f - p.ln_p_drop) : 1.f;
  const uint32_t thresh = keep_thresh_h2(p.ln_p_drop);
  const uint64_t seed = mix_seed(reinterpret_cast<const uint64_t*>(p.ln_seed_dev), p.ln_seed);
  const __nv_bfloat16* res_row = p.ln_res + (long long)t.z * p.res_batch_stride + (long long)(in_rows ? gm : 0) * p.ld_res + colb;
  float* red0 = s_red + (parity * 2 + 0) * 256;  // [half][128]
  float* red1 = s_red + (parity * 2 + 1) * 256;
  // ---- pass 1: v = dropout(bf16(acc + bias)) / (1 - p) + residual; v -> TMEM (fp32) and -> ln_v (bf16)
  float sum = 0.f, sumsq = 0.f;
  uint32_t kbytes[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  uint32_t wv[32];
  // the residual row pieces (64 B per 32 columns, one row per thread) are requested one chunk ahead: their global
  // latency hides behind the tensor-memory load and the Philox rounds of the current chunk
  uint4 rnext[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) rnext[j] = row_ok ? __ldg(reinterpret_cast<const uint4*>(res_row) + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int col0 = colb + c * 32;
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    uint4 rr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rr[j] = rnext[j];
    if (c < 3) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        rnext[j] = row_ok ? __ldg(reinterpret_cast<const uint4*>(res_row + (c + 1) * 32) + j) : make_uint4(0u, 0u, 0u, 0u);
    }
    tmem_ld_wait();
    uint32_t w[16];
    {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(b4 + j);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(__uint_as_float(v[4 * j]) + bv.x, __uint_as_float(v[4 * j + 1]) + bv.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(__uint_as_float(v[4 * j + 2]) + bv.z, __uint_as_float(v[4 * j + 3]) + bv.w);
        w[2 * j] = *reinterpret_cast<const uint32_t*>(&lo);
        w[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&hi);
      }
    }
    if (drop) {
      uint32_t kb = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        KeepMask km;
        km.draw(seed, (uint64_t)R * (C / 8) + ((col0 >> 3) + g), thresh);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[4 * g + i] &= km.m[i];
        kb |= km.to_byte() << (8 * g);
      }
      kbytes[c] = kb;
    }
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(rr);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float f0 = __uint_as_float(w[j] << 16), f1 = __uint_as_float(w[j] & 0xFFFF0000u);
      const float r0 = __uint_as_float(rw[j] << 16), r1 = __uint_as_float(rw[j] & 0xFFFF0000u);
      const float v0 = row_ok ? fmaf(f0, scale, r0) : 0.f, v1 = row_ok ? fmaf(f1, scale, r1) : 0.f;
      sum += v0 + v1;
      sumsq = fmaf(v0, v0, fmaf(v1, v1, sumsq));
      v[2 * j] = __float_as_uint(v0);
      v[2 * j + 1] = __float_as_uint(v1);
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
      wv[(c & 1) * 16 + j] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    tmem_st32(taddr + c * 32, v);
    if (c & 1)
      ln_flush64(stg, lane, wv, p.ln_v + ((long long)t.z * p.M + m_w0) * C + colb + (c - 1) * 32, C, m_w0, p.M);
  }
  tmem_st_wait();
  if (in_rows && p.ln_keep && drop)
    *reinterpret_cast<uint4*>(p.ln_keep + R * (C / 8) + chalf * 16) = make_uint4(kbytes[0], kbytes[1], kbytes[2], kbytes[3]);
  // ---- row statistics from the two column halves: one exchange.  var = E[v^2] - mean^2 in fp32: the rows are
  //      residual-stream activations (|mean| is a fraction of the standard deviation), so the cancellation costs
  //      ~1e-6 of the variance -- and it saves a whole pass over tensor memory and a second barrier per tile
  red0[chalf * 128 + row] = sum;
  red1[chalf * 128 + row] = sumsq;
  bar_sync_epi2();
  const float mean = (red0[row] + red0[128 + row]) * (1.f / C);
  const float var = fmaxf((red1[row] + red1[128 + row]) * (1.f / C) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  if (chalf == 0 && in_rows) {
    p.ln_mean[R] = row_ok ? mean : 0.f;
    p.ln_rstd[R] = row_ok ? rstd : 0.f;
  }
  // ---- pass 2: normalise, affine, zero the padded rows, store
  const float nmr = -mean * rstd;
  const long long base_off = (long long)(t.z / p.d_zdiv) * p.d_zdiv_stride + (long long)(t.z % p.d_zdiv) * p.d_zmod_stride;
  __nv_bfloat16* dbase = static_cast<__nv_bfloat16*>(p.d) + base_off + (long long)m_w0 * p.ldd + colb;
#pragma unroll 1
  for (int c = 0; c < 4; c += 2) {  // 64 columns per tensor-memory round trip
    uint32_t va[32], vb[32];
    tmem_ld32(taddr + c * 32, va);
    tmem_ld32(taddr + c * 32 + 32, vb);
    tmem_ld_wait();
    const float* gam = s_gb + colb + c * 32;
    const float* bet = s_gb + 256 + colb + c * 32;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float o0 = row_ok ? fmaf(fmaf(__uint_as_float(va[2 * j]), rstd, nmr), gam[2 * j], bet[2 * j]) : 0.f;
      const float o1 = row_ok ? fmaf(fmaf(__uint_as_float(va[2 * j + 1]), rstd, nmr), gam[2 * j + 1], bet[2 * j + 1]) : 0.f;
      const float o2 = row_ok ? fmaf(fmaf(__uint_as_float(vb[2 * j]), rstd, nmr), gam[32 + 2 * j], bet[32 + 2 * j]) : 0.f;
      const float o3 = row_ok ? fmaf(fmaf(__uint_as_float(vb[2 * j + 1]), rstd, nmr), gam[32 + 2 * j + 1], bet[32 + 2 * j + 1]) : 0.f;
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(o0, o1), b3 = __floats2bfloat162_rn(o2, o3);
      wv[j] = *reinterpret_cast<const uint32_t*>(&b2);
      wv[16 + j] = *reinterpret_cast<const uint32_t*>(&b3);
    }
    ln_flush64(stg, lane, wv, dbase + c * 32, p.ldd, m_w0, p.M);
  }
}

template <bool LN>
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
  constexpr int NST = LN ? LN_STAGES : STAGES;  // operand ring depth

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.a_colsum_on ? 5 : 1);  // MMA commit (+ four reader warps of the leader CTA)
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
  if (LN && warp == 2) {  // LayerNorm affine parameters -> shared memory (broadcast reads in the epilogue)
    float* gb = reinterpret_cast<float*>(sgen + LN_GB_OFF);
    for (int i = lane; i < 256; i += 32) {
      gb[i] = p.ln_gamma[i];
      gb[256 + i] = p.ln_beta[i];
    }
  }
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
            if (++s == NST) {
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
          if (++s == NST) {
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
          if (++s == NST) {
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
    int as = 0, ln_parity = 0;
    uint32_t aph = 0;
    uint8_t* stg = sgen + STAGING_OFF + (warp - 4) * 4096;
    const uint32_t leader_tempty0 = mapa_rank(tempty_bar(0), 0);
    if (p.ragged && p.mode == FS2_GEMM_NORMAL)
      zero_fill_padded<BN>(p, cum, threadIdx.x - 128, blockIdx.x, gridDim.x);
    // fused column sums of A (WGRAD, fs2_gemm::a_colsum): the scheme of wgrad_taps.cu -- the leader CTA's warps 4-7
    // read its dY tile, 8-11 the peer's (ld.shared::cluster), one extra arrival per warp on the tile's empty barrier;
    // the column tiles of a (row pair tile, split) take turns over its frame blocks
    const bool reader = !LN && p.a_colsum_on && leader;
    const int rtile = (warp - 4) >> 2, rh = (warp - 4) & 1, rf = ((warp - 4) >> 1) & 1;
    const uint32_t rbase = mapa_rank(sbase, (uint32_t)rtile) + rh * kChunkBytes + rf * 32 * 128 + (lane & 3) * 4;
    const uint32_t rempty0 = mapa_rank(empty_bar(0), (uint32_t)rtile);
    int rs = 0;
    uint32_t rph = 0;
    for (int ptile = pair; ptile < total_pair_tiles; ptile += num_pairs) {
      const TileCoord t = decode_pair(ptile);
      if (reader) {
        float c0 = 0.f, c1 = 0.f;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait_cluster(full_bar(rs), rph);
          if ((t.kb0 + kb) % p.tiles_n == t.tn) {
            colsum_tile_rows32(rbase + rs * STAGE_BYTES, lane, c0, c1);  // frames rf * 32 .. + 31 of this tile
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(rempty0 + 8u * rs);
          if (++rs == STAGES) {
            rs = 0;
            rph ^= 1u;
          }
        }
        if (t.nkb > 0) {
          const int co = (t.tm - (int)rank + rtile) * BM + rh * 64 + lane * 2;  // leader: t.tm = 2 * pair tile
          float* dst = p.a_colsum;
          int c = co;
          if (p.seg_rows > 0 && co < p.M) {
            const int sg = co / p.seg_rows;
            dst = p.a_colsum_seg[sg];
            c = co - sg * p.seg_rows;
          }
          if (dst && co < p.M) atomicAdd(dst + c, c0);
          if (dst && co + 1 < p.M) atomicAdd(dst + c + 1, c1);
        }
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (LN) {
        epilogue_ln_tile(p, t, tmem_base + as * BN, stg, q, chalf, lane, reinterpret_cast<const float*>(sgen + LN_GB_OFF),
                         reinterpret_cast<float*>(sgen + LN_RED_OFF), ln_parity);
        ln_parity ^= 1;
      } else if (!(p.dbg & 2)) {
        epilogue_tile<BN>(p, t, tmem_base + as * BN, stg, q, chalf, lane);
      }
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
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
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
  if (g.ln_gamma) {
    if (g.mode != FS2_GEMM_NORMAL || g.N != 256 || g.d_f32 || taps != 1 || g.epilogue != FS2_EPI_NONE || g.alpha != 1.f ||
        !g.ln_beta || !g.ln_res || !g.ln_v || !g.ln_mean || !g.ln_rstd || (g.ld_res & 7) || (g.res_batch_stride & 7) ||
        (reinterpret_cast<uintptr_t>(g.ln_res) & 15) || (reinterpret_cast<uintptr_t>(g.ln_v) & 15) ||
        (g.ln_p_drop > 0.f && (!g.ln_keep || (reinterpret_cast<uintptr_t>(g.ln_keep) & 15))) || g.ln_p_drop >= 1.f)
      return set_error("gemm: the fused LayerNorm epilogue needs NORMAL mode, taps = 1, N = 256, a bf16 output, no other "
                       "epilogue and all ln_* buffers (16-byte aligned)");
    static bool ln_attr = false;
    if (!ln_attr) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
      if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(gemm_tc2<LN>)", e);
      ln_attr = true;
    }
    kp.ln_gamma = g.ln_gamma; kp.ln_beta = g.ln_beta;
    kp.ln_res = static_cast<const __nv_bfloat16*>(g.ln_res);
    kp.ld_res = g.ld_res; kp.res_batch_stride = g.res_batch_stride;
    kp.ln_p_drop = g.ln_p_drop; kp.ln_seed = g.ln_seed;
    kp.ln_seed_dev = reinterpret_cast<const unsigned long long*>(g.ln_seed_dev);
    kp.ln_v = static_cast<__nv_bfloat16*>(g.ln_v);
    kp.ln_mean = g.ln_mean; kp.ln_rstd = g.ln_rstd; kp.ln_keep = g.ln_keep;
    FS2_LAUNCH((gemm_tc2_kernel<true>), 2 * pairs, kThreads2, DYN_BYTES, stream, tmA, tmB, kp);
    count_launch();
    return check_launch("gemm_tc2_kernel<LN>");
  }
  FS2_LAUNCH((gemm_tc2_kernel<false>), 2 * pairs, kThreads2, DYN_BYTES, stream, tmA, tmB, kp);
  count_launch();
  return check_launch("gemm_tc2_kernel");
}

}  // namespace fs2
