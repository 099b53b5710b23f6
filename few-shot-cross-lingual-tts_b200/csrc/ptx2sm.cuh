// cta_group::2 / thread-block-cluster PTX wrappers shared by the 2-CTA GEMM (gemm_tc2.cu) and the 2-CTA
// implicit-GEMM Conv1d with operand-halo reuse (conv_tc2.cu).
#pragma once
#include <cstdint>
#include <cstdio>

namespace fs2 {

// ---- cluster / cta_group::2 PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// generic-proxy reads of a peer CTA's shared memory whose TMA fill was signalled on OUR mbarrier: wait with
// cluster-scope acquire, then ld.shared::cluster through a mapa address (wgrad_taps.cu: fused column sums)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) {
      printf("fs2: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}
// Column sums of one MN-major operand tile (two 64 x 64 SWIZZLE_128B boxes, a row = 64 bf16 of one reduction row): this
// lane adds the bf16 pair at column pair `lane` of 32 consecutive rows starting at `a` (the cluster address of the
// lane's pair in the first of those rows).  All 32 loads are in flight at once: one (distributed) shared-memory
// latency per tile.  Used by the fused bias gradients of wgrad_taps.cu and gemm_tc2.cu.
__device__ __forceinline__ void colsum_tile_rows32(uint32_t a, int lane, float& c0, float& c1) {
  uint32_t v[32];
#pragma unroll
  for (int f = 0; f < 32; ++f)  // 16-byte chunk index ^ (row & 7): the rows start at a multiple of 8
    v[f] = ld_shared_cluster_u32(a + f * 128 + ((((uint32_t)lane >> 2) ^ (f & 7)) << 4));
#pragma unroll
  for (int f = 0; f < 32; ++f) {
    c0 += __uint_as_float(v[f] << 16);
    c1 += __uint_as_float(v[f] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs once the previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}

}  // namespace fs2
