// Device helpers shared by the memory-bound kernels: 16-byte bf16 vectors, warp/block reductions,
// Philox4x32-7 counter RNG for dropout and the bf16x2 keep-mask form the LayerNorm kernels use.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace fs2 {

constexpr int kWarp = 32;

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
// Always ONE 16-byte access: copying the struct member-wise (__nv_bfloat162 has user-provided copy operations, so
// bf16x8 is not trivially copyable) let the compiler emit four 32-bit LDG / STG at a 16-byte lane stride -- four
// times the LSU requests and partially written sectors -- in most of the memory-bound kernels.
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) {
  bf16x8 r;
  *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-7 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3": 7 rounds is the smallest
// crush-resistant variant; the default 10 is a safety margin), one call -> 4 x 32 random bits.  The dropout masks
// are regenerated (never stored) in the backward kernels, and these rounds are a visible share of the otherwise
// HBM-bound LayerNorm / BatchNorm kernels' instruction stream.
constexpr int kPhiloxRounds = 7;
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32), c2 = 0x243F6A88u,
           c3 = 0x85A308D3u;
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Dropout keep-mask for 8 consecutive elements starting at linear index `elem0` (multiple of 8).
// bit i of the result = keep element elem0+i.  `thresh` = p * 2^16 (16 random bits per element).
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t elem0, uint32_t thresh) {
  const uint4 r = philox4x32(seed, elem0 >> 3);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xFFFFu) >= thresh ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= thresh ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}
// ---- dropout masks in bf16x2 lane form ------------------------------------------------------------------------
// One Philox call gives 8 x 16 random bits for the 8 channels of a lane.  A pair of channels is kept / dropped by
// ONE half2 compare on 13 of its 16 bits (mapped into [2, 4): finite, normal fp16 numbers, ordered like the
// integers) that yields 0xFFFF / 0x0000 per half -- the mask is ANDed onto the packed bf16 pair, and the 1/(1-p)
// scale rides on the FMA that adds the residual.  (The per-element form -- extract, compare, select, multiply --
// was ~6 instructions per element of an otherwise HBM-bound kernel.)  Resolution of p: 1/8192.
struct KeepMask {
  uint32_t m[4];  // channel pairs (0,1) (2,3) (4,5) (6,7): 0xFFFF per kept half
  __device__ __forceinline__ void draw(uint64_t seed, uint64_t ctr, uint32_t thresh_h2) {
    const uint4 r = philox4x32(seed, ctr);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const __half2 th = *reinterpret_cast<const __half2*>(&thresh_h2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = (w[i] & 0x1FFF1FFFu) | 0x40004000u;
      m[i] = __hge2_mask(*reinterpret_cast<const __half2*>(&a), th);
    }
  }
  __device__ __forceinline__ void all() { m[0] = m[1] = m[2] = m[3] = 0xFFFFFFFFu; }
  __device__ __forceinline__ uint32_t to_byte() const {
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) b |= ((m[i] & 1u) | ((m[i] >> 15) & 2u)) << (2 * i);
    return b;
  }
  __device__ __forceinline__ void from_byte(uint32_t b) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t t = b >> (2 * i);
      m[i] = ((t & 1u) | ((t & 2u) << 15)) * 0xFFFFu;
    }
  }
  __device__ __forceinline__ void apply(bf16x8& v) const {
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] &= m[i];
  }
};
// threshold pair for KeepMask::draw: keep <=> 13 random bits >= p * 8192
__host__ __device__ __forceinline__ uint32_t keep_thresh_h2(float p) {
  float t = p * 8192.f;
  const uint32_t k = t <= 0.f ? 0u : (t >= 8191.f ? 8191u : static_cast<uint32_t>(t + 0.5f));
  const uint32_t h = 0x4000u | k;
  return h | (h << 16);
}


// Per-step seed: a device-resident counter (so that a replayed CUDA graph draws fresh masks every
// step) mixed with a per-call-site salt.
__device__ __forceinline__ uint64_t mix_seed(const uint64_t* seed_dev, uint64_t salt) {
  const uint64_t s = seed_dev ? *seed_dev : 0ull;
  return s * 0x9E3779B97F4A7C15ull + salt;
}
__host__ __device__ __forceinline__ uint32_t dropout_thresh(float p) {
  float t = p * 65536.f;
  return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : static_cast<uint32_t>(t));
}

}  // namespace fs2
