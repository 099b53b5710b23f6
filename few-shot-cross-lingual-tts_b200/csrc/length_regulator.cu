// LengthRegulator as an integer inclusive cumsum + a vectorised row gather (forward) and a
// contiguous segment sum (backward).
//
// Replaces lightning/model/modules.py:169-196 (python double loop with one `.item()` host sync per
// phoneme, `vec.expand`, `torch.cat`) and `pad` (lightning/utils/tool.py:168-186).  Semantics kept:
//   * each duration goes through max(int(d), 0) (modules.py:188-189) -- int() truncates toward 0;
//   * out[b, t] = x[b, i] where i = #{j : cum[b, j] <= t}, zero rows for t >= cum[b, Ts-1];
//   * max_len smaller than the expansion CROPS (negative F.pad), larger zero-pads;
//   * mel_len[b] = sum_i max(int(d[b,i]),0) (NOT cropped).
// The gather is a pure copy, so the output is bit-exact for any element type.  Optional fused adds
// (speaker row, sinusoid table row: fastspeech2m.py:132-136, transformer/Models.py:224-226) are only
// used by the fused FastSpeech2 forward, never by the stand-alone LengthRegulator module.
// HBM-bound: algorithmic bytes = (B*Ts + B*max_len) * C * sizeof(elt) + 8*B*Ts.
#include "common.h"
#include "util.cuh"

namespace fs2 {

// One block per utterance: clamp + inclusive scan of durations, then idx for every output frame.
template <typename DurT>
__global__ void __launch_bounds__(256)
lr_index_kernel(const DurT* __restrict__ dur, int Ts, int max_len, int64_t* __restrict__ cum,
                int32_t* __restrict__ idx, int64_t* __restrict__ mel_len) {
  pdl_sync();
  extern __shared__ long long s_cum[];  // Ts entries
  __shared__ long long s_warp[8];
  __shared__ long long s_carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < Ts; base += 256) {
    const int i = base + tid;
    long long v = 0;
    if (i < Ts) {
      const DurT d = dur[(long long)b * Ts + i];
      v = static_cast<long long>(d);  // float -> truncation toward zero == python int()
      if (v < 0) v = 0;
    }
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    long long off = s_carry;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (i < Ts) {
      s_cum[i] = x + off;
      cum[(long long)b * Ts + i] = x + off;
    }
    __syncthreads();
    if (tid == 255) s_carry = x + off;
    __syncthreads();
  }
  const long long total = Ts > 0 ? s_carry : 0;
  if (tid == 0) mel_len[b] = total;
  for (int t = tid; t < max_len; t += 256) {
    int r = -1;
    if (t < total) {
      int lo = 0, hi = Ts;  // first i with cum[i] > t
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_cum[mid] <= t) lo = mid + 1; else hi = mid;
      }
      r = lo;
    }
    idx[(long long)b * max_len + t] = r;
  }
}

// One warp per output row; VEC 16-byte vectors per row are spread over the lanes.
__global__ void __launch_bounds__(256)
lr_gather_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ idx, int B, int Ts,
                 int max_len, int out_len, int vec_per_row, uint4* __restrict__ out) {
  pdl_sync();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * out_len) return;
  const int lane = threadIdx.x & 31;
  const int b = row / out_len, t = row - (long long)b * out_len;
  const int i = idx[(long long)b * max_len + t];
  const uint4* src = i >= 0 ? x + ((long long)b * Ts + i) * vec_per_row : nullptr;
  uint4* dst = out + row * vec_per_row;
  for (int v = lane; v < vec_per_row; v += 32) dst[v] = src ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
}

// bf16 gather fused with "+ speaker row + sinusoid row" (decoder input of the fused forward).
// One warp owns ONE frame index t for a group of kLrGroup utterances: the fp32 sinusoid row of t is read
// once into registers and reused for every utterance of the group (it is 2x the bytes of the bf16 output
// row).  The group's row indices are fetched by one load (lane u holds utterance b0+u's) so that the gathers
// depend on a single latency, four gathered rows are in flight per lane, and the group's fp32 speaker rows are
// staged in shared memory once per block (read per row from L1/L2 they were a 16-step dependent latency chain:
// 19 -> 22 us for a 40 MB kernel).
constexpr int kLrGroup = 16;
constexpr int kLrFlight = 4;
constexpr int kLrSpkSmemMax = 48 * 1024;
__global__ void __launch_bounds__(256, 4)
lr_gather_fused_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ idx,
                       const float* __restrict__ spk, const float* __restrict__ pe, int B, int Ts,
                       int max_len, int out_len, int C, int spk_smem, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  extern __shared__ __align__(16) float s_spk[];  // [kLrGroup][C] when spk_smem
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * kLrGroup;
  const int b1 = min(b0 + kLrGroup, B);
  if (spk && spk_smem) {
    const float4* src = reinterpret_cast<const float4*>(spk + (long long)b0 * C);
    for (int v = threadIdx.x; v < (b1 - b0) * (C >> 2); v += blockDim.x)
      reinterpret_cast<float4*>(s_spk)[v] = __ldg(src + v);
    __syncthreads();
  }
  if (t >= out_len) return;
  const int my_i = (lane < kLrGroup && b0 + lane < b1) ? idx[(long long)(b0 + lane) * max_len + t] : -1;
  for (int c = lane * 8; c < C; c += 256) {
    float p8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p8[j] = 0.f;
    if (pe) {
      const float4 p0 = *reinterpret_cast<const float4*>(pe + (long long)t * C + c);
      const float4 p1 = *reinterpret_cast<const float4*>(pe + (long long)t * C + c + 4);
      p8[0] = p0.x; p8[1] = p0.y; p8[2] = p0.z; p8[3] = p0.w; p8[4] = p1.x; p8[5] = p1.y; p8[6] = p1.z; p8[7] = p1.w;
    }
    for (int bb = b0; bb < b1; bb += kLrFlight) {
      int i[kLrFlight];
      bf16x8 xv[kLrFlight];
#pragma unroll
      for (int u = 0; u < kLrFlight; ++u) i[u] = __shfl_sync(0xffffffffu, my_i, (bb - b0 + u) & 31);
#pragma unroll
      for (int u = 0; u < kLrFlight; ++u)
        if (bb + u < b1 && i[u] >= 0) xv[u] = ld8(x + ((long long)(bb + u) * Ts + i[u]) * C + c);
#pragma unroll
      for (int u = 0; u < kLrFlight; ++u) {
        const int b = bb + u;
        if (b >= b1) break;
        float f[8];
        if (i[u] >= 0) {
          unpack8(xv[u], f);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
        if (spk) {
          // the reference adds in two rounded steps (x + spk, then + pe), each in fp32
          const float* sp = spk_smem ? s_spk + (b - b0) * C + c : spk + (long long)b * C + c;
          const float4 s0 = *reinterpret_cast<const float4*>(sp);
          const float4 s1 = *reinterpret_cast<const float4*>(sp + 4);
          f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w; f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
        }
        if (pe) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += p8[j];
        }
        st8(out + ((long long)b * out_len + t) * C + c, pack8(f));
      }
    }
  }
}

// dx[b,i,:] = sum_{t in [cum[i-1], min(cum[i], n_rows))} dout[b,t,:]   (contiguous, deterministic)
__global__ void __launch_bounds__(256)
lr_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const int64_t* __restrict__ cum, int B, int Ts,
              int n_rows, int C, __nv_bfloat16* __restrict__ dx) {
  pdl_sync();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * Ts) return;
  const int lane = threadIdx.x & 31;
  const int b = row / Ts, i = row - (long long)b * Ts;
  long long t0 = i > 0 ? cum[row - 1] : 0, t1 = cum[row];
  if (t1 > n_rows) t1 = n_rows;
  for (int c = lane * 8; c < C; c += 256) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long t = t0; t < t1; t += 4) {  // four frames in flight, added in frame order
      bf16x8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t + u < t1) v[u] = ld8(dout + ((long long)b * n_rows + t + u) * C + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (t + u >= t1) break;
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    st8(dx + row * C + c, pack8(acc));
  }
}

__global__ void __launch_bounds__(256)
lr_bwd_f32_kernel(const float* __restrict__ dout, const int64_t* __restrict__ cum, int B, int Ts,
                  int n_rows, int C, float* __restrict__ dx) {
  pdl_sync();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * Ts) return;
  const int lane = threadIdx.x & 31;
  const int b = row / Ts, i = row - (long long)b * Ts;
  long long t0 = i > 0 ? cum[row - 1] : 0, t1 = cum[row];
  if (t1 > n_rows) t1 = n_rows;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (long long t = t0; t < t1; ++t) acc += dout[((long long)b * n_rows + t) * C + c];
    dx[row * C + c] = acc;
  }
}

}  // namespace fs2

extern "C" {

// dur: int64 [B][Ts] (dur_is_f32 = 0) or f32 (inference path, modules.py:133-139).
// Outputs: cum int64 [B][Ts] (inclusive clamped cumsum), idx int32 [B][max_len] (-1 = zero row),
// mel_len int64 [B].
int fs2_lr_index(const void* dur, int dur_is_f32, int B, int Ts, int max_len, int64_t* cum,
                 int32_t* idx, int64_t* mel_len, void* stream) {
  if (B <= 0) return 0;
  if (Ts > 12000) return fs2::set_error("lr_index: Ts too large for the shared-memory scan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)(Ts > 0 ? Ts : 1) * sizeof(long long);
  if (dur_is_f32)
    FS2_LAUNCH((fs2::lr_index_kernel<float>), B, 256, smem, s, static_cast<const float*>(dur), Ts, max_len, cum,
                                                     idx, mel_len);
  else
    FS2_LAUNCH((fs2::lr_index_kernel<int64_t>), B, 256, smem, s, static_cast<const int64_t*>(dur), Ts, max_len,
                                                       cum, idx, mel_len);
  fs2::count_launch();
  return fs2::check_launch("lr_index_kernel");
}

// Pure copy gather; row_bytes = C * sizeof(element) must be a multiple of 16.
// out: [B][out_len][C] with out_len <= max_len (the decoder truncates to max_seq_len).
int fs2_lr_gather(const void* x, const int32_t* idx, int B, int Ts, int max_len, int out_len,
                  int row_bytes, void* out, void* stream) {
  if (row_bytes % 16) return fs2::set_error("lr_gather: row bytes must be a multiple of 16");
  const long long rows = (long long)B * out_len;
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::lr_gather_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(x), idx, B, Ts, max_len, out_len, row_bytes / 16,
      static_cast<uint4*>(out));
  fs2::count_launch();
  return fs2::check_launch("lr_gather_kernel");
}

// bf16 gather + optional f32 speaker row [B][C] + optional f32 position table [>=out_len][C].
int fs2_lr_gather_fused_bf16(const void* x, const int32_t* idx, const float* spk, const float* pe,
                             int B, int Ts, int max_len, int out_len, int C, void* out, void* stream) {
  if (C % 8) return fs2::set_error("lr_gather_fused: C must be a multiple of 8");
  const long long rows = (long long)B * out_len;
  if (rows <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(spk) & 15) || (reinterpret_cast<uintptr_t>(pe) & 15) || (C % 4))
    return fs2::set_error("lr_gather_fused: spk / pe rows must be 16-byte aligned");
  const dim3 grid((unsigned)((out_len + 7) / 8), (unsigned)((B + fs2::kLrGroup - 1) / fs2::kLrGroup));
  const size_t spk_bytes = spk ? (size_t)fs2::kLrGroup * C * sizeof(float) : 0;
  const int spk_smem = spk_bytes > 0 && spk_bytes <= (size_t)fs2::kLrSpkSmemMax;
  FS2_LAUNCH((fs2::lr_gather_fused_kernel), grid, 256, spk_smem ? spk_bytes : 0, static_cast<cudaStream_t>(stream),
      static_cast<const __nv_bfloat16*>(x), idx, spk, pe, B, Ts, max_len, out_len, C, spk_smem,
      static_cast<__nv_bfloat16*>(out));
  fs2::count_launch();
  return fs2::check_launch("lr_gather_fused_kernel");
}

// dout: bf16 [B][n_rows][C] (n_rows = rows that exist in the forward output) -> dx bf16 [B][Ts][C].
int fs2_lr_bwd_bf16(const void* dout, const int64_t* cum, int B, int Ts, int n_rows, int C, void* dx,
                    void* stream) {
  if (C % 8) return fs2::set_error("lr_bwd: C must be a multiple of 8");
  const long long rows = (long long)B * Ts;
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::lr_bwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(dout), cum, B, Ts, n_rows, C, static_cast<__nv_bfloat16*>(dx));
  fs2::count_launch();
  return fs2::check_launch("lr_bwd_kernel");
}

int fs2_lr_bwd_f32(const float* dout, const int64_t* cum, int B, int Ts, int n_rows, int C, float* dx,
                   void* stream) {
  const long long rows = (long long)B * Ts;
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::lr_bwd_f32_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      dout, cum, B, Ts, n_rows, C, dx);
  fs2::count_launch();
  return fs2::check_launch("lr_bwd_f32_kernel");
}
}
