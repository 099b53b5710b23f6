// FastSpeech2Loss: masked L1 on mel / postnet mel, masked MSE on pitch / energy / log-duration,
// forward (deterministic two-stage reduction in ONE kernel: the last block to arrive sums the per-block partials in
// block order) and backward.
//
// Replaces lightning/model/loss.py:15-89: the nine `masked_select` compactions, two nn.L1Loss and
// three nn.MSELoss (each a mean over its OWN number of valid elements), the `log(d + 1)` target
// (loss.py:38), the slicing of the mel target to the (possibly truncated) mask length (loss.py:39)
// and the unweighted total (loss.py:77-79).  Validity comes from the lengths instead of bool masks:
//   mel frame (b,t) valid  iff t < min(mel_lens[b], Tm)      (Tm = decoder output length)
//   phoneme  (b,i) valid  iff i < min(src_lens[b], Ts)
// HBM-bound: algorithmic bytes fwd = 3*B*Tm*n_mel*4 + 5*B*Ts*4, bwd adds 2*B*Tm*n_mel*4 of writes.
#include "common.h"
#include "util.cuh"

namespace fs2 {

constexpr int kLossThreads = 256;
constexpr int kElemsPerBlock = kLossThreads * 4 * 4;  // 4 float4 per thread

struct LossArgs {
  const float* mel_pred;
  const float* post_pred;
  const float* mel_tgt;
  const float* p_pred;
  const float* p_tgt;
  const float* e_pred;
  const void* e_tgt;
  int e_tgt_is_f64;
  const float* d_pred;
  const int64_t* d_tgt;
  const int64_t* src_lens;
  const int64_t* mel_lens;
  int B, Ts, Tm, Tm_tgt, n_mel;
  int p_T, p_ld, e_T, e_ld;   // pitch / energy: prediction row length, target row stride
  const int64_t* p_lens;      // src_lens (phoneme level) or mel_lens (frame level)
  const int64_t* e_lens;
  int mel_blocks_per_b, n_blocks;
  float* partials;  // [5][n_blocks] + arrival counter (zero between calls)
  float* out;       // [10] total, mel, post, pitch, energy, duration, N_mel, N_pitch, N_energy, N_duration
  // backward
  const float* gout;  // [6] d(loss)/d(out[0..5])
  float* d_mel;
  float* d_post;
  float* d_p;
  float* d_e;
  float* d_d;
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

__device__ __forceinline__ float energy_target(const LossArgs& a, long long i) {
  return a.e_tgt_is_f64 ? static_cast<float>(static_cast<const double*>(a.e_tgt)[i])
                        : static_cast<const float*>(a.e_tgt)[i];
}
// number of valid positions of a [B][T] feature, counted by ONE WARP (exact integer sum; every lane returns it)
__device__ __forceinline__ double count_valid_warp(const int64_t* lens, int B, int T) {
  long long n = 0;
  for (int b = threadIdx.x & 31; b < B; b += 32) {
    const long long l = min((long long)lens[b], (long long)T);
    n += l > 0 ? l : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  return (double)n;
}

__device__ __forceinline__ void loss_final(const LossArgs& a, float* sh) {
  __shared__ double cnt[4];
  // the four normalisers, one warp each (this used to be 4 * B dependent loads of thread 0: 15 us at B = 64)
  const int warp = threadIdx.x >> 5;
  if (warp < 4) {
    const double c = warp == 0 ? count_valid_warp(a.mel_lens, a.B, a.Tm) * a.n_mel
                   : warp == 1 ? count_valid_warp(a.p_lens, a.B, a.p_T)
                   : warp == 2 ? count_valid_warp(a.e_lens, a.B, a.e_T)
                               : count_valid_warp(a.src_lens, a.B, a.Ts);
    if ((threadIdx.x & 31) == 0) cnt[warp] = c;
  }
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i0 = threadIdx.x; i0 < a.n_blocks; i0 += 4 * kLossThreads) {  // 20 loads in flight per thread
    float v[4][5];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kLossThreads;
#pragma unroll
      for (int k = 0; k < 5; ++k) v[u][k] = i < a.n_blocks ? __ldcg(a.partials + (long long)k * a.n_blocks + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 5; ++k) s[k] += v[u][k];
  }
  float r[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) r[k] = block_sum(s[k], sh);  // (its barriers also publish cnt[])
  if (threadIdx.x == 0) {
    const double n_mel = cnt[0], n_p = cnt[1], n_e = cnt[2], n_d = cnt[3];
    const float mel = r[0] / (float)n_mel, post = r[1] / (float)n_mel;
    const float pitch = r[2] / (float)n_p, energy = r[3] / (float)n_e, dur = r[4] / (float)n_d;
    a.out[0] = mel + post + dur + pitch + energy;  // same association order as loss.py:77-79
    a.out[1] = mel;
    a.out[2] = post;
    a.out[3] = pitch;
    a.out[4] = energy;
    a.out[5] = dur;
    a.out[6] = (float)n_mel;
    a.out[7] = (float)n_p;
    a.out[8] = (float)n_e;
    a.out[9] = (float)n_d;
  }
}

__global__ void __launch_bounds__(kLossThreads) loss_partial_kernel(const LossArgs a) {
  pdl_sync();
  __shared__ float sh[8];
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int nmel_blocks = a.B * a.mel_blocks_per_b;
  if ((int)blockIdx.x < nmel_blocks) {
    const int b = blockIdx.x / a.mel_blocks_per_b, chunk = blockIdx.x % a.mel_blocks_per_b;
    const long long len = min((long long)a.mel_lens[b], (long long)a.Tm);
    const long long valid = (len > 0 ? len : 0) * a.n_mel;
    const float* mp = a.mel_pred + (long long)b * a.Tm * a.n_mel;
    const float* pp = a.post_pred + (long long)b * a.Tm * a.n_mel;
    const float* tp = a.mel_tgt + (long long)b * a.Tm_tgt * a.n_mel;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = (long long)chunk * kElemsPerBlock + (u * kLossThreads + threadIdx.x) * 4;
      if (i < valid) {  // valid is a multiple of n_mel; n_mel % 4 == 0 checked on the host
        const float4 m = *reinterpret_cast<const float4*>(mp + i);
        const float4 p = *reinterpret_cast<const float4*>(pp + i);
        const float4 t = *reinterpret_cast<const float4*>(tp + i);
        s[0] += (fabsf(m.x - t.x) + fabsf(m.y - t.y)) + (fabsf(m.z - t.z) + fabsf(m.w - t.w));
        s[1] += (fabsf(p.x - t.x) + fabsf(p.y - t.y)) + (fabsf(p.z - t.z) + fabsf(p.w - t.w));
      }
    }
  } else {
    const int blk = blockIdx.x - nmel_blocks;
    const long long stride = (long long)(a.n_blocks - nmel_blocks) * kLossThreads;
    const long long i0 = (long long)blk * kLossThreads + threadIdx.x;
    for (long long i = i0; i < (long long)a.B * a.p_T; i += stride) {
      const int b = i / a.p_T, t = i - (long long)b * a.p_T;
      if (t < a.p_lens[b]) {
        const float dp = a.p_pred[i] - a.p_tgt[(long long)b * a.p_ld + t];
        s[2] += dp * dp;
      }
    }
    for (long long i = i0; i < (long long)a.B * a.e_T; i += stride) {
      const int b = i / a.e_T, t = i - (long long)b * a.e_T;
      if (t < a.e_lens[b]) {
        const float de = a.e_pred[i] - energy_target(a, (long long)b * a.e_ld + t);
        s[3] += de * de;
      }
    }
    for (long long i = i0; i < (long long)a.B * a.Ts; i += stride) {
      const int b = i / a.Ts, t = i - (long long)b * a.Ts;
      if (t < a.src_lens[b]) {
        const float dd = a.d_pred[i] - logf(static_cast<float>(a.d_tgt[i]) + 1.f);
        s[4] += dd * dd;
      }
    }
  }
  // block sums of the five terms with one barrier pair, then the fixed-order final reduction by whichever block
  // arrives last (partials are [5][n_blocks]; the arrival counter behind them is left at zero for the next call)
  __shared__ float sh5[5][kLossThreads / 32];
  __shared__ bool s_last;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float r = warp_sum(s[k]);
      if (lane == 0) sh5[k][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < kLossThreads / 32; ++w) r += sh5[threadIdx.x][w];
      a.partials[(long long)threadIdx.x * a.n_blocks + blockIdx.x] = r;
      __threadfence();
    }
    __syncwarp();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned* counter = reinterpret_cast<unsigned*>(a.partials + 5ll * a.n_blocks);
      s_last = atomicAdd(counter, 1u) == (unsigned)a.n_blocks - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
  }
  loss_final(a, sh);
  if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(a.partials + 5ll * a.n_blocks) = 0u;
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(kLossThreads) loss_bwd_kernel(const LossArgs a) {
  pdl_sync();
  const float g0 = a.gout[0];
  const float n_mel = a.out[6];
  const int nmel_blocks = a.B * a.mel_blocks_per_b;
  if ((int)blockIdx.x < nmel_blocks) {
    const float wm = (g0 + a.gout[1]) / n_mel, wp = (g0 + a.gout[2]) / n_mel;
    const int b = blockIdx.x / a.mel_blocks_per_b, chunk = blockIdx.x % a.mel_blocks_per_b;
    const long long len = min((long long)a.mel_lens[b], (long long)a.Tm);
    const long long valid = (len > 0 ? len : 0) * a.n_mel;
    const long long total = (long long)a.Tm * a.n_mel;
    const long long base = (long long)b * total;
    const float* tp = a.mel_tgt + (long long)b * a.Tm_tgt * a.n_mel;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = (long long)chunk * kElemsPerBlock + (u * kLossThreads + threadIdx.x) * 4;
      if (i >= total) continue;
      float4 dm = make_float4(0.f, 0.f, 0.f, 0.f), dp = dm;
      if (i < valid) {
        const float4 m = *reinterpret_cast<const float4*>(a.mel_pred + base + i);
        const float4 p = *reinterpret_cast<const float4*>(a.post_pred + base + i);
        const float4 t = *reinterpret_cast<const float4*>(tp + i);
        dm = make_float4(wm * sgn(m.x - t.x), wm * sgn(m.y - t.y), wm * sgn(m.z - t.z),
                         wm * sgn(m.w - t.w));
        dp = make_float4(wp * sgn(p.x - t.x), wp * sgn(p.y - t.y), wp * sgn(p.z - t.z),
                         wp * sgn(p.w - t.w));
      }
      *reinterpret_cast<float4*>(a.d_mel + base + i) = dm;
      *reinterpret_cast<float4*>(a.d_post + base + i) = dp;
    }
  } else {
    const float w_p = 2.f * (g0 + a.gout[3]) / a.out[7], w_e = 2.f * (g0 + a.gout[4]) / a.out[8],
                w_d = 2.f * (g0 + a.gout[5]) / a.out[9];
    const int blk = blockIdx.x - nmel_blocks;
    const long long stride = (long long)(a.n_blocks - nmel_blocks) * kLossThreads;
    const long long i0 = (long long)blk * kLossThreads + threadIdx.x;
    for (long long i = i0; i < (long long)a.B * a.p_T; i += stride) {
      const int b = i / a.p_T, t = i - (long long)b * a.p_T;
      a.d_p[i] = t < a.p_lens[b] ? w_p * (a.p_pred[i] - a.p_tgt[(long long)b * a.p_ld + t]) : 0.f;
    }
    for (long long i = i0; i < (long long)a.B * a.e_T; i += stride) {
      const int b = i / a.e_T, t = i - (long long)b * a.e_T;
      a.d_e[i] = t < a.e_lens[b] ? w_e * (a.e_pred[i] - energy_target(a, (long long)b * a.e_ld + t)) : 0.f;
    }
    for (long long i = i0; i < (long long)a.B * a.Ts; i += stride) {
      const int b = i / a.Ts, t = i - (long long)b * a.Ts;
      a.d_d[i] = t < a.src_lens[b] ? w_d * (a.d_pred[i] - logf(static_cast<float>(a.d_tgt[i]) + 1.f)) : 0.f;
    }
  }
}

static int fill_grid(LossArgs& a) {
  if (a.n_mel % 4) return set_error("loss: n_mel must be a multiple of 4");
  if (a.Tm > a.Tm_tgt) return set_error("loss: mel target shorter than the prediction");
  if (a.p_T > a.p_ld || a.e_T > a.e_ld) return set_error("loss: pitch / energy target shorter than the prediction");
  const long long per_b = (long long)a.Tm * a.n_mel;
  a.mel_blocks_per_b = (int)((per_b + kElemsPerBlock - 1) / kElemsPerBlock);
  long long nsrc = (long long)a.B * a.Ts;
  if ((long long)a.B * a.p_T > nsrc) nsrc = (long long)a.B * a.p_T;
  if ((long long)a.B * a.e_T > nsrc) nsrc = (long long)a.B * a.e_T;
  int src_blocks = (int)((nsrc + kLossThreads - 1) / kLossThreads);
  if (src_blocks > 64) src_blocks = 64;
  if (src_blocks < 1) src_blocks = 1;
  a.n_blocks = a.B * a.mel_blocks_per_b + src_blocks;
  return 0;
}

}  // namespace fs2

extern "C" {

// Number of floats the caller must provide in `partials` (T_feat = the longest of Ts and the pitch / energy rows).
int64_t fs2_loss_workspace_floats(int B, int T_feat, int Tm, int n_mel) {
  fs2::LossArgs a{};
  a.B = B; a.Ts = T_feat; a.Tm = Tm; a.Tm_tgt = Tm; a.n_mel = n_mel;
  if (fs2::fill_grid(a)) return -1;
  return (int64_t)a.n_blocks * 5 + 4;  // + the arrival counter (and padding)
}

static void loss_common(fs2::LossArgs& a, const float* mel_pred, const float* post_pred, const float* mel_tgt,
                        const float* p_pred, const float* p_tgt, int p_T, int p_ld, const int64_t* p_lens,
                        const float* e_pred, const void* e_tgt, int e_tgt_is_f64, int e_T, int e_ld,
                        const int64_t* e_lens, const float* d_pred, const int64_t* d_tgt, const int64_t* src_lens,
                        const int64_t* mel_lens, int B, int Ts, int Tm, int Tm_tgt, int n_mel) {
  a.mel_pred = mel_pred; a.post_pred = post_pred; a.mel_tgt = mel_tgt;
  a.p_pred = p_pred; a.p_tgt = p_tgt; a.e_pred = e_pred; a.e_tgt = e_tgt;
  a.e_tgt_is_f64 = e_tgt_is_f64; a.d_pred = d_pred; a.d_tgt = d_tgt;
  a.src_lens = src_lens; a.mel_lens = mel_lens;
  a.p_T = p_T; a.p_ld = p_ld; a.p_lens = p_lens; a.e_T = e_T; a.e_ld = e_ld; a.e_lens = e_lens;
  a.B = B; a.Ts = Ts; a.Tm = Tm; a.Tm_tgt = Tm_tgt; a.n_mel = n_mel;
}

int fs2_loss_fwd(const float* mel_pred, const float* post_pred, const float* mel_tgt, const float* p_pred,
                 const float* p_tgt, int p_T, int p_ld, const int64_t* p_lens, const float* e_pred, const void* e_tgt,
                 int e_tgt_is_f64, int e_T, int e_ld, const int64_t* e_lens, const float* d_pred,
                 const int64_t* d_tgt, const int64_t* src_lens, const int64_t* mel_lens, int B, int Ts, int Tm,
                 int Tm_tgt, int n_mel, float* partials, float* out10, void* stream) {
  fs2::LossArgs a{};
  loss_common(a, mel_pred, post_pred, mel_tgt, p_pred, p_tgt, p_T, p_ld, p_lens, e_pred, e_tgt, e_tgt_is_f64, e_T, e_ld,
              e_lens, d_pred, d_tgt, src_lens, mel_lens, B, Ts, Tm, Tm_tgt, n_mel);
  a.partials = partials; a.out = out10;
  if (int rc = fs2::fill_grid(a)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FS2_LAUNCH((fs2::loss_partial_kernel), a.n_blocks, fs2::kLossThreads, 0, s, a);
  fs2::count_launch();
  return fs2::check_launch("loss_partial_kernel");
}

int fs2_loss_bwd(const float* gout6, const float* out10, const float* mel_pred, const float* post_pred,
                 const float* mel_tgt, const float* p_pred, const float* p_tgt, int p_T, int p_ld,
                 const int64_t* p_lens, const float* e_pred, const void* e_tgt, int e_tgt_is_f64, int e_T, int e_ld,
                 const int64_t* e_lens, const float* d_pred, const int64_t* d_tgt, const int64_t* src_lens,
                 const int64_t* mel_lens, int B, int Ts, int Tm, int Tm_tgt, int n_mel, float* d_mel, float* d_post,
                 float* d_p, float* d_e, float* d_d, void* stream) {
  fs2::LossArgs a{};
  loss_common(a, mel_pred, post_pred, mel_tgt, p_pred, p_tgt, p_T, p_ld, p_lens, e_pred, e_tgt, e_tgt_is_f64, e_T, e_ld,
              e_lens, d_pred, d_tgt, src_lens, mel_lens, B, Ts, Tm, Tm_tgt, n_mel);
  a.gout = gout6; a.out = const_cast<float*>(out10);
  a.d_mel = d_mel; a.d_post = d_post; a.d_p = d_p; a.d_e = d_e; a.d_d = d_d;
  if (int rc = fs2::fill_grid(a)) return rc;
  FS2_LAUNCH((fs2::loss_bwd_kernel), a.n_blocks, fs2::kLossThreads, 0, static_cast<cudaStream_t>(stream), a);
  fs2::count_launch();
  return fs2::check_launch("loss_bwd_kernel");
}
}
