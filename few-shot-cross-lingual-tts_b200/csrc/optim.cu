// Fused gradient clipping + Adam + learning-rate schedule on the flat, bucket-ordered fp32 buffers of
// runtime/dp.py -- the other half of the reference's training step:
//   pytorch_lightning `gradient_clip_val=1.0`  (main.py:104-110)  == torch.nn.utils.clip_grad_norm_(max_norm, 2)
//   torch.optim.Adam(betas, eps, weight_decay) (lightning/optimizer.py:5-16)
//   LambdaLR(sqrt_schedule | const_schedule)   (lightning/scheduler.py:5-62)
// Three stream-ordered launches, all CUDA-graph friendly: the step counter lives on the device, so the
// learning rate, the bias corrections and the clip coefficient are computed inside the kernels.
//   1. fs2_sumsq_f32      : sum g^2 over the whole flat gradient (one read of 138 MB)
//   2. fs2_adam_step_f32  : p, m, v updated in place (4 reads + 3 writes of the flat buffers)
//   3. fs2_optim_advance  : step += 1, and re-zeroes the sum-of-squares cell
// HBM-bound: algorithmic bytes = (1 + 7) * 4 * n.
#include "common.h"
#include "util.cuh"

namespace fs2 {

struct AdamArgs {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  const float* gnorm_sq;     // [1] sum of squares of g (device)
  const long long* step;     // [1] number of optimizer steps taken so far (device)
  float lr0, beta1, beta2, eps, weight_decay, max_norm;
  int sched_type;            // 0: constant lr0, 1: sqrt_schedule, 2: const_schedule (warm-up then flat)
  int warmup;
  int n_anneal;
  int anneal_steps[8];
  float anneal_rate;
};

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = g4[i];
    s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float x = g[(n4 << 2) + threadIdx.x];
    s += x * x;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

// learning-rate factor of lightning/scheduler.py:21-62 for the optimizer step with 0-based index `step`
__device__ __forceinline__ float lr_factor(const AdamArgs& a, long long step) {
  if (a.sched_type == 0) return 1.f;
  const double cur = (double)(step + 1);
  double f = 1.0;
  if (a.warmup > 0) {
    if (cur <= (double)a.warmup) f = cur / (double)a.warmup;
    else if (a.sched_type == 1) f = sqrt((double)a.warmup / cur);
  }
  for (int i = 0; i < a.n_anneal; ++i)
    if (cur > (double)a.anneal_steps[i]) f *= (double)a.anneal_rate;
  return (float)f;
}

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  pdl_sync();
  const long long step = *a.step;  // steps already taken; this is step number step + 1
  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  float clip = 1.f;
  if (a.max_norm > 0.f) {
    const float total = sqrtf(*a.gnorm_sq);
    clip = fminf(a.max_norm / (total + 1e-6f), 1.f);
  }
  const float lr = a.lr0 * lr_factor(a, step);
  const double t = (double)(step + 1);
  const float bc1 = (float)(1.0 - pow((double)a.beta1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, t));
  const float step_size = lr / bc1;
  const long long n4 = a.n >> 2;
  float4* p4 = reinterpret_cast<float4*>(a.p);
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    float* pp = &p.x;
    float* gg = &g.x;
    float* mm = &m.x;
    float* vv = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = gg[j] * clip;
      if (a.weight_decay != 0.f) gj = fmaf(a.weight_decay, pp[j], gj);
      mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gj;
      vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gj * gj;
      const float denom = sqrtf(vv[j]) / bc2_sqrt + a.eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
  }
}

__global__ void optim_advance_kernel(long long* step, float* gnorm_sq, float* gnorm_out) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *step += 1;
    if (gnorm_out) *gnorm_out = sqrtf(*gnorm_sq);  // for logging: the un-clipped global gradient norm
    *gnorm_sq = 0.f;
  }
}

}  // namespace fs2

namespace fs2 {
// y += alpha * x on flat fp32 buffers: the inner-loop SGD step of the first-order few-shot adaptation
// (theta' = theta - lr * grad, runtime/fomaml.py) and the restore-free outer step bookkeeping.
__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha,
                                                   long long n) {
  pdl_sync();
  const long long n4 = n >> 2;
  float4* y4 = reinterpret_cast<float4*>(y);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = y4[i];
    const float4 b = x4[i];
    a.x = fmaf(alpha, b.x, a.x);
    a.y = fmaf(alpha, b.y, a.y);
    a.z = fmaf(alpha, b.z, a.z);
    a.w = fmaf(alpha, b.w, a.w);
    y4[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    y[i] = fmaf(alpha, x[i], y[i]);
  }
}
}  // namespace fs2

extern "C" {

int fs2_sumsq_f32(const float* g, int64_t n, float* out, void* stream) {
  if (n <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(g) & 15) return fs2::set_error("sumsq: buffer must be 16-byte aligned");
  FS2_LAUNCH((fs2::sumsq_kernel), 148 * 8, 256, 0, static_cast<cudaStream_t>(stream), g, n, out);
  fs2::count_launch();
  return fs2::check_launch("sumsq_kernel");
}

int fs2_adam_step_f32(float* p, const float* g, float* m, float* v, int64_t n, const float* gnorm_sq,
                      const int64_t* step, float lr0, float beta1, float beta2, float eps, float weight_decay,
                      float max_norm, int sched_type, int warmup, const int32_t* anneal_steps, int n_anneal,
                      float anneal_rate, void* stream) {
  if (n <= 0) return 0;
  if (n & 3) return fs2::set_error("adam: flat buffers are padded to multiples of 4 floats");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return fs2::set_error("adam: buffers must be 16-byte aligned");
  if (n_anneal < 0 || n_anneal > 8) return fs2::set_error("adam: at most 8 anneal steps");
  if (max_norm > 0.f && !gnorm_sq) return fs2::set_error("adam: clipping needs the gradient sum of squares");
  fs2::AdamArgs a{};
  a.p = p; a.g = g; a.m = m; a.v = v; a.n = n;
  a.gnorm_sq = gnorm_sq;
  a.step = reinterpret_cast<const long long*>(step);
  a.lr0 = lr0; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
  a.sched_type = sched_type; a.warmup = warmup; a.n_anneal = n_anneal; a.anneal_rate = anneal_rate;
  for (int i = 0; i < n_anneal; ++i) a.anneal_steps[i] = anneal_steps[i];
  FS2_LAUNCH((fs2::adam_step_kernel), 148 * 8, 256, 0, static_cast<cudaStream_t>(stream), a);
  fs2::count_launch();
  return fs2::check_launch("adam_step_kernel");
}

int fs2_optim_advance(int64_t* step, float* gnorm_sq, float* gnorm_out, void* stream) {
  FS2_LAUNCH((fs2::optim_advance_kernel), 1, 32, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<long long*>(step),
                                                                            gnorm_sq, gnorm_out);
  fs2::count_launch();
  return fs2::check_launch("optim_advance_kernel");
}

// y[i] += alpha * x[i], f32 [n] device, 16-byte aligned.
int fs2_axpy_f32(float* y, const float* x, float alpha, int64_t n, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(x) & 15))
    return fs2::set_error("axpy: buffers must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  FS2_LAUNCH((fs2::axpy_kernel), (unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream), y, x, alpha,
             (long long)n);
  fs2::count_launch();
  return fs2::check_launch("axpy_kernel");
}
}
