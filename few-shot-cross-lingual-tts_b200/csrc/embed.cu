// Embedding-shaped ops of the variance adaptor and the phoneme front-end.
//
//  * fs2_bucket_embed_add_bf16: y = x + table[bucketize(target, bins)]
//      replaces lightning/model/modules.py:82-102,119-128 (`torch.bucketize(target, bins)` with the
//      default right=False -> index = #{bins < v}, `nn.Embedding(n_bins, d)`, `x = x + embedding`).
//      Targets may be f32 or f64 (energies can arrive as f64, collates/utils.py:82): the comparison is
//      done in double in both cases, which equals torch's type-promoted comparison.
//  * fs2_embedding_fwd_bf16 / fs2_embedding_bwd_f32: row gather and scatter-add
//      replaces F.embedding (lightning/systems/language/embeddings.py:25-31) and the backward of both
//      embeddings above (fp32 atomics into the [rows][C] table gradient; order is non-deterministic
//      at the 1e-7 level, like torch's own embedding_dense_backward on CUDA).
#include "common.h"
#include "util.cuh"

namespace fs2 {

template <typename TgtT>
__global__ void __launch_bounds__(256)
bucket_embed_add_kernel(const __nv_bfloat16* __restrict__ x, const TgtT* __restrict__ target,
                        const float* __restrict__ bins, int nb, const float* __restrict__ table,
                        long long rows, int C, __nv_bfloat16* __restrict__ y,
                        int32_t* __restrict__ idx_out) {
  pdl_sync();
  extern __shared__ float s_bins[];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = bins[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double v = static_cast<double>(target[row]);
  int lo = 0, hi = nb;  // first i with bins[i] >= v  (== count of bins < v)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (static_cast<double>(s_bins[mid]) < v) lo = mid + 1; else hi = mid;
  }
  if (lane == 0) idx_out[row] = lo;
  const float* e = table + (long long)lo * C;
  for (int c = lane * 8; c < C; c += 256) {
    float f[8];
    unpack8(ld8(x + row * C + c), f);
    const float4 e0 = *reinterpret_cast<const float4*>(e + c);
    const float4 e1 = *reinterpret_cast<const float4*>(e + c + 4);
    f[0] += e0.x; f[1] += e0.y; f[2] += e0.z; f[3] += e0.w;
    f[4] += e1.x; f[5] += e1.y; f[6] += e1.z; f[7] += e1.w;
    st8(y + row * C + c, pack8(f));
  }
}

__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table, long long rows,
                     int C, int n_rows_table, int pad_idx, __nv_bfloat16* __restrict__ y) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = ids[row];
  if (id < 0 || id >= n_rows_table) id = pad_idx >= 0 ? pad_idx : 0;
  const float* e = table + id * C;
  for (int c = lane * 8; c < C; c += 256) {
    const float4 e0 = *reinterpret_cast<const float4*>(e + c);
    const float4 e1 = *reinterpret_cast<const float4*>(e + c + 4);
    const float f[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    st8(y + row * C + c, pack8(f));
  }
}

// Each warp walks a contiguous run of rows and keeps a running sum while the table index stays the
// same (padded positions all bucketise to one bin and are contiguous), flushing with atomics only
// when the index changes: removes the same-address atomic storm of the naive scatter-add.
constexpr int kEmbRowsPerWarp = 16;

template <typename IdxT>
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const IdxT* __restrict__ ids,
                     long long rows, int C, int n_rows_table, int pad_idx,
                     float* __restrict__ dtable) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long r0 = w * kEmbRowsPerWarp;
  if (r0 >= rows) return;
  const long long r1 = min(r0 + (long long)kEmbRowsPerWarp, rows);
  for (int c = lane * 8; c < C; c += 256) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    long long cur = -1;
    for (long long r = r0; r < r1; ++r) {
      long long id = static_cast<long long>(ids[r]);
      if (id < 0 || id >= n_rows_table || id == pad_idx) id = -1;  // padding_idx rows get no gradient
      if (id != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dtable + cur * C + c + j, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        cur = id;
      }
      if (id >= 0) {
        float f[8];
        unpack8(ld8(dy + r * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    if (cur >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dtable + cur * C + c + j, acc[j]);
    }
  }
}

// Small tables (the 256-bin pitch / energy embeddings: 12800 rows hit 256 table rows, and ALL padded phonemes hit the
// bin of 0.0 -- the scatter-add above spends 55-63 us on 3.3 M atomics, thousands of them on the same addresses).
// Counting sort instead: block (x, y, z) owns 64 channels, 32 table rows ("bins") and one slice of the input rows.
// It histograms the ids of its row slice that fall into its bins, turns the counts into segment starts and scatters
// (bin, row) pairs into a shared-memory list ordered by bin (integer shared-memory atomics only); the 32 warps then
// take EQUAL shares of that list (a bin with thousands of rows is spread over many warps), add up 128-byte row slices
// with eight loads in flight and flush a partial sum whenever the bin changes.  One global atomic per touched table
// element and block at the end (instead of one per input element).
constexpr int kEmbSliceC = 64, kEmbBinsPerBlock = 32;
template <typename IdxT>
__global__ void __launch_bounds__(1024)
embedding_bwd_sorted_kernel(const __nv_bfloat16* __restrict__ dy, const IdxT* __restrict__ ids, int rows, int C,
                            int n_rows_table, int pad_idx, int rows_per_z, float* __restrict__ dtable) {
  pdl_sync();
  extern __shared__ int s_order[];  // [rows_per_z] (bin << 20 | row - r0), grouped by bin
  __shared__ int s_cnt[kEmbBinsPerBlock], s_cur[kEmbBinsPerBlock], s_total;
  __shared__ float s_acc[kEmbBinsPerBlock][kEmbSliceC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * kEmbSliceC;
  const int bin0 = blockIdx.y * kEmbBinsPerBlock;
  const int r0 = blockIdx.z * rows_per_z, r1 = min(r0 + rows_per_z, rows);
  if (threadIdx.x < kEmbBinsPerBlock) s_cnt[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < kEmbBinsPerBlock * kEmbSliceC; i += 1024) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  auto local_bin = [&](int r) {
    const long long id = static_cast<long long>(ids[r]);
    if (id < 0 || id >= n_rows_table || id == pad_idx) return -1;  // padding_idx rows get no gradient
    const long long lb = id - bin0;
    return (lb >= 0 && lb < kEmbBinsPerBlock) ? (int)lb : -1;
  };
  for (int r = r0 + threadIdx.x; r < r1; r += 1024) {
    const int lb = local_bin(r);
    if (lb >= 0) atomicAdd(&s_cnt[lb], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 32 counts
    const int v = s_cnt[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    s_cur[lane] = inc - v;
    if (lane == 31) s_total = inc;
  }
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += 1024) {
    const int lb = local_bin(r);
    if (lb >= 0) s_order[atomicAdd(&s_cur[lb], 1)] = (lb << 20) | (r - r0);
  }
  __syncthreads();
  const int n = s_total;
  const int col = c0 + 2 * lane;
  if (n > 0 && col < C) {
    const int per = (n + 31) / 32;
    int k = warp * per;
    const int kend = min(k + per, n);
    float ax = 0.f, ay = 0.f;
    int cur = k < kend ? (s_order[k] >> 20) : 0;
    auto flush = [&](int bin) {
      atomicAdd(&s_acc[bin][2 * lane], ax);
      atomicAdd(&s_acc[bin][2 * lane + 1], ay);
      ax = ay = 0.f;
    };
    const __nv_bfloat16* base = dy + (long long)r0 * C + col;
    for (; k + 8 <= kend; k += 8) {
      int e[8];
      __nv_bfloat162 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        e[u] = s_order[k + u];
        v[u] = *reinterpret_cast<const __nv_bfloat162*>(base + (long long)(e[u] & 0xFFFFF) * C);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int bin = e[u] >> 20;
        if (bin != cur) {
          flush(cur);
          cur = bin;
        }
        const float2 f = __bfloat1622float2(v[u]);
        ax += f.x;
        ay += f.y;
      }
    }
    for (; k < kend; ++k) {
      const int e = s_order[k], bin = e >> 20;
      if (bin != cur) {
        flush(cur);
        cur = bin;
      }
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + (long long)(e & 0xFFFFF) * C));
      ax += f.x;
      ay += f.y;
    }
    if (warp * per < kend) flush(cur);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kEmbBinsPerBlock * kEmbSliceC; i += 1024) {
    const int bin = i / kEmbSliceC, c = c0 + (i % kEmbSliceC);
    const float v = s_acc[bin][i % kEmbSliceC];
    if (v != 0.f && c < C && bin0 + bin < n_rows_table) atomicAdd(dtable + (long long)(bin0 + bin) * C + c, v);
  }
}

// ---- lookup in the CONCATENATION of several tables without materialising it -----------------------------------
// MultilingualEmbedding.forward (lightning/systems/language/embeddings.py:25-31) does torch.cat(all tables) on every
// call and F.embedding(padding_idx) on the result; here the per-language tables stay where they are and a row id is
// resolved against the cumulative row counts (<= kMaxTables tables, passed by value in the kernel parameters).
constexpr int kMaxTables = 32;
struct EmbTables {
  const float* ptr[kMaxTables];  // forward: table values; backward: table gradients
  int row0[kMaxTables + 1];      // first concatenated row of table k; row0[n] = total rows
  int n;
};
__device__ __forceinline__ int find_table(const EmbTables& t, long long id) {
  int k = 0;
  while (k + 1 < t.n && id >= t.row0[k + 1]) ++k;
  return k;
}

__global__ void __launch_bounds__(256)
embedding_multi_fwd_kernel(const int64_t* __restrict__ ids, const __grid_constant__ EmbTables tabs, long long rows,
                           int C, int pad_idx, __nv_bfloat16* __restrict__ y) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = ids[row];
  if (id < 0 || id >= tabs.row0[tabs.n]) id = pad_idx >= 0 ? pad_idx : 0;
  const int k = find_table(tabs, id);
  const float* e = tabs.ptr[k] + (id - tabs.row0[k]) * C;
  for (int c = lane * 8; c < C; c += 256) {
    const float4 e0 = *reinterpret_cast<const float4*>(e + c);
    const float4 e1 = *reinterpret_cast<const float4*>(e + c + 4);
    const float f[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    st8(y + row * C + c, pack8(f));
  }
}

__global__ void __launch_bounds__(256)
embedding_multi_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const int64_t* __restrict__ ids,
                           const __grid_constant__ EmbTables tabs, long long rows, int C, int pad_idx) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long r0 = w * kEmbRowsPerWarp;
  if (r0 >= rows) return;
  const long long r1 = min(r0 + (long long)kEmbRowsPerWarp, rows);
  for (int c = lane * 8; c < C; c += 256) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    long long cur = -1;
    auto flush = [&]() {
      if (cur < 0) return;
      const int k = find_table(tabs, cur);
      float* d = const_cast<float*>(tabs.ptr[k]) + (cur - tabs.row0[k]) * C + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(d + j, acc[j]);
    };
    for (long long r = r0; r < r1; ++r) {
      long long id = ids[r];
      if (id < 0 || id >= tabs.row0[tabs.n] || id == pad_idx) id = -1;  // padding_idx row gets no gradient
      if (id != cur) {
        flush();
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        cur = id;
      }
      if (id >= 0) {
        float f[8];
        unpack8(ld8(dy + r * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    flush();
  }
}

static int fill_tables(EmbTables& t, const void* const* tables, const int32_t* table_rows, int n_tables) {
  if (n_tables < 1 || n_tables > kMaxTables) return set_error("embedding_multi: 1..32 tables");
  t.n = n_tables;
  int r = 0;
  for (int k = 0; k < n_tables; ++k) {
    if (!tables[k] || table_rows[k] <= 0) return set_error("embedding_multi: empty table");
    t.ptr[k] = static_cast<const float*>(tables[k]);
    t.row0[k] = r;
    r += table_rows[k];
  }
  t.row0[n_tables] = r;
  return 0;
}

}  // namespace fs2

extern "C" {

// tables / table_rows: HOST arrays (n_tables device pointers to f32 [rows_k][C] tables, and their row counts); ids
// index the concatenation of the tables in array order.  pad_idx: concatenated row that gets no gradient
// (F.embedding padding_idx), -1 = none.
int fs2_embedding_multi_fwd_bf16(const int64_t* ids, const void* const* tables, const int32_t* table_rows,
                                 int n_tables, int64_t rows, int C, int pad_idx, void* y, void* stream) {
  if (C % 8) return fs2::set_error("embedding_multi_fwd: C must be a multiple of 8");
  if (rows <= 0) return 0;
  fs2::EmbTables t;
  if (int rc = fs2::fill_tables(t, tables, table_rows, n_tables)) return rc;
  FS2_LAUNCH((fs2::embedding_multi_fwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream),
             ids, t, rows, C, pad_idx, static_cast<__nv_bfloat16*>(y));
  fs2::count_launch();
  return fs2::check_launch("embedding_multi_fwd_kernel");
}

// dtables: HOST array of device pointers to the f32 [rows_k][C] table GRADIENTS (accumulated with atomics).
int fs2_embedding_multi_bwd_f32(const void* dy, const int64_t* ids, const void* const* dtables,
                                const int32_t* table_rows, int n_tables, int64_t rows, int C, int pad_idx,
                                void* stream) {
  if (C % 8) return fs2::set_error("embedding_multi_bwd: C must be a multiple of 8");
  if (rows <= 0) return 0;
  fs2::EmbTables t;
  if (int rc = fs2::fill_tables(t, dtables, table_rows, n_tables)) return rc;
  const unsigned grid = (unsigned)((rows + 8 * fs2::kEmbRowsPerWarp - 1) / (8 * fs2::kEmbRowsPerWarp));
  FS2_LAUNCH((fs2::embedding_multi_bwd_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(dy), ids, t, rows, C, pad_idx);
  fs2::count_launch();
  return fs2::check_launch("embedding_multi_bwd_kernel");
}


int fs2_bucket_embed_add_bf16(const void* x, const void* target, int target_is_f64, const float* bins,
                              int n_bins_minus_1, const float* table, int64_t rows, int C, void* y,
                              int32_t* idx_out, void* stream) {
  if (C % 8) return fs2::set_error("bucket_embed_add: C must be a multiple of 8");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const size_t smem = (size_t)n_bins_minus_1 * sizeof(float);
  if (target_is_f64)
    FS2_LAUNCH((fs2::bucket_embed_add_kernel<double>), grid, 256, smem, s, 
        static_cast<const __nv_bfloat16*>(x), static_cast<const double*>(target), bins,
        n_bins_minus_1, table, rows, C, static_cast<__nv_bfloat16*>(y), idx_out);
  else
    FS2_LAUNCH((fs2::bucket_embed_add_kernel<float>), grid, 256, smem, s, 
        static_cast<const __nv_bfloat16*>(x), static_cast<const float*>(target), bins, n_bins_minus_1,
        table, rows, C, static_cast<__nv_bfloat16*>(y), idx_out);
  fs2::count_launch();
  return fs2::check_launch("bucket_embed_add_kernel");
}

int fs2_embedding_fwd_bf16(const int64_t* ids, const float* table, int64_t rows, int C,
                           int n_rows_table, int pad_idx, void* y, void* stream) {
  if (C % 8) return fs2::set_error("embedding_fwd: C must be a multiple of 8");
  if (rows <= 0) return 0;
  FS2_LAUNCH((fs2::embedding_fwd_kernel), (unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      ids, table, rows, C, n_rows_table, pad_idx, static_cast<__nv_bfloat16*>(y));
  fs2::count_launch();
  return fs2::check_launch("embedding_fwd_kernel");
}

// ids: int32 (ids_is_i64 = 0, the bucket indices saved by the forward) or int64.
int fs2_embedding_bwd_f32(const void* dy, const void* ids, int ids_is_i64, int64_t rows, int C,
                          int n_rows_table, int pad_idx, float* dtable, void* stream) {
  if (C % 8) return fs2::set_error("embedding_bwd: C must be a multiple of 8");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int zs = (int)(rows / 2048);
  zs = zs < 1 ? 1 : (zs > 8 ? 8 : zs);
  const int rows_per_z = (int)((rows + zs - 1) / zs);
  const size_t smem = (size_t)rows_per_z * sizeof(int);
  if (smem <= 160 * 1024 && rows_per_z < (1 << 20) && rows >= 8 * (int64_t)n_rows_table && !(C & 1) && rows < (1ll << 31)) {
    // many writers per table row: counting sort, see embedding_bwd_sorted_kernel
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(fs2::embedding_bwd_sorted_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cudaFuncSetAttribute(fs2::embedding_bwd_sorted_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_set = true;
    }
    const dim3 grid((C + fs2::kEmbSliceC - 1) / fs2::kEmbSliceC,
                    (n_rows_table + fs2::kEmbBinsPerBlock - 1) / fs2::kEmbBinsPerBlock, (unsigned)((rows + rows_per_z - 1) / rows_per_z));
    if (ids_is_i64)
      FS2_LAUNCH((fs2::embedding_bwd_sorted_kernel<int64_t>), grid, 1024, smem, s, static_cast<const __nv_bfloat16*>(dy),
                 static_cast<const int64_t*>(ids), (int)rows, C, n_rows_table, pad_idx, rows_per_z, dtable);
    else
      FS2_LAUNCH((fs2::embedding_bwd_sorted_kernel<int32_t>), grid, 1024, smem, s, static_cast<const __nv_bfloat16*>(dy),
                 static_cast<const int32_t*>(ids), (int)rows, C, n_rows_table, pad_idx, rows_per_z, dtable);
    fs2::count_launch();
    return fs2::check_launch("embedding_bwd_sorted_kernel");
  }
  const unsigned grid = (unsigned)((rows + 8 * fs2::kEmbRowsPerWarp - 1) / (8 * fs2::kEmbRowsPerWarp));
  if (ids_is_i64)
    FS2_LAUNCH((fs2::embedding_bwd_kernel<int64_t>), grid, 256, 0, s, static_cast<const __nv_bfloat16*>(dy),
                                                            static_cast<const int64_t*>(ids), rows, C,
                                                            n_rows_table, pad_idx, dtable);
  else
    FS2_LAUNCH((fs2::embedding_bwd_kernel<int32_t>), grid, 256, 0, s, static_cast<const __nv_bfloat16*>(dy),
                                                            static_cast<const int32_t*>(ids), rows, C,
                                                            n_rows_table, pad_idx, dtable);
  fs2::count_launch();
  return fs2::check_launch("embedding_bwd_kernel");
}
}
