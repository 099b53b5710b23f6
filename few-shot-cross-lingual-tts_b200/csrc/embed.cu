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

// Small tables (the 256-bin pitch / energy embeddings: 12800 rows hit 256 table rows, ~50 writers per address):
// block (x, y) owns 64 channels and a contiguous range of input rows and accumulates into its own shared-memory copy
// of that table slice; only the non-zero sums go to the gradient with global atomics (one per touched table element
// and block instead of one per input element: 55-63 us -> a few us at C2).
constexpr int kEmbSliceC = 64;
template <typename IdxT>
__global__ void __launch_bounds__(256)
embedding_bwd_smem_kernel(const __nv_bfloat16* __restrict__ dy, const IdxT* __restrict__ ids, long long rows, int C,
                          int n_rows_table, int pad_idx, int rows_per_block, float* __restrict__ dtable) {
  pdl_sync();
  extern __shared__ float acc[];  // [n_rows_table][64]
  const int n = n_rows_table * kEmbSliceC;
  for (int i = threadIdx.x; i < n; i += 256) acc[i] = 0.f;
  __syncthreads();
  const int c0 = blockIdx.x * kEmbSliceC;
  const int sub = threadIdx.x & 7, rl = threadIdx.x >> 3;  // 8 threads x 8 channels per row, 32 rows per pass
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, rows);
  if (c0 + sub * 8 < C) {
    for (long long r = r0 + rl; r < r1; r += 32) {
      const long long id = static_cast<long long>(ids[r]);
      if (id < 0 || id >= n_rows_table || id == pad_idx) continue;  // padding_idx rows get no gradient
      float f[8];
      unpack8(ld8(dy + r * C + c0 + sub * 8), f);
      float* a = acc + id * kEmbSliceC + sub * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(a + j, f[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = acc[i];
    const int row = i / kEmbSliceC, c = c0 + (i % kEmbSliceC);
    if (v != 0.f && c < C) atomicAdd(dtable + (long long)row * C + c, v);
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
  const size_t smem = (size_t)n_rows_table * fs2::kEmbSliceC * sizeof(float);
  if (smem <= 200 * 1024 && rows >= 8 * (int64_t)n_rows_table) {  // many writers per table row: shared-memory slices
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(fs2::embedding_bwd_smem_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(fs2::embedding_bwd_smem_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    int chunks = (int)(rows / 1024);
    chunks = chunks < 1 ? 1 : (chunks > 16 ? 16 : chunks);
    const int rpb = (int)((rows + chunks - 1) / chunks);
    const dim3 grid((C + fs2::kEmbSliceC - 1) / fs2::kEmbSliceC, (unsigned)((rows + rpb - 1) / rpb));
    if (ids_is_i64)
      FS2_LAUNCH((fs2::embedding_bwd_smem_kernel<int64_t>), grid, 256, smem, s, static_cast<const __nv_bfloat16*>(dy),
                 static_cast<const int64_t*>(ids), rows, C, n_rows_table, pad_idx, rpb, dtable);
    else
      FS2_LAUNCH((fs2::embedding_bwd_smem_kernel<int32_t>), grid, 256, smem, s, static_cast<const __nv_bfloat16*>(dy),
                 static_cast<const int32_t*>(ids), rows, C, n_rows_table, pad_idx, rpb, dtable);
    fs2::count_launch();
    return fs2::check_launch("embedding_bwd_smem_kernel");
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
