/*
 * fs2b200.h -- C ABI of libfs2b200.so: hand-written sm_100a kernels for the FastSpeech2
 * acoustic-model training step (forward + losses + backward).
 *
 * The reference (hhhaaahhhaa/Few-Shot-Cross-Lingual-TTS) is pure Python and has NO FFI of its own
 * (SURVEY.md section 8b); every entry point below therefore cites the reference *Python* op it
 * replaces.  The Python boundary classes (same names / signatures as the reference's
 * transformer.* and lightning.model.* modules) call these through ctypes.
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer owned by the caller (torch tensors); the library
 *     never allocates, frees or synchronises; every call is ordered on `stream` (a cudaStream_t);
 *   - activations are channels-last row-major [B][T][C]; `bf16` = __nv_bfloat16, `f32` = float,
 *     lengths / durations are int64 as in the reference's batch tuple (collates/utils.py:70-85);
 *   - return value 0 = OK, non-zero = error; fs2_last_error() gives the thread-local message;
 *   - no global mutable state (the only cache is the driver entry point of cuTensorMapEncodeTiled).
 */
#ifndef FS2B200_H_
#define FS2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS2_ABI_VERSION 1

int fs2_version(void);
const char* fs2_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t fs2_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* tcgen05 / TMEM / TMA GEMM family                                                            */
/* ------------------------------------------------------------------------------------------ */

/* One bf16 GEMM operand viewed as a 3-D tensor [batches][rows][inner] (inner contiguous). */
typedef struct fs2_operand {
  const void* ptr;      /* bf16 device pointer (16-byte aligned) */
  int64_t ld;           /* elements between rows (multiple of 8) */
  int64_t batch_stride; /* elements between batches (multiple of 8) */
  int32_t inner;        /* extent of the contiguous dimension */
  int32_t rows;         /* extent of the row dimension */
  int32_t batches;      /* extent of the batch dimension */
  int32_t mn_major;     /* 0: contiguous dim is the reduction (K-major)
                           1: contiguous dim is M (operand A) / N (operand B) */
  int32_t inner_base;   /* first contiguous-dim coordinate used */
  int32_t zdiv;         /* batch coordinate = z / zdiv (>= 1)  */
  int32_t zmod_stride;  /* contiguous coordinate += (z % zdiv) * zmod_stride (head select) */
} fs2_operand;

enum { FS2_GEMM_NORMAL = 0, FS2_GEMM_WGRAD = 1 };
enum { FS2_EPI_NONE = 0, FS2_EPI_RELU = 1, FS2_EPI_RELU_BWD = 2, FS2_EPI_ADD_AUX = 3 };

/*
 * NORMAL:  D[z][m][n] = epi( alpha * sum_{tap<taps} sum_{k<K} A_z[m + tap_shift0 + tap][k]
 *                                              * B_z[n][tap*b_tap_kstride + k]  + bias[n] )
 *          Out-of-range rows of A read as zero (per-sequence zero halo of Conv1d).
 *          Replaces nn.Linear (transformer/SubLayers.py:18-25,39-41,54), nn.Conv1d on
 *          channels-last data (SubLayers.py:68-80,87-89; lightning/model/modules.py:211-242;
 *          transformer/Layers.py:33-64), torch.bmm (transformer/Modules.py:16,23) and their
 *          autograd input-gradients.
 * WGRAD:   D[m][tap*N + n] (+)= sum_{zb<a.batches} sum_{r<a.rows} A_zb[r][m] * B_zb[r+tap_shift0+tap][n]
 *          (both operands MN-major; split-K over `splits` CTAs groups with fp32 atomics).
 *          Replaces autograd's weight gradients of the same Linear / Conv1d modules.
 */
typedef struct fs2_gemm {
  fs2_operand a, b;
  int32_t mode;
  int32_t M, N, K;
  int32_t Z;
  int32_t taps;
  int32_t tap_shift0;
  int32_t b_tap_kstride;
  int32_t splits;
  int32_t epilogue;
  int32_t d_f32;       /* 0: D is bf16, 1: D is f32 */
  int32_t d_atomic;    /* 1: accumulate into D with atomics (f32 only) */
  int32_t d_zdiv;      /* D offset = (z/d_zdiv)*d_zdiv_stride + (z%d_zdiv)*d_zmod_stride */
  float alpha;
  void* d;
  int64_t ldd;
  int64_t d_col_stride; /* elements between consecutive n (1 except WGRAD into [Cout][Cin][k]) */
  int64_t d_tap_stride; /* WGRAD: elements between taps */
  int64_t d_zdiv_stride;
  int64_t d_zmod_stride;
  const float* bias;    /* [N] f32 or NULL */
  const void* aux;      /* bf16 [Z][M][ld_aux] or NULL */
  int64_t ld_aux;
  int64_t aux_batch_stride;
} fs2_gemm;

/* impl: 0 = tcgen05 (product path), 1 = plain CUDA-core kernel (debug cross-check only). */
int fs2_gemm_bf16(const fs2_gemm* g, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2B200_H_ */
