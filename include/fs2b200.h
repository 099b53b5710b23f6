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

#define FS2_ABI_VERSION 2

int fs2_version(void);
const char* fs2_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t fs2_launch_count(void);
/* name of the kernel this thread launched last (which engine a GEMM descriptor was routed to: tools / bench) */
const char* fs2_last_kernel(void);
/* id of the CUDA-graph capture `stream` is in (cudaStreamGetCaptureInfo), 0 when it is not capturing.  Callers that
   hand out self-resetting scratch (fs2_gemm::workspace, the fs2_loss_fwd workspace) use one buffer per capture and
   zero its counters INSIDE the captured graph, so that no graph depends on another one having run first. */
int64_t fs2_stream_capture_id(void* stream);

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
  /* optional row segmentation of D (fused Q|K|V weight gradients land in three separate tensors):
     row m goes to d_seg[m / d_seg_rows] + (m % d_seg_rows) * ldd; d_seg_rows = 0 disables it. */
  int32_t d_seg_rows;
  int32_t d_seg_pad;
  void* d_seg[4];
  /* optional ragged rows (key/frame padding: transformer/Layers.py:25,28 zero those rows anyway, so their
     GEMM work is skipped instead of computed and thrown away).  row_lens: int64 [.] valid-row counts;
     NORMAL: rows m >= row_lens[z / lens_zdiv] of D[z] are written as ZERO and output tiles made only of
             such rows skip their TMA loads and MMAs (compact tile schedule built on the device);
     WGRAD:  the caller PROMISES that rows r >= row_lens[zb / lens_zdiv] of A_zb are zero; 64-row reduction
             blocks made only of such rows are skipped (split-K ranges are cut over the remaining blocks). */
  const int64_t* row_lens;
  int32_t lens_zdiv;
  /* NORMAL + row_lens: what happens to the rows of D[z] behind the last scheduled row tile (scheduling unit:
     128 rows, or 256 when the 2-CTA kernel runs).  0: all of them are written as zero (every element of D is
     defined); n > 0: only the first n are zeroed (enough for a consumer that reads an n-row halo, e.g. the
     input-gradient of a Conv1d with kernel 2n+1), the rest of D is left untouched; n < 0: none is written
     (consumers that skip padded rows themselves: LayerNorm, attention, row_lens GEMMs with k = 1). */
  int32_t tail_zero_rows;
  /* optional 1-bit ReLU mask (NORMAL mode, bf16 D, N % 64 == 0): uint64 [Z*M rows][N/64], bit j of word w of a row =
     column 64*w + j.  FS2_EPI_RELU writes it (output > 0); FS2_EPI_RELU_BWD reads it INSTEAD of the bf16 `aux`
     tensor -- the backward of Conv1d -> ReLU -> Conv1d (transformer/SubLayers.py:87-89) then re-reads 1/16 of the
     bytes of the hidden activation.  NULL = off. */
  void* relu_mask;
  /* optional scratch of the Conv1d kernel (NORMAL mode, taps > 1, bf16 D): with it, the tiles of the LAST, partially
     filled wave of the persistent schedule are split over the reduction (channel blocks) across the idle CTA pairs;
     partial accumulators meet in this buffer and are added in a fixed order (bit-reproducible).  Its first 4096
     bytes (arrival counters) must be ZERO on first use (the kernel leaves them zero again); at least
     fs2_gemm_workspace_bytes() bytes; not shared by launches that can run concurrently (one buffer per stream).
     NULL = no split. */
  void* workspace;
  int64_t workspace_bytes;
  /* optional fused LayerNorm epilogue (NORMAL mode, taps = 1, N == 256, bf16 D; transformer/SubLayers.py:88-91
     `layer_norm(dropout(w_2(h)) + residual)` + transformer/Layers.py:28 masked_fill in the GEMM that computes w_2):
       f = bf16(acc + bias); v = dropout(f) / (1 - p) + ln_res; D = (v - mean) * rstd * ln_gamma + ln_beta,
     rows m >= row_lens[z] of D are zero.  ln_gamma != NULL enables it.  Also written, all indexed by the dense row
     z * M + m: ln_v (bf16 [Z*M][256], the pre-norm sum the backward needs), ln_mean / ln_rstd (f32 [Z*M]) and
     ln_keep (uint8 [Z*M][32] dropout keep bits, same stream and layout as fs2_ln_fwd_bf16 with the same seed). */
  const float* ln_gamma;
  const float* ln_beta;
  const void* ln_res;          /* bf16 [Z][M][ld_res] */
  int64_t ld_res;
  int64_t res_batch_stride;
  float ln_p_drop;
  int32_t ln_pad0;
  uint64_t ln_seed;
  const uint64_t* ln_seed_dev;
  void* ln_v;
  float* ln_mean;
  float* ln_rstd;
  uint8_t* ln_keep;
  /* optional, WGRAD mode: a_colsum[m] += sum over z and over the reduction rows r of A[z][r][m] (f32 [M]) -- the
     bias gradient of the Conv1d / Linear whose weight gradient this call computes (autograd of `nn.Conv1d.bias`).
     The tap-group weight-gradient kernel (taps >= 2, N % 128 == 0, M >= 256) sums the dY tiles it has in shared
     memory anyway; any other kernel choice runs a separate column-sum launch on the same stream first (A must then
     be densely batched: batch_stride == rows * ld, lens_zdiv == 1, M % 8 == 0).  NULL = off.
     With a segmented output (d_seg_rows > 0) the sums of row block i go to a_colsum_seg[i] instead (NULL entries
     are skipped; a_colsum must then be NULL): the Q / K / V bias gradients next to the fused QKV weight gradient.
     The 2-CTA weight-gradient kernel reads its dY tiles the same way as the tap-group kernel. */
  float* a_colsum;
  float* a_colsum_seg[4];
} fs2_gemm;

/* impl: 0 = tcgen05 (product path), 1 = plain CUDA-core kernel (debug cross-check only). */
int fs2_gemm_bf16(const fs2_gemm* g, int impl, void* stream);
/* bytes of fs2_gemm::workspace that serve every launch on this device */
int64_t fs2_gemm_workspace_bytes(void);


/* ------------------------------------------------------------------------------------------ */
/* Fused (dropout +) residual + LayerNorm (+ dropout) + padding-row zeroing                    */
/*   transformer/SubLayers.py:54-55,90-91; transformer/Layers.py:25,28;                        */
/*   lightning/model/modules.py:222-225,234-237                                                */
/*   x,res,y,dy,dx,dres: bf16 [B][T][C], C in {256,512,1024}; mean,rstd: f32 [B*T];           */
/*   lens: int64 [B] or NULL (rows t >= lens[b] are zeroed); drop_mode 1: LN(drop(x)+res),      */
/*   2: drop(LN(x+res)); the mask comes from a Philox4x32-7 stream keyed by seed_dev[0]        */
/*   (device step counter, may be NULL) mixed with the call-site salt `seed`; p is resolved to */
/*   1/8192.  keep_out / keep_in: uint8 [B*T][C/8] keep bits (bit j of byte v = channel 8v+j),  */
/*   written by the forward (optional there) and READ by the backward (required when p > 0).   */
/*   Backward only, drop_mode 3: x is the pre-norm sum written by the fused GEMM epilogue        */
/*   (fs2_gemm::ln_v), res must be NULL; dx = mask / (1 - p) * dpre, dres = dpre.               */
/*   dgamma/dbeta: f32 [C], accumulated with atomics (zero them first); dbias (optional, f32    */
/*   [C]): column sums of dx, i.e. the bias gradient of the GEMM / conv that produced x.        */
/*   Rows t >= lens[b] are never read: y / dx / dres are zero there.                            */
/* ------------------------------------------------------------------------------------------ */
int fs2_ln_fwd_bf16(const void* x, const void* res, const float* gamma, const float* beta,
                    const int64_t* lens, int B, int T, int C, float p_drop, int drop_mode,
                    uint64_t seed, const uint64_t* seed_dev, void* y, float* mean, float* rstd,
                    uint8_t* keep_out, void* stream);
int fs2_ln_bwd_bf16(const void* dy, const void* x, const void* res, const float* gamma,
                    const float* mean, const float* rstd, const int64_t* lens, int B, int T, int C,
                    float p_drop, int drop_mode, int relu_x, const uint8_t* keep_in,
                    void* dx, void* dres, float* dgamma, float* dbeta, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Key-padding-masked softmax (transformer/Modules.py:17-22), z = b*H + h                      */
/*   S,dP: f32 [Z][T][Tp]; P,dS: bf16 [Z][T][Tp]; lens: int64 [Z/H]; Tp % 128 == 0, <= 2048     */
/* ------------------------------------------------------------------------------------------ */
int fs2_softmax_fwd(const float* S, const int64_t* lens, int Z, int H, int T, int Tp, void* P,
                    void* stream);
int fs2_softmax_bwd(const void* P, const float* dP, const int64_t* lens, int Z, int H, int T, int Tp,
                    float alpha, void* dS, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused masked-softmax attention, d_k = 128 (transformer/Modules.py:14-25 + the head split /   */
/* merge of transformer/SubLayers.py:39-52).  qkv: bf16 [B][T][3*H*dk]; out: bf16 [B][T][H*dk]; */
/* lse2: f32 [B*H][T] log2-domain log-sum-exp of the scaled scores (+inf on padded rows).       */
/* ------------------------------------------------------------------------------------------ */
/* sched (optional, NULL = natural order): int32 [B*H*ceil(T/128)] work order written by          */
/* fs2_attn_schedule from the device-side lengths -- tiles of the longest utterances first,      */
/* padded-only tiles last (the CTAs of a ragged batch differ 8x in work).                        */
int fs2_attn_schedule(const int64_t* lens, int B, int T, int H, int32_t* sched, void* stream);
int fs2_attn_fwd_bf16(const void* qkv, const int64_t* lens, const int32_t* sched, int B, int T, int H, int dk,
                      void* out, float* lse2, void* stream);
/* backward: o, d_o: bf16 [B][T][H*dk]; dsum: f32 [B*H][T] workspace; dqkv: bf16 [B][T][3*H*dk];  */
/* dbias_q / dbias_v (optional, f32 [H*dk], ACCUMULATED): column sums of dQ / dV over all frames = */
/* the gradients of the w_qs / w_vs biases (transformer/SubLayers.py:18-20), taken from the tiles   */
/* on their way out instead of re-reading dqkv (the key bias has no gradient: softmax ignores it). */
int fs2_attn_bwd_bf16(const void* qkv, const void* o, const void* d_o, const float* lse2, const int64_t* lens,
                      const int32_t* sched, int B, int T, int H, int dk, float* dsum, void* dqkv, float* dbias_q,
                      float* dbias_v, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* LengthRegulator (lightning/model/modules.py:169-196, lightning/utils/tool.py:168-186)       */
/*   integer cumsum + vectorised gather; bit-exact copy semantics; see csrc/length_regulator.cu */
/* ------------------------------------------------------------------------------------------ */
int fs2_lr_index(const void* dur, int dur_is_f32, int B, int Ts, int max_len, int64_t* cum,
                 int32_t* idx, int64_t* mel_len, void* stream);
int fs2_lr_gather(const void* x, const int32_t* idx, int B, int Ts, int max_len, int out_len,
                  int row_bytes, void* out, void* stream);
int fs2_lr_gather_fused_bf16(const void* x, const int32_t* idx, const float* spk, const float* pe,
                             int B, int Ts, int max_len, int out_len, int C, void* out, void* stream);
int fs2_lr_bwd_bf16(const void* dout, const int64_t* cum, int B, int Ts, int n_rows, int C, void* dx,
                    void* stream);
int fs2_lr_bwd_f32(const float* dout, const int64_t* cum, int B, int Ts, int n_rows, int C, float* dx,
                   void* stream);

/* ------------------------------------------------------------------------------------------ */
/* bucketize + embedding (+ add) and plain embedding (modules.py:82-102,119-128;               */
/*   lightning/systems/language/embeddings.py:25-31); tables are the f32 master parameters     */
/* ------------------------------------------------------------------------------------------ */
int fs2_bucket_embed_add_bf16(const void* x, const void* target, int target_is_f64, const float* bins,
                              int n_bins_minus_1, const float* table, int64_t rows, int C, void* y,
                              int32_t* idx_out, void* stream);
int fs2_embedding_fwd_bf16(const int64_t* ids, const float* table, int64_t rows, int C,
                           int n_rows_table, int pad_idx, void* y, void* stream);
int fs2_embedding_bwd_f32(const void* dy, const void* ids, int ids_is_i64, int64_t rows, int C,
                          int n_rows_table, int pad_idx, float* dtable, void* stream);
/* MultilingualEmbedding (lightning/systems/language/embeddings.py:25-31) without the per-call torch.cat: ids index the
   concatenation of n_tables (<= 32) f32 tables; tables / dtables / table_rows are HOST arrays (device pointers, row
   counts) that are copied into the kernel parameters. */
int fs2_embedding_multi_fwd_bf16(const int64_t* ids, const void* const* tables, const int32_t* table_rows,
                                 int n_tables, int64_t rows, int C, int pad_idx, void* y, void* stream);
int fs2_embedding_multi_bwd_f32(const void* dy, const int64_t* ids, const void* const* dtables,
                                const int32_t* table_rows, int n_tables, int64_t rows, int C, int pad_idx,
                                void* stream);


/* ------------------------------------------------------------------------------------------ */
/* Small helpers: sinusoid add (transformer/Models.py:155-157,224-226), casts, Conv1d weight   */
/* packing [Co][Ci][k] f32 -> [Co][k][Cpad] bf16, broadcast row add (fastspeech2m.py:89,101,   */
/* 136), column sums (bias / speaker-row gradients), the 256->1 predictor head                 */
/* (modules.py:242,246-252)                                                                    */
/* ------------------------------------------------------------------------------------------ */
int fs2_posenc_add(const void* x, int x_is_f32, int64_t x_batch_stride, const float* pe, int B, int T,
                   int C, void* y, void* stream);
int fs2_cast_f32_bf16(const float* x, int64_t n, void* y, void* stream);
int fs2_cast_bf16_f32(const void* x, int64_t n, float* y, void* stream);
int fs2_pack_conv_weight(const float* w, int Co, int Ci, int k, int Cpad, void* wp, void* stream);
int fs2_add_rowvec_bf16(const void* x, const float* e, int B, int T, int C, void* y, void* stream);
int fs2_add_f32_bf16(const float* a, const void* b, int64_t n, float* out, void* stream);
int fs2_colsum_bf16(const void* x, int64_t ld, int groups, int rows_per_group, int C, float* out,
                    void* stream);
/* out[c] += sum_b sum_{t < lens[b]} x[b][t][c]: bias gradients of ragged batches (rows at / after lens[b]
   are padding whose gradient is zero by construction, transformer/Layers.py:25,28; they are not read) */
int fs2_colsum_ragged_bf16(const void* x, int64_t ld, int B, int T, int C, const int64_t* lens, float* out,
                           void* stream);
/* two column segments of the same rows in one launch: out0 += colsum(x[:, 0:C]), out1 += colsum(x[:, seg_stride:
   seg_stride + C]) -- the Q and V bias gradients from the fused dQ|dK|dV buffer */
int fs2_colsum_ragged2_bf16(const void* x, int64_t ld, int B, int T, int C, const int64_t* lens, int seg_stride,
                            float* out0, float* out1, void* stream);
int fs2_colsum_f32(const float* x, int64_t ld, int rows, int C, float* out, void* stream);
/* column sums of x[:, 0:3*seg_cols] split into three outputs (fused Q|K|V bias gradients) */
int fs2_colsum3_bf16(const void* x, int64_t ld, int rows, int seg_cols, float* out0, float* out1, float* out2,
                     void* stream);
/* grad[co][ci][tap] += packed[co][tap][ci]  (Conv1d weight gradient accumulated in the GEMM-friendly layout) */
int fs2_unpack_add_conv_grad(const float* packed, int Co, int Ci, int k, float* grad, void* stream);
int fs2_rowdot_fwd(const void* x, const float* w, const float* bias, const int64_t* lens, int B, int T,
                   int C, float* out, void* stream);
int fs2_rowdot_bwd(const float* dout, const void* x, const float* w, const int64_t* lens, int B, int T,
                   int C, void* dx, float* dw, float* db, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* FastSpeech2Loss (lightning/model/loss.py:15-89): out10 = total, mel, postnet, pitch, energy, */
/* duration, N_mel, N_pitch, N_energy, N_duration.  partials: f32 workspace of                  */
/* fs2_loss_workspace_floats(B, max(Ts, p_T, e_T), Tm, n_mel) elements that the caller ZEROES    */
/* ONCE (it ends with the block arrival counter of the one-kernel reduction, which every call   */
/* leaves at zero) and does not share between concurrently running calls.  Pitch / energy are   */
/* phoneme-level (p_T = Ts, p_lens = src_lens) or frame-level (p_T = Tm, p_lens = mel_lens),    */
/* loss.py:47-60; p_ld / e_ld = row stride of the target (>= p_T / e_T).                        */
/* ------------------------------------------------------------------------------------------ */
int64_t fs2_loss_workspace_floats(int B, int T_feat, int Tm, int n_mel);
int fs2_loss_fwd(const float* mel_pred, const float* post_pred, const float* mel_tgt, const float* p_pred,
                 const float* p_tgt, int p_T, int p_ld, const int64_t* p_lens, const float* e_pred, const void* e_tgt,
                 int e_tgt_is_f64, int e_T, int e_ld, const int64_t* e_lens, const float* d_pred,
                 const int64_t* d_tgt, const int64_t* src_lens, const int64_t* mel_lens, int B, int Ts, int Tm,
                 int Tm_tgt, int n_mel, float* partials, float* out10, void* stream);
int fs2_loss_bwd(const float* gout6, const float* out10, const float* mel_pred, const float* post_pred,
                 const float* mel_tgt, const float* p_pred, const float* p_tgt, int p_T, int p_ld,
                 const int64_t* p_lens, const float* e_pred, const void* e_tgt, int e_tgt_is_f64, int e_T, int e_ld,
                 const int64_t* e_lens, const float* d_pred, const int64_t* d_tgt, const int64_t* src_lens,
                 const int64_t* mel_lens, int B, int Ts, int Tm, int Tm_tgt, int n_mel, float* d_mel, float* d_post,
                 float* d_p, float* d_e, float* d_d, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* PostNet BatchNorm1d(train) + tanh + dropout (transformer/Layers.py:129-137,                 */
/* fastspeech2m.py:145); y: bf16 [M][C] conv output; stats: f32 [2][C] (sum, sum of squares),   */
/* WRITTEN by fs2_bn_stats_bf16 with a fixed summation order (bit-reproducible, no atomics);    */
/* ws: f32 scratch of fs2_bn_workspace_floats(M, C) elements (per-block partial sums).          */
/* fs2_bn_stats_bf16 also performs nn.BatchNorm1d's running-statistics update (momentum,        */
/* unbiased variance, num_batches_tracked += 1) when running_mean is non-NULL.                  */
/* fs2_bn_bwd: dstats f32 [2][C] written (dbeta, dgamma); dbeta_acc / dgamma_acc (optional,     */
/* f32 [C]) accumulate them into the parameter gradients.  keep_out / keep_in (uint8 [M][C/8]):   */
/* dropout keep bits (bit j of byte [m][c/8] = channel 8*(c/8)+j kept) written by the forward    */
/* apply and READ by the backward (required there when p_drop > 0: nothing is regenerated).      */
/* ------------------------------------------------------------------------------------------ */
int64_t fs2_bn_workspace_floats(int64_t M, int C);
int fs2_bn_stats_bf16(const void* y, int64_t M, int C, float* ws, float* stats, float momentum, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, void* stream);
int fs2_bn_apply_fwd(const void* y, const float* stats, const float* gamma, const float* beta, int64_t M,
                     int C, int act_tanh, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* out_bf16, float* out_f32, const float* res_f32, uint8_t* keep_out, void* stream);
int fs2_bn_bwd(const void* dout, int dout_is_f32, const void* y, const float* stats, const float* gamma,
               const float* beta, int64_t M, int C, int act_tanh, float p_drop, uint64_t seed,
               const uint64_t* seed_dev, const uint8_t* keep_in, float* ws, float* dstats, float* dbeta_acc,
               float* dgamma_acc, void* dy, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Phoneme-embedding front-end of the few-shot systems                      */
/*  - PhonemeQueryExtractor "average" (lightning/model/reduction.py:42-110): per utterance,       */
/*    table_sum[cls[i]] += mean (two_stage) or sum of the frames of segment i; x: f32 [T][D],     */
/*    dur / cls: int64 [L] device; count[c] += 1 (two_stage) or #frames; then finalize = sum/count */
/*  - SoftMultiAttCodebook2 (lightning/systems/language/embeddings.py:77-142): layer-weighted sum */
/*    (softmax(w_raw) over n_layer, NaN -> 0; w_raw NULL = plain cast of [rows][1][D]), then       */
/*    multi-head attention of q [rows][E] over att_banks / emb_banks [C][E] with H heads           */
/*    (transformer/Modules.py:28-47); p: f32 [rows][H][C] saved for the backward; d_att / d_emb    */
/*    are accumulated with atomics (zero them first).                                             */
/* ------------------------------------------------------------------------------------------ */
int fs2_segment_class_accum_f32(const float* x, const int64_t* dur, const int64_t* cls, int L, int64_t T, int64_t D,
                                int n_classes, int two_stage, float* table_sum, float* count, void* stream);
int fs2_class_mean_finalize_f32(float* table, const float* count, int n_classes, int64_t D, void* stream);
int fs2_layer_weighted_sum_bf16(const float* ref, const float* w_raw, int64_t rows, int n_layer, int D, void* out,
                                void* stream);
int fs2_codebook_attn_fwd_f32(const float* q, const float* att_banks, const float* emb_banks, int rows, int C, int E,
                              int H, float inv_temp, float* out, float* p, void* stream);
int fs2_codebook_attn_bwd_f32(const float* dout, const float* q, const float* att_banks, const float* emb_banks,
                              const float* p, int rows, int C, int E, int H, float inv_temp, float* dq, float* d_att,
                              float* d_emb, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Device-side collate: zero-padding of ragged rows -- pad_1D / pad_2D of    */
/* lightning/utils/tool.py:134-165 as used by lightning/collates/utils.py:8-85.                   */
/* src: rows of all B items back to back ([sum len_b][row_bytes]); offsets: int64 [B+1] (device);  */
/* dst: [B][max_len][row_bytes]; row_bytes % 4 == 0.                                               */
/* ------------------------------------------------------------------------------------------ */
int fs2_pad_ragged(const void* src, const int64_t* offsets, int B, int max_len, int row_bytes, void* dst,
                   void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Multi-tensor weight refresh: every bf16 operand copy of the fp32 master weights in ONE launch */
/* (nn.Linear: cast; nn.Conv1d [Co][Ci][k] -> packed [Co][k][Cpad], as fs2_pack_conv_weight;     */
/* kind 1: plain f32 copy, used to gather the Q|K|V biases).  `table` is a DEVICE array.          */
/* ------------------------------------------------------------------------------------------ */
typedef struct fs2_prep_entry {
  const float* src;  /* f32 [rows][ci][k] */
  void* dst;         /* kind 0: bf16 [rows][k][cpad]; kind 1: f32 [rows][ci*k] */
  int32_t rows, ci, k, cpad;
  int32_t kind;
  int32_t row0;      /* sum of `rows` of the entries before this one */
} fs2_prep_entry;
int fs2_weight_prep(const void* table, int n_entries, int total_rows, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused gradient clipping + Adam + LR schedule on flat fp32 buffers:       */
/*   clip  : pytorch_lightning gradient_clip_val (main.py:104-110) = clip_grad_norm_(max_norm, 2) */
/*   Adam  : torch.optim.Adam(betas, eps, weight_decay)        (lightning/optimizer.py:5-16)      */
/*   sched : LambdaLR sqrt_schedule (1) / const_schedule (2)   (lightning/scheduler.py:21-62)     */
/* p, g, m, v: f32 [n] device (16-byte aligned, n % 4 == 0); gnorm_sq: f32 [1] device, sum of g^2 */
/* (fs2_sumsq_f32 accumulates into it; zero it first); step: int64 [1] device = optimizer steps    */
/* already taken.  anneal_steps is a HOST array of n_anneal <= 8 ints (copied into the launch).    */
/* fs2_optim_advance: step += 1, *gnorm_out = sqrt(*gnorm_sq) (optional), *gnorm_sq = 0.           */
/* ------------------------------------------------------------------------------------------ */
int fs2_sumsq_f32(const float* g, int64_t n, float* out, void* stream);
int fs2_adam_step_f32(float* p, const float* g, float* m, float* v, int64_t n, const float* gnorm_sq,
                      const int64_t* step, float lr0, float beta1, float beta2, float eps, float weight_decay,
                      float max_norm, int sched_type, int warmup, const int32_t* anneal_steps, int n_anneal,
                      float anneal_rate, void* stream);
int fs2_optim_advance(int64_t* step, float* gnorm_sq, float* gnorm_out, void* stream);
/* y += alpha * x on flat f32 buffers (16-byte aligned): the inner-loop SGD step of the first-order few-shot      */
/* adaptation, theta' = theta - lr * grad (config/algorithm/language/fscl-orig.yaml:38-43).      */
int fs2_axpy_f32(float* y, const float* x, float alpha, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2B200_H_ */
