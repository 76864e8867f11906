/* dicow_b200.h -- C ABI of libdicow_b200.so: the B200 (sm_100a) kernels behind the DiCoW / SE-DiCoW hot path.
 *
 * The reference (BUTSpeechFIT/TS-ASR-Whisper) has NO native interface: every FLOP of its hot path is a PyTorch /
 * HF-transformers library call.  Each entry point below therefore cites the reference *Python call site* it
 * replaces (paths relative to /root/reference, "HF:" = transformers/models/whisper).  See INTEGRATION.md for the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", POD arguments only: raw device pointers, sizes, a cudaStream_t passed as void*.
 *   - The caller owns every buffer (inputs, outputs, workspaces); the library never allocates or frees caller
 *     memory and never synchronises the device: all work is enqueued on the given stream.
 *   - Every function returns 0 on success or a dicow_status_t; dicow_last_error(h) gives the message.
 *   - A handle is not thread-safe; distinct handles/streams are independent.
 *   - bf16 tensors are row-major with the stated leading dimension (elements); "ld" must keep rows 16-byte aligned.
 */
#ifndef DICOW_B200_H
#define DICOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DICOW_API __attribute__((visibility("default")))
#else
#define DICOW_API
#endif

typedef struct dicow_ctx* dicow_handle_t;

typedef enum {
  DICOW_OK = 0,
  DICOW_ERR_INVALID_ARG = 1,
  DICOW_ERR_CUDA = 2,
  DICOW_ERR_UNSUPPORTED = 3,
  DICOW_ERR_ARCH = 4 /* not an sm_100 device */
} dicow_status_t;

DICOW_API int dicow_create(int device, dicow_handle_t* out);
DICOW_API int dicow_destroy(dicow_handle_t h);
DICOW_API const char* dicow_last_error(dicow_handle_t h);
/* returns DICOW_ERR_CUDA (and records the message) if the device has a pending/sticky error */
DICOW_API int dicow_check(dicow_handle_t h);
/* Number of SMs the persistent kernels (GEMM, attention, LayerNorm rings) size their grids for in calls made AFTER this one;
 * sms <= 0 or >= the device's SM count restores the whole device.  Returns the value in effect.  Used to run one encoder
 * pass on part of the device next to latency-bound decode steps on another stream (long-form speculation: the reference's
 * seek loop, src/models/dicow/generation.py:415-534, encodes window n + 1 only after window n is decoded). */
DICOW_API int dicow_set_sm_budget(dicow_handle_t h, int sms);
/* ABI version of this header (bumped when a struct changes) */
DICOW_API int dicow_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores:  D[b, m, :] = epilogue( sum_k A[b, m, k] * W[:, k] ),  bf16 x bf16 -> fp32.
 * Replaces every nn.Linear / nn.Conv1d on the path:
 *   q/k/v/out_proj   HF:modeling_whisper.py:310,331-332,355     fc1/fc2  HF:modeling_whisper.py:404-408
 *   conv1/conv2      src/models/dicow/encoder.py:167-168        (implicit GEMM: A is an overlapping-row view)
 *   SCB ffn          src/models/dicow/layers.py:138-143,163-166 (concat folded: A = [A | A2] split along K)
 *   lm_head/proj_out src/models/dicow/encoder.py:236, modeling_dicow.py:302
 *
 * The M dimension is a batch of nb independent row blocks of Mb rows (nb = 1 for a plain GEMM).  Row m of batch b
 * of A starts at A + b*a_batch_stride + m*lda (elements) -- lda may be smaller than K (overlapping rows), which is
 * how Conv1d(k=3) over a zero-padded channels-last buffer becomes a GEMM without materialising im2col.
 * ------------------------------------------------------------------------------------------------------------ */
typedef enum {
  DICOW_EPI_BIAS_BF16 = 0,      /* out_bf16 = acc + bias                                                        */
  DICOW_EPI_BIAS_GELU_BF16 = 1, /* out_bf16 = gelu_erf(acc + bias)                                              */
  DICOW_EPI_RESIDUAL_F32 = 2,   /* out_f32  = resid + alpha * (acc + bias); alpha = tanh(*gate) or 1 if gate NULL */
  DICOW_EPI_BIAS_F32 = 3,       /* out_f32  = acc + bias                                                        */
  DICOW_EPI_GELU_FDDT_POS_F32 = 4, /* out_f32 = FDDT(gelu_erf(acc + bias), stno) + pos[m, :]   (conv2 epilogue:
                                      src/models/dicow/encoder.py:168-179)                                       */
  DICOW_EPI_ACCUM_F32 = 5,         /* out_f32 += alpha * acc  (atomic fp32 adds; alpha = *gate or 1): weight-gradient
                                      accumulation, contraction split over CTAs (splits)                          */
  DICOW_EPI_GELU_SAVE_BF16 = 6,    /* training forward: aux_bf16 = acc + bias (pre-activation), out_bf16 = gelu_erf(.)  */
  DICOW_EPI_DGELU_BF16 = 7         /* backward: out_bf16 = acc * gelu_erf'(aux_bf16)   (aux has out's leading dimension) */
} dicow_epilogue_t;

typedef struct {
  size_t struct_size; /* = sizeof(dicow_gemm_args_t) */
  /* A operand(s), bf16 */
  const void* A;
  int64_t lda;            /* row stride, elements */
  int64_t a_batch_stride; /* elements */
  const void* A2;         /* optional second source for k >= K1 (NULL if unused) */
  int64_t lda2;
  int64_t a2_batch_stride;
  int32_t K1; /* split point (multiple of 64) when A2 != NULL */
  /* W operand: [N, K] row-major bf16 (nn.Linear weight layout) */
  const void* W;
  int64_t ldw;
  int32_t nb, Mb, N, K;
  const float* bias; /* [N] fp32 or NULL */
  /* output */
  void* out;
  int64_t ldo;              /* elements */
  int64_t out_batch_stride; /* elements */
  int32_t epilogue;         /* dicow_epilogue_t */
  /* DICOW_EPI_RESIDUAL_F32 */
  const float* resid; /* may alias out */
  int64_t ldr;
  int64_t resid_batch_stride;
  const float* gate; /* device scalar; alpha = tanhf(*gate) */
  /* DICOW_EPI_GELU_FDDT_POS_F32 */
  const float* stno;   /* [nb, 4, Mb] fp32, class order S,T,N,O (src/models/dicow/FDDT.py:41-63) */
  int64_t stno_batch_stride; /* elements between batches (4*Mb when dense) */
  const float* fddt_w; /* [4, N] fp32, rows in S,T,N,O order; NULL = bias-only FDDT */
  const float* fddt_b; /* [4, N] */
  const float* pos;    /* [Mb, N] fp32 (embed_positions.weight) or NULL */
  int32_t flags;       /* 0 = auto; bit 0: force the single-CTA kernel; bit 1: force the CTA-pair (cta_group::2) kernel;
                          bit 2: A is given transposed, At[k][m] with row stride lda (contraction index on the rows);
                          bit 3: W is given transposed, Wt[k][n] with row stride ldw.  Transposed operands are consumed
                          MN-major by the tensor core -- no transposed copies: dgrad dX = dY W uses bit 3 with the
                          forward weight, wgrad dW = dY^T X uses bits 2 | 3 with the activations                      */
  int32_t splits;      /* DICOW_EPI_ACCUM_F32: 0 = choose, n > 1 = split the contraction n ways, 1 = no split        */
  void* aux_bf16;      /* DICOW_EPI_GELU_SAVE_BF16 (written) / DICOW_EPI_DGELU_BF16 (read): pre-activation, ld = ldo   */
} dicow_gemm_args_t;

DICOW_API int dicow_gemm_bf16(dicow_handle_t h, const dicow_gemm_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused FDDT + LayerNorm over the fp32 residual stream (HBM-bound, one warp per row, warp-shuffle reductions).
 *   x'[b,t,:] = sum_c stno[b,c,t] * (w_c (.) x[b,t,:] + b_c)          src/models/dicow/FDDT.py:52-62
 *   ln        = LayerNorm(x') * gamma + beta  (eps inside the sqrt)     HF:modeling_whisper.py:393,403; encoder.py:228
 * Replaces the ~15 eager element-wise launches of src/models/dicow/encoder.py:205-206 plus nn.LayerNorm.
 * Optional pending residual updates delta1 / delta2 (the bf16 outputs of out_proj and fc2: under the reference's bf16
 * autocast the Linear output is bf16 and the residual sum fp32, HF:modeling_whisper.py:400,411) are added first:
 *   x' = FDDT(x + delta1 + delta2).  x' is written back to x when store_x != 0.  Any of the outputs may be NULL.
 * fddt_w == NULL selects the bias-only FDDT variant (FDDT.py:43-51).  rows = B*T; row r uses mask column (r / T, :, r % T).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  float* x; /* [rows, d] fp32 */
  int32_t rows, d, T;
  const float* stno; /* [B, 4, T] fp32 or NULL (no FDDT) */
  int64_t stno_batch_stride;
  const float* fddt_w; /* [4, d] fp32 rows S,T,N,O, or NULL */
  const float* fddt_b; /* [4, d] */
  const float* gamma;  /* [d] or NULL (no LayerNorm) */
  const float* beta;
  float eps;
  void* ln_out_bf16; /* [rows, d] bf16 or NULL */
  float* ln_out_f32; /* [rows, d] fp32 or NULL */
  void* x_out_bf16;  /* [rows, d] bf16 copy of x' or NULL */
  const void* delta1_bf16; /* [rows, d] bf16 or NULL */
  const void* delta2_bf16; /* [rows, d] bf16 or NULL */
  int32_t store_x;         /* write x' back into x (or into x_out) */
  int32_t flags;           /* 0 = default (TMA-pipelined rows); bit 0: warp-per-row kernel, bit 1: column-owner kernel (comparison) */
  float* x_out;            /* NULL: x' overwrites x; else x' goes to this [rows, d] fp32 buffer and x stays as it was */
} dicow_fddt_ln_args_t;

DICOW_API int dicow_fddt_layernorm(dicow_handle_t h, const dicow_fddt_ln_args_t* args, void* stream);

/* full-matrix FDDT (src/models/dicow/layers.py:7-47, FDDT.py:52-62): y bf16 [rows, >= 4 d] holds the four class transforms
 * of every row side by side (one dicow_gemm_bf16 against the stacked [4 d, d] weight, class order S, T, N, O);
 * x[r, :] = sum_c stno[r / T, c, r % T] * y[r, c d : (c + 1) d] (+ pos[r % T, :] when pos != NULL), fp32. */
DICOW_API int dicow_fddt_full_combine(dicow_handle_t h, const void* y_bf16, int64_t ldy, const float* stno,
                                      int64_t stno_batch_stride, int T, int rows, int d, const float* pos, float* x,
                                      void* stream);
/* backward of dicow_fddt_full_combine: dy[r, c d : (c + 1) d] = stno[r / T, c, r % T] * g[r, :] (bf16 [rows, >= 4 d]); what
 * autograd computes for `sum_c mask_c * CustomLinear_c(x)` (FDDT.py:52-62) before the four Linear backward passes. */
DICOW_API int dicow_fddt_full_scatter(dicow_handle_t h, const float* g, const float* stno, int64_t stno_batch_stride, int T,
                                      int rows, int d, void* dy_bf16, int64_t ldy, void* stream);
/* input_features fp32 [B, C, F] -> zero-padded channels-last bf16 [B, F + 2, C] (the buffer conv1's implicit GEMM
 * reads; src/models/dicow/encoder.py:167 nn.Conv1d(padding=1)). */
DICOW_API int dicow_features_to_channels_last(dicow_handle_t h, const float* in, void* out_bf16, int B, int C, int F,
                                              void* stream);
/* zero rows 0 and T+1 of a channels-last [B, T + 2, C] bf16 buffer */
DICOW_API int dicow_zero_pad_rows(dicow_handle_t h, void* buf_bf16, int B, int T, int C, void* stream);
/* fp32 -> bf16 cast (weight preparation) */
DICOW_API int dicow_cast_f32_bf16(dicow_handle_t h, const float* in, void* out_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-head attention, head_dim 64, flash-style on tcgen05: S = Q K^T into TMEM, online softmax (one thread per
 * query row), P (bf16) V accumulated in TMEM.  softmax(Q K^T) V with NO extra scaling (the reference scales q by
 * hd^-0.5 before the product, HF:modeling_whisper.py:305-310 -- folded into the q projection weights) and no padding
 * mask; causal = 1 applies the decoder's lower-triangular mask (key <= query + Tk - Tq).
 * Replaces HF WhisperAttention's SDPA call (HF:modeling_whisper.py:342-352) for the encoder self-attention, the
 * SE-DiCoW enrollment cross-attention (src/models/dicow/layers.py:156-160) and decoder prefill.
 *
 * Element (b, t, h, e) of Q lives at Q + b*q_batch_stride + t*q_row_stride + h*64 + e (elements, bf16); likewise
 * K, V (kv strides) and the output.  All strides must be multiples of 8 elements, bases 16-byte aligned.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  const void* Q;
  const void* K;
  const void* V;
  void* out; /* bf16 */
  int32_t B, H, Tq, Tk;
  int64_t q_row_stride, q_batch_stride;
  int64_t kv_row_stride, kv_batch_stride;
  int64_t o_row_stride, o_batch_stride;
  int32_t causal;
  int32_t variant; /* 0 = default (persistent CTAs, two query tiles per CTA in ping-pong, every exponential on the SFU: the
                      lowest energy per launch, which is what a power-capped step pays); 1, 2, 3 = 2, 4, 6 of every 8
                      exponentials as a polynomial on the FMA pipe; 4 = packed bf16 exponentials; 16+ = single-tile
                      comparison kernels */
  float* lse;      /* optional [B, H, Tq] fp32: log2(sum_k 2^(s_k log2 e)) per query row, saved for dicow_attention_bwd_bf16 */
} dicow_attention_args_t;

DICOW_API int dicow_attention_bf16(dicow_handle_t h, const dicow_attention_args_t* args, void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Attention backward (training): dQ, dK, dV from dO, the forward's Q / K / V / O and its saved lse.  Non-causal shapes with
 * Tq >= 256 and a workspace of B * H * Tq * 65 + 4 floats: ONE tcgen05 pass per (batch, head, 128 keys) -- 5 GEMMs, dQ partial
 * sums reduced with fp32 red.global into the workspace (summation order not deterministic).  Otherwise two passes
 * (dQ; dK + dV) that recompute the probabilities; no atomics (workspace: B * H * Tq floats).  dO shares O's strides; dK / dV
 * share one stride pair.  Replaces autograd through the SDPA call of HF WhisperAttention
 * (HF:modeling_whisper.py:342-352) in the fine-tuning step.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  const void* Q;
  const void* K;
  const void* V;
  const void* O;
  const void* dO;
  const float* lse; /* [B, H, Tq] from dicow_attention_bf16 */
  void* dQ;
  void* dK;
  void* dV;
  int32_t B, H, Tq, Tk;
  int64_t q_row_stride, q_batch_stride;
  int64_t kv_row_stride, kv_batch_stride;
  int64_t o_row_stride, o_batch_stride;
  int64_t dq_row_stride, dq_batch_stride;
  int64_t dkv_row_stride, dkv_batch_stride;
  int32_t causal;
  float* workspace;
  int64_t workspace_floats; /* size of workspace; >= B*H*Tq*65 + 4 enables the single-pass kernel (0: B*H*Tq, two passes) */
} dicow_attention_bwd_args_t;
DICOW_API int dicow_attention_bwd_bf16(dicow_handle_t h, const dicow_attention_bwd_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward of dicow_fddt_layernorm: given dy = dL/d(LayerNorm output) (bf16) and g_in = dL/dx' arriving through the
 * residual stream (fp32, or NULL), recompute x' = FDDT(x + delta1 + delta2) and write g_out = dL/d(x + delta1 + delta2)
 * (fp32, + optional bf16 copy for the next dgrad / wgrad GEMM); accumulate (+=) dgamma, dbeta [d] and the FDDT table
 * gradients dfddt_w, dfddt_b [4, d].  gamma == NULL: no LayerNorm on the path (FDDT backward of g_in only).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  const float* x;
  const void* delta1_bf16;
  const void* delta2_bf16;
  int32_t rows, d, T;
  const float* stno;
  int64_t stno_batch_stride;
  const float* fddt_w;
  const float* fddt_b;
  const float* gamma;
  float eps;
  const void* dy_bf16;
  const float* g_in;
  float* g_out;
  void* g_out_bf16;
  float* dgamma;
  float* dbeta;
  float* dfddt_w;
  float* dfddt_b;
  float* g_colsum; /* [d] += column sums of the rows written to g_out_bf16 (= the bias gradient of the Linear whose output
                      was a pending delta of this LayerNorm), or NULL */
} dicow_ln_bwd_args_t;
DICOW_API int dicow_layernorm_fddt_bwd(dicow_handle_t h, const dicow_ln_bwd_args_t* args, void* stream);

/* out[n] += alpha * sum_rows x[row, n]  (bias gradients); x bf16 (is_bf16 = 1) or fp32, leading dimension ld */
DICOW_API int dicow_colsum(dicow_handle_t h, const void* x, int is_bf16, int64_t ld, int rows, int N, float* out, float alpha,
                           void* stream);
/* col2im of Conv1d(k = 3, padding = 1, stride): dcol bf16 [B, T_out, 3 C] (the dgrad GEMM's output in im2col layout)
 * -> dx bf16 [B, T, C] addressed by dx_batch_stride / dx_row_stride (elements) */
DICOW_API int dicow_conv1d_col2im(dicow_handle_t h, const void* dcol_bf16, void* dx_bf16, int B, int T, int T_out, int C,
                                  int stride, int64_t dx_batch_stride, int64_t dx_row_stride, void* stream);

/* out_bf16[r, c] = g[r, c] * gelu_erf'(pre_bf16[r, c]); g fp32 or bf16 (g_is_bf16); leading dimensions in elements.
 * Backward of the GELUs of the conv stem (src/models/dicow/encoder.py:167-168) where no dgrad GEMM epilogue sits. */
DICOW_API int dicow_dgelu_mul(dicow_handle_t h, const void* g, int g_is_bf16, int64_t ldg, const void* pre_bf16, int64_t ldp,
                              void* out_bf16, int64_t ldo, int rows, int cols, void* stream);
/* backward of the SE-DiCoW gate (src/models/dicow/layers.py:79-93,168: q + tanh(gate) * upd) for g = dL/d(output) fp32:
 * dupd_bf16[r, c] = tanh(*gate) * g[r, c];  *dgate += (1 - tanh^2(*gate)) * sum_rc g[r, c] * upd_bf16[r, c]  (dgate may be
 * NULL).  cols and the leading dimensions (elements) must be even. */
DICOW_API int dicow_gate_bwd(dicow_handle_t h, const float* g, int64_t ldg, const void* upd_bf16, int64_t ldu, const float* gate,
                             void* dupd_bf16, int64_t ldd, int rows, int cols, float* dgate, void* stream);
/* out_bf16[r, c] = in[r, c] for c < cols, 0 for cols <= c < cols_out: an fp32 gradient (e.g. d logits handed over by
 * autograd) re-laid as the zero-padded bf16 operand of the dgrad / wgrad GEMMs */
DICOW_API int dicow_cast_f32_bf16_2d(dicow_handle_t h, const float* in, int64_t ldi, void* out_bf16, int64_t ldo, int rows,
                                     int cols, int cols_out, void* stream);
/* backward of dicow_embed_tokens (HF:modeling_whisper.py:739-760 through autograd): g fp32 [rows, d], rows = B * S;
 * d_embed_tokens[ids[r], :] += g[r, :] (ids outside [0, vocab) skipped), d_embed_positions[past + r % S, :] += g[r, :];
 * either gradient pointer may be NULL. */
DICOW_API int dicow_embedding_bwd(dicow_handle_t h, const float* g, const int64_t* ids, int rows, int S, int d, int past,
                                  int vocab, float* d_embed_tokens, float* d_embed_positions, void* stream);

/* CTC backward (src/models/dicow/encoder.py:123-134 through autograd): dlogits_bf16[b, t, :V1] = loss_scale * dL/dlogits,
 * columns [V1, ldd) zeroed (ldd: leading dimension, a multiple of 8 so the buffer feeds the wgrad GEMM).
 * lse: per-row natural-log sum exp (B * T floats, e.g. the first B * T floats of dicow_ctc_loss's workspace).
 * workspace: (2 * B * T * (2 Lmax + 1) + B) floats. */
typedef struct {
  size_t struct_size;
  const float* logits;
  const float* lse;
  int32_t B, T, V1;
  const int64_t* labels;
  int32_t Lmax;
  int32_t reduction_mean;
  float loss_scale;
  float* workspace;
  void* dlogits_bf16;
  int64_t ldd;
  const float* scale_dev; /* optional device scalar multiplying loss_scale (the upstream gradient stays on the GPU) */
  int32_t out_f32;        /* != 0: dlogits is fp32 (what autograd hands to a caller-visible logits tensor) */
  int64_t ld;             /* row stride of logits in elements (0 = V1: contiguous rows) */
} dicow_ctc_bwd_args_t;
DICOW_API int dicow_ctc_loss_bwd(dicow_handle_t h, const dicow_ctc_bwd_args_t* args, void* stream);

/* backward of dicow_softlabel_ce: dlogits_bf16[r, :V] = scale * (softmax(logits[r]) - target of the winning case stream),
 * 0 for masked rows; scale = upstream gradient / (number of unmasked rows | rows) supplied by the caller. */
typedef struct {
  size_t struct_size;
  const float* logits;
  int64_t ld;
  int32_t rows, V;
  const int64_t* labels;
  const int64_t* upp_labels;
  int32_t ts_begin, n_ts;
  const float* smoothing;
  int32_t soft_mode;
  float scale;
  void* dlogits_bf16;
  int64_t ldd;
  const float* scale_dev; /* optional device scalar multiplying scale */
} dicow_softlabel_ce_bwd_args_t;
DICOW_API int dicow_softlabel_ce_bwd(dicow_handle_t h, const dicow_softlabel_ce_bwd_args_t* args, void* stream);

/* debug aid: clock64 stamps of one CTA's KV loop are written to buf ([steps][8] int64); NULL disables */
DICOW_API int dicow_debug_set_attention_profile(dicow_handle_t h, void* buf);

/* ------------------------------------------------------------------------------------------------------------
 * Whisper log-mel front-end (n_fft 400, hop 160, periodic Hann, centred reflect-padded STFT, last frame dropped, power,
 * mel filterbank, log10(max(., 1e-10)), floor at (max over the whole recording - 8), (x + 4) / 4).
 * Replaces WhisperFeatureExtractor._torch_extract_fbank_features (HF:models/whisper/feature_extraction_whisper.py:135-164)
 * as called at src/data/local_datasets.py:208-214; one batch row = one (possibly multi-window) recording, zero-padded
 * by the caller to n_pad samples (a multiple of 480000 in the reference; any multiple of 160 here).
 * out[b, m, t], t < n_pad / 160.  attention_mask[b, t] = (t * 160 < lengths[b])  (feature_extraction_whisper.py:328-337).
 * workspace: B * 4 bytes of device scratch (per-recording running max).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  const float* audio; /* [B, n_pad] fp32 */
  int64_t audio_batch_stride;
  int32_t B;
  int64_t n_pad;
  const int64_t* lengths;   /* [B] valid samples per recording (device) or NULL */
  const float* mel_filters; /* [201, n_mels] fp32 (WhisperFeatureExtractor.mel_filters) */
  int32_t n_mels;
  float* out;              /* [B, n_mels, n_pad / 160] fp32 */
  int32_t* attention_mask; /* [B, n_pad / 160] or NULL */
  void* workspace;
} dicow_logmel_args_t;

DICOW_API int dicow_logmel(dicow_handle_t h, const dicow_logmel_args_t* args, void* stream);

/* STNO mask of one recording from per-speaker sample-level activity (src/data/local_datasets.py:162-196: get_stno_mask /
 * _create_stno_masks).  activity: uint8 [n_speakers, ld >= n_samples] (non-zero = speaking); target = row of the target
 * speaker or -1; frames = padded length / frame_samples (320); out[t * frame_stride + c * class_stride], c = S, T, N, O. */
DICOW_API int dicow_stno_mask(dicow_handle_t h, const uint8_t* activity, int64_t ld, int n_speakers, int64_t n_samples,
                              int target, int frame_samples, int64_t frames, float* out, int64_t frame_stride,
                              int64_t class_stride, void* stream);

/* Training-time augmentations of a padded batch (src/data/collators.py:184-210, DataCollator.__call__: soft_segment_augmentation
 * :77-136, add_gaussian_noise_and_rescale :50-75, SpecAug on [mel || STNO repeated x factor] :205-210 with
 * src/data/augmentations.py time_warp :70-98 / mask_along_axis :20-66 / SpecAug.forward :414-434).  The reference draws every
 * random number from torch's CPU generator in a data-independent order; the caller draws them in that order (the "plan") and
 * this call applies the plan, in the reference's order: segments, noise (both in place on `stno`), SpecAug (feats, stno ->
 * feats_out, stno_out; skipped when spec == 0).  All device pointers; fp32 tensors contiguous. */
typedef struct {
  size_t struct_size;
  float* stno;             /* [B, C, Ts] class probabilities (C = 4: S, T, N, O) */
  int32_t B, C, Ts;
  const int32_t* seg;      /* [n_seg, 4]: batch row, start, end, index into the classes other than the dominant one */
  const float* seg_soft;   /* [n_seg, 2]: softness, 1 - softness (the latter rounded from double, as the reference's scalar is) */
  int32_t n_seg;
  const int32_t* noise_rows; /* [n_noise] distinct batch rows */
  const float* noise;      /* [n_noise, C, Ts], already scaled by sqrt(variance) */
  int32_t n_noise;
  int32_t spec;            /* 0: no SpecAug */
  const float* feats;      /* [B, M, Tf], Tf = factor * Ts */
  float* feats_out;        /* [B, M, Tf] */
  float* stno_out;         /* [B, C, Ts] */
  int32_t M, Tf, factor;
  int32_t center, warped;  /* time warp: frames [0, center) -> [0, warped), the rest -> the rest (bicubic); center < 0: none */
  const int32_t* freq_masks; /* [B, n_freq_masks, 2] (pos, length) over the first min(mask_channels, M + C) channels */
  int32_t n_freq_masks;
  const int32_t* time_masks; /* [B, n_time_masks, 2] (pos, length) in mel frames */
  int32_t n_time_masks;
  int32_t mask_channels;   /* 128 in the reference (augmentations.py:425-431), whatever M is */
} dicow_augment_args_t;

DICOW_API int dicow_augment_batch(dicow_handle_t h, const dicow_augment_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Token-by-token decoder step (greedy generate()).  Replaces, per generated token, HF WhisperDecoder.forward with a KV
 * cache (HF:modeling_whisper.py:449-506, 691-796), proj_out (src/models/dicow/modeling_dicow.py:302) and the greedy
 * branch of DiCoWGenerationMixin._sample (src/models/dicow/generation.py:707-782).  Every entry point optionally reads
 * the current position from a device scalar `pos` so that one step is a fixed launch sequence (CUDA-graph replayable).
 * ------------------------------------------------------------------------------------------------------------ */

/* out[m, :] = epilogue(A[m, :] W^T + bias) for M <= 64 rows: weights streamed once from HBM (mma.sync m16n8k16).
 * epilogue: DICOW_EPI_BIAS_BF16 | DICOW_EPI_BIAS_GELU_BF16 | DICOW_EPI_RESIDUAL_F32 (out = resid + acc + bias) |
 * DICOW_EPI_BIAS_F32.  If pos != NULL the output base is advanced by (*pos) * pos_stride elements (KV-cache append). */
typedef struct {
  size_t struct_size;
  const void* A; /* bf16 [M, K] */
  int64_t lda;
  const void* W; /* bf16 [N, K] */
  int64_t ldw;
  int32_t M, N, K;
  const float* bias;
  void* out;
  int64_t ldo;
  int32_t epilogue;
  const float* resid;
  int64_t ldr;
  const int32_t* pos;
  int64_t pos_stride;
} dicow_gemm_skinny_args_t;
DICOW_API int dicow_gemm_skinny_bf16(dicow_handle_t h, const dicow_gemm_skinny_args_t* args, void* stream);

/* [LayerNorm ->] Linear of one decode step in ONE kernel (M <= 64 rows):
 *   A = x != NULL ? LayerNorm(x) * gamma + beta  (fp32 rows, two-pass statistics, K <= 1280; HF:modeling_whisper.py:471,
 *       483,498 + 779) : A (bf16);   out[m, n] = epilogue(sum_k A[m, k] W[n, k] + bias[n])   (epilogues as above).
 * Columns n >= n_split go to out2[m * ldo2 + (n - n_split)] when out2 != NULL (fused q | k,v projection:
 * HF:modeling_whisper.py:310,331-332); `pos` advances the base of out2 (of out when out2 == NULL) by *pos * pos_stride
 * elements (KV-cache append).  The kernel is launched with programmatic stream serialisation: it requests its weight
 * slab before waiting for the previous kernel of the step.  K may be any multiple of 32 that splits into <= 8 slices of
 * <= 1280 (multiples of 32); small-N layers split K over a thread-block cluster (deterministic DSMEM reduction). */
typedef struct {
  size_t struct_size;
  const float* x; /* LayerNorm source [M, K] fp32, or NULL */
  int64_t ldx;
  const float* gamma;
  const float* beta;
  float eps;
  const void* A; /* bf16 [M, K] when x == NULL */
  int64_t lda;
  const void* W; /* bf16 [N, K] */
  int64_t ldw;
  int32_t M, N, K;
  const float* bias;
  void* out;
  int64_t ldo;
  int32_t epilogue;
  const float* resid;
  int64_t ldr;
  int32_t n_split;
  void* out2;
  int64_t ldo2;
  const int32_t* pos;
  int64_t pos_stride;
} dicow_decode_linear_args_t;
DICOW_API int dicow_decode_linear(dicow_handle_t h, const dicow_decode_linear_args_t* args, void* stream);

/* The decoder LAYERS of one greedy token step in ONE persistent kernel (decode batch <= 16 rows): token + position
 * embedding, then per layer LayerNorm -> q | k,v (k,v appended to the self-attention cache at *pos), self-attention over the
 * cache, out_proj + residual, LayerNorm -> q, cross-attention over the head-major encoder K/V, out_proj + residual,
 * LayerNorm -> fc1 + GELU, fc2 + residual -- one CTA per SM, grid-wide barriers between the phases, every weight tile of a
 * phase requested at once.  Replaces the layer loop of HF WhisperDecoder.forward with a KV cache
 * (HF:models/whisper/modeling_whisper.py:449-506, 691-796) inside one step of DiCoWGenerationMixin._sample
 * (src/models/dicow/generation.py:707-782); what dicow_embed_tokens + 12 launches per layer of dicow_fddt_layernorm /
 * dicow_decode_linear / dicow_decode_attention_bf16 compute.  The final LayerNorm, proj_out, the logits rules and the
 * position advance stay separate calls.  `layers`: DEVICE array of L entries; weights bf16 [N, K] as in dicow_decode_linear
 * (wqkv = [Wq * hd^-0.5 ; Wk ; Wv], bqkv = [bq * hd^-0.5 ; 0 ; bv]); self_kv = this layer's cache [B, S_max, 2 d] (k | v);
 * cross_kv = [B, H, T, 128] (k(64) | v(64) per key).  x / q / ctx / hidden: step scratch [16, d] fp32, [16, d], [16, d],
 * [16, ffn] bf16; attn_workspace: scratch; `barrier`: one zero-initialised uint64 owned by the caller for the lifetime of the buffers (monotonic
 * arrival counter).  d_model <= 1280 (multiple of 64, head_dim 64), ffn <= 5120. */
typedef struct {
  const float* ln1_g;
  const float* ln1_b;
  const float* ln2_g;
  const float* ln2_b;
  const float* ln3_g;
  const float* ln3_b;
  const void* wqkv;
  const float* bqkv;
  const void* wo_self;
  const float* bo_self;
  const void* wq_cross;
  const float* bq_cross;
  const void* wo_cross;
  const float* bo_cross;
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
  void* self_kv;
  const void* cross_kv;
} dicow_decode_layer_args_t;
typedef dicow_decode_layer_args_t dicow_decode_layer_t;

typedef struct {
  size_t struct_size;
  int32_t B, d, H, ffn, L, T, S_max, vocab;
  const dicow_decode_layer_args_t* layers;
  const int64_t* ids;
  int64_t ids_row_stride;
  const float* embed_tokens;
  const float* embed_positions;
  const int32_t* pos;
  float* x;
  void* q;
  void* ctx;
  void* hidden;
  void* barrier;
  float* attn_workspace; /* B * H * 136 floats: cross-attention partials exchanged between two phases */
  float eps;
  int32_t flags; /* tuning switches (bit 0: cp.async staging of activations, bit 1: L2 prefetch of the next phase's weights) */
} dicow_decode_layers_args_t;
DICOW_API int dicow_decode_layers(dicow_handle_t h, const dicow_decode_layers_args_t* args, void* stream);

/* one query row per (batch, head), head_dim 64: out[b, h*64:] = softmax(q . K^T) V over Tk keys (Tk = *pos + 1 if pos).
 * Q/out bf16 [B, H*64]; K/V bf16 rows at K + b*kv_batch_stride + h*kv_head_stride + k*kv_row_stride. */
typedef struct {
  size_t struct_size;
  const void* Q;
  int64_t q_batch_stride;
  const void* K;
  const void* V;
  int64_t kv_row_stride, kv_batch_stride;
  void* out;
  int64_t o_batch_stride;
  int32_t B, H, Tk;
  const int32_t* pos;
  int64_t kv_head_stride; /* elements between heads; 0 = 64 (heads adjacent inside a row) */
  int32_t kv_batch_div;   /* > 1: K/V batch index = query row / kv_batch_div (the beams of an utterance share its cross K/V) */
  const int32_t* ancestry; /* optional [B, ancestry_stride]: key t of query row b is read from cache row ancestry[b][t] --
                              beam search keeps every hypothesis' K/V where it was computed and only re-links prefixes
                              (replaces past_key_values.reorder_cache(beam_idx), src/models/dicow/generation.py:1080-1085) */
  int64_t ancestry_stride;
} dicow_decode_attention_args_t;
DICOW_API int dicow_decode_attention_bf16(dicow_handle_t h, const dicow_decode_attention_args_t* args, void* stream);

/* cross-attention K/V of one window, [B*T, (k | v) x H x 64] bf16 as the projection GEMM writes it -> [B, H, T, 128]
 * (k | v of one key adjacent): decode_attention then reads one contiguous T x 256 B stream per (batch, head) with
 * kv_row_stride = 128, kv_head_stride = T * 128, kv_batch_stride = H * T * 128, V = K + 64. */
DICOW_API int dicow_kv_to_head_major(dicow_handle_t h, const void* kv_bf16, void* out_bf16, int B, int T, int H, void* stream);

/* x[b, s, :] = embed_tokens[ids[b, p + s]] + embed_positions[p + s], p = pos ? *pos : past   (fp32 tables, fp32 out)
 * HF:modeling_whisper.py:741-760 */
DICOW_API int dicow_embed_tokens(dicow_handle_t h, const int64_t* ids, int64_t ids_row_stride, const float* embed_tokens,
                                 const float* embed_positions, float* x, int B, int S, int d, int vocab, int past,
                                 const int32_t* pos, void* stream);
/* *pos += by (device scalar) */
DICOW_API int dicow_advance(dicow_handle_t h, int32_t* pos, int by, void* stream);

/* SuppressTokensLogitsProcessor + WhisperTimeStampLogitsProcessor (HF:generation/logits_process.py:1905-2043) + the DiCoW
 * EOS-at-begin exception (src/models/dicow/utils.py:5-14) + argmax + finished-row bookkeeping
 * (src/models/dicow/generation.py:728-779) in one pass over logits [B, V]; writes the new token to ids[b, len] where
 * len = pos ? *pos + 1 : cur_len, and clears unfinished[b] on EOS.  No host synchronisation. */
typedef struct {
  size_t struct_size;
  const float* logits;
  int64_t ld;
  int32_t B, V;
  int64_t* ids; /* [B, >= len + 1] */
  int64_t ids_row_stride;
  const int32_t* pos;
  int32_t cur_len;
  int32_t begin_index; /* prompt length (forced_decoder_ids) */
  int32_t eos, pad, no_timestamps, ts_begin;
  int32_t max_initial_timestamp_index; /* -1 = None */
  int32_t timestamp_rules;             /* 1 = return_timestamps (timestamp processor on); 0 = suppress list + argmax only */
  const uint32_t* suppress_bitmap;     /* ceil(V / 32) words, bit v set = suppressed; or NULL */
  int32_t* unfinished;                 /* [B] */
  float* processed_scores;             /* optional [B, V]: the scores after all processors (tests; input of the joint CTC step) */
  int32_t no_select;                   /* 1 = only write processed_scores: leave ids / unfinished alone (joint CTC decoding
                                          selects after the CTC rescoring, dicow_ctc_joint_step) */
} dicow_logits_rules_args_t;
DICOW_API int dicow_logits_rules_argmax(dicow_handle_t h, const dicow_logits_rules_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Joint CTC / attention decoding, one generated token for B hypotheses.  Replaces
 * LogSoftmaxProcessor + CTCRescorerLogitsProcessor.__call__ + argmax + CTCRescorerLogitsProcessor.update_state
 * (src/models/dicow/generation.py:250-268, 756-769; src/models/dicow/decoding.py:8-159, 253-338):
 *   scores  = log_softmax(processed_scores)                      (the suppress / timestamp processors ran before)
 *   cands   = top-K text ids of scores (EOS forced in)
 *   ctc[c]  = CTC prefix score of prefix + c for c in cands (forward variables over the T CTC frames), LOGZERO elsewhere,
 *             row maximum for timestamp ids
 *   token   = argmax (1 - w) scores + w (ctc - ctc_prev);  finished rows emit pad;  ids[b, len] = token
 *   a text token moves (r_prev, score_prev) of the hypothesis to its candidate's forward variables / prefix score.
 * ctc_logp: [B, T, V1] fp32 log-posteriors of the window (dicow_log_softmax_rows of the CTC logits; blank = V1 - 1 class).
 * workspace_i32: 4 B + 4 + B K ints; workspace_f32: B + 2 B K floats; states: B T 2 K floats; r_prev: [B, T, 2] initialised
 * to (LOGZERO, cumsum of the blank log-posteriors), score_prev: [B] zeros (decoding.py:37-44).  B <= 64, K <= 512.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  int64_t* ids; /* [B, >= len + 1] */
  int64_t ids_row_stride;
  const int32_t* pos; /* device scalar: len = *pos + 1 (or cur_len when NULL) */
  int32_t cur_len;
  int32_t B, V, T, V1, K;
  int32_t bos, eos, pad, blank, first_timestamp, prefix_len; /* prefix_len = len(tokenizer.prefix_tokens) */
  float ctc_weight;
  const float* ctc_logp;
  const float* processed_scores; /* [B, V] */
  int32_t* workspace_i32;
  float* workspace_f32;
  float* states;
  float* r_prev;
  float* score_prev;
  int32_t* unfinished;
  const float* raw_logits; /* optional [B, V]: the log-softmax normaliser is taken over these (beam search applies
                              log_softmax BEFORE the processors, generation.py:1003-1004) instead of over processed_scores */
  int32_t score_only;      /* 1: stop after the prefix scores (no selection / state update): beam search selects across
                              the beams of an utterance (dicow_beam_step) */
} dicow_ctc_joint_args_t;
DICOW_API int dicow_ctc_joint_step(dicow_handle_t h, const dicow_ctc_joint_args_t* args, void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Beam search, one step for U utterances x NB beams (rows r = u * NB + k), on the device.  Replaces the bookkeeping of
 * DiCoWGenerationMixin._beam_search (src/models/dicow/generation.py:992-1107 = HF GenerationMixin._beam_search with the
 * CTC hook): accumulated score = (1 - w) log-prob + w (ctc - ctc_prev) + running beam score over the candidates of every
 * beam (the top-K text ids scored by dicow_ctc_joint_step(score_only) + all timestamp ids), top-2NB continuations,
 * EOS / max_length hits, next running beams, finished set with length penalty, early-stop heuristic, then the re-linking
 * of token sequences, self-attention ancestry and CTC states to the chosen parents (update_state, generation.py:1087-1088).
 * flags[u] = {any continuation without a hit, all finished slots full, early-stop heuristic unsatisfied}: the caller ends
 * the loop when HF's _beam_search_has_unfinished_sequences over all utterances turns false.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  int32_t U, NB, V, K;
  const float* processed_scores; /* [U NB, V] */
  const int32_t* joint_workspace_i32; /* of dicow_ctc_joint_step(score_only): candidates */
  const float* joint_workspace_f32;   /* lse, attention log-probs and prefix scores of the candidates */
  float ctc_weight;                   /* 0: attention-only beam search */
  const float* ctc_states;            /* [U NB, T, 2, K] (ctc_weight > 0) */
  float* ctc_r_prev;                  /* [U NB, T, 2] */
  float* ctc_score_prev;              /* [U NB] */
  float* ctc_r_tmp;                   /* scratch like ctc_r_prev */
  int32_t T;
  float* run_score;  /* [U NB] running beam scores (first beam 0, others -1e9 at start) */
  float* fin_score;  /* [U NB] (-1e9 at start) */
  int32_t* fin_flag; /* [U NB] */
  int32_t* unsat;    /* [U] (1 at start) */
  int64_t* ids;      /* [U NB, ids_row_stride] running sequences */
  int64_t* fin_ids;  /* [U NB, ids_row_stride] finished sequences */
  int64_t* ids_tmp;  /* scratch [2 U NB, ids_row_stride] */
  int64_t ids_row_stride;
  int32_t* ancestry;     /* [U NB, ancestry_stride] */
  int32_t* ancestry_tmp; /* scratch */
  int64_t ancestry_stride;
  const int32_t* pos; /* device scalar: the sequences hold *pos + 1 tokens */
  int32_t eos, pad, first_timestamp, max_length, prompt_len;
  float length_penalty;
  int32_t early_stopping; /* 0 = False, 1 = True, 2 = "never" */
  int32_t* scratch_i32;   /* 3 U NB ints */
  float* scratch_f32;     /* 2 U NB floats */
  int32_t* flags;         /* [U, 4] */
} dicow_beam_step_args_t;
DICOW_API int dicow_beam_step(dicow_handle_t h, const dicow_beam_step_args_t* args, void* stream);

/* out[r, :] = in[r, :] - logsumexp(in[r, :]) over V columns (in place allowed) */
DICOW_API int dicow_log_softmax_rows(dicow_handle_t h, const float* in, float* out, int64_t rows, int V, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Losses.  soft-label CE: src/models/dicow/modeling_dicow.py:95-144 (soft_mode = 1: Gaussian-smoothed timestamp targets,
 * lower/upper-case streams, per-token min, mean over labels != -100) or :312-323 (soft_mode = 0: hard labels, mean over
 * all rows).  workspace: 2 * rows floats.  CTC: src/models/dicow/encoder.py:108-135 (blank = V1 - 1, every frame valid,
 * zero_infinity, reduction mean|sum); workspace: B * T + 2 * B floats.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  size_t struct_size;
  const float* logits; /* [rows, V] fp32 */
  int64_t ld;
  int32_t rows, V;
  const int64_t* labels;     /* [rows], -100 = padding */
  const int64_t* upp_labels; /* [rows] or NULL */
  int32_t ts_begin, n_ts;    /* n_ts = 0: no timestamp smoothing */
  const float* smoothing;    /* [n_ts, n_ts] row-normalised */
  int32_t soft_mode;
  float* workspace;
  float* loss; /* [1] */
} dicow_softlabel_ce_args_t;
DICOW_API int dicow_softlabel_ce(dicow_handle_t h, const dicow_softlabel_ce_args_t* args, void* stream);

typedef struct {
  size_t struct_size;
  const float* logits; /* [B, T, V1] fp32, rows `ld` elements apart */
  int32_t B, T, V1;
  const int64_t* labels; /* [B, Lmax], negative = padding */
  int32_t Lmax;
  int32_t reduction_mean; /* 1 = "mean" (per-utterance loss / target length, batch mean), 0 = "sum" */
  float* workspace;
  float* loss; /* [1] */
  int64_t ld;  /* row stride of logits in elements (0 = V1).  The CTC head writes its logits with a stride that is a multiple of
                  4 (V + 1 = 51 867 is odd): the GEMM epilogue then stores 16-byte vectors instead of 4-byte scalars */
} dicow_ctc_loss_args_t;
DICOW_API int dicow_ctc_loss(dicow_handle_t h, const dicow_ctc_loss_args_t* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * AdamW step over a list of fp32 tensors in ONE launch (the update rule of torch.optim.AdamW, which the reference's
 * get_optimizer builds: src/models/containers.py:100-114).  tensors: device array, one entry per parameter; chunks: device array of
 * n_chunks {tensor index, chunk index} int32 pairs covering every tensor in pieces of dicow_adamw_chunk_elems() elements.
 * Per tensor: p *= 1 - lr * weight_decay; m = lerp(m, g, 1 - beta1); v = beta2 v + (1 - beta2) g^2;
 *             p -= (lr / bias_correction1) * m / (sqrt(v) / bias_correction2_sqrt + eps).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  float* p;        /* parameter, updated in place */
  const float* g;  /* gradient */
  float* m;        /* exp_avg, updated in place */
  float* v;        /* exp_avg_sq, updated in place */
  int64_t n;       /* elements */
  float lr, weight_decay;
  float bias_correction1;      /* 1 - beta1^step */
  float bias_correction2_sqrt; /* sqrt(1 - beta2^step) */
} dicow_adamw_tensor_args_t;
DICOW_API int dicow_adamw_chunk_elems(void);
DICOW_API int dicow_adamw_step(dicow_handle_t h, const dicow_adamw_tensor_args_t* tensors, const int32_t* chunks, int n_chunks,
                               float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DICOW_B200_H */
