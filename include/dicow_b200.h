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
  DICOW_EPI_GELU_FDDT_POS_F32 = 4 /* out_f32 = FDDT(gelu_erf(acc + bias), stno) + pos[m, :]   (conv2 epilogue:
                                     src/models/dicow/encoder.py:168-179)                                        */
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
  const float* fddt_w; /* [4, N] fp32, rows in S,T,N,O order */
  const float* fddt_b; /* [4, N] */
  const float* pos;    /* [Mb, N] fp32 (embed_positions.weight) or NULL */
} dicow_gemm_args_t;

DICOW_API int dicow_gemm_bf16(dicow_handle_t h, const dicow_gemm_args_t* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DICOW_B200_H */
