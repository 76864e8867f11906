// attention_common.h -- parameters shared by the attention kernels (attention_fa.cu: ping-pong kernel; attention.cu:
// C-ABI entry + the single-tile kernel kept for comparison).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

struct dicow_ctx;

namespace dicow {

struct AttnParams {
  int B, H, Tq, Tk, causal;
  __nv_bfloat16* out;
  long long o_rs, o_bs;
  long long* prof;  // debug: per-step clock64 stamps of one CTA (dicow_debug_set_attention_profile), else NULL
  float* lse;       // optional [B, H, Tq]: log2-sum-exp of every score row (training: attention backward), else NULL
};

// two 128-row query tiles per CTA, ping-pong softmax (attention_fa.cu); emu selects how many of every 8 exponentials run
// as a polynomial on the FMA pipe
int launch_attention_fa(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                        const AttnParams& p, int emu, cudaStream_t stream);

// single-pass backward (attention_bwd_fused.cu): non-causal, query blocks of 128; dq_acc is an fp32 [B, Tq, H * 64] workspace
struct FusedBwdArgs {
  int B, H, Tq, Tk;
  const float* lse;
  const float* D;
  float* dq_acc;
  __nv_bfloat16* dQ;
  __nv_bfloat16* dK;
  __nv_bfloat16* dV;
  long long dq_rs, dq_bs, dkv_rs, dkv_bs;
};
int launch_attention_bwd_fused(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmdO, const CUtensorMap& tmK,
                               const CUtensorMap& tmV, const FusedBwdArgs& a, cudaStream_t stream);

}  // namespace dicow
