// decode.cu -- HBM-bound kernels of the token-by-token decoder step (greedy generate()):
//   * skinny GEMM  out[M <= 64, N] = epilogue(A[M, K] W[N, K]^T): every weight byte is read exactly once from HBM and
//     multiplied against all M rows (the batch) with warp-level mma.sync m16n8k16 -- at M = 16 the CUDA-core FMA rate
//     would cap the kernel below the HBM roofline, and a 128-row tcgen05 tile would leave most SMs without work
//     (N / 256 tiles) while re-reading W once per batch row group;
//   * decode attention: one query row per (batch, head) against a K/V cache (self: Tk = pos + 1; cross: Tk = 1500);
//   * token + position embedding gather;
//   * fused logits rules + argmax: SuppressTokensLogitsProcessor + WhisperTimeStampLogitsProcessor + the DiCoW
//     EOS-at-begin exception + argmax + finished-row handling in one pass over the [B, V] logits, no host sync.
// Every kernel can take the current position from a device scalar, so a whole decode step is a fixed sequence of
// launches that is captured once in a CUDA graph and replayed per token.
//
// Replaces, per generated token, HF WhisperDecoder.forward with a KV cache (HF:models/whisper/modeling_whisper.py:
// 449-506, 691-796), proj_out (src/models/dicow/modeling_dicow.py:302) and the greedy branch of
// DiCoWGenerationMixin._sample (src/models/dicow/generation.py:707-782) with its logits processors
// (HF:generation/logits_process.py:1905-2043; src/models/dicow/utils.py:5-14).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// skinny GEMM
// ------------------------------------------------------------------------------------------------------------------
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;

struct SkinnyParams {
  const __nv_bfloat16* A;
  long long lda;
  const __nv_bfloat16* W;
  long long ldw;
  int M, N, K;
  const float* bias;
  void* out;
  long long ldo;
  int epilogue;
  const float* resid;
  long long ldr;
  const int* pos;
  long long pos_stride;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {  // weights: read once, do not pollute L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// One CTA = 8 output columns; its 8 warps split K in 32-element blocks (warp w takes blocks w, w + 8, ...).  Lane
// (g = lane / 4, t = lane % 4) loads 16 contiguous bytes of W row n0 + g at k = 32 kb + 8 t and the same bytes of A rows
// g and g + 8 of every 16-row tile; the 8 k values feed two m16n8k16 MMAs with a k permutation that is identical on the
// A and B side (so the contraction is unchanged).  Partial sums meet in shared memory.
template <int MT>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const SkinnyParams p) {
  __shared__ float red[SK_WARPS][MT * 16][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8;
  const int nrow = n0 + g;
  const bool n_ok = nrow < p.N;
  const __nv_bfloat16* wrow = p.W + (long long)(n_ok ? nrow : 0) * p.ldw + t * 8;
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
  const int kblocks = p.K >> 5;
#pragma unroll 4
  for (int kb = warp; kb < kblocks; kb += SK_WARPS) {
    uint4 wv = make_uint4(0u, 0u, 0u, 0u);
    if (n_ok) wv = ldg_stream_128(wrow + kb * 32);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int r0 = m * 16 + g, r1 = r0 + 8;
      uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
      if (r0 < p.M) lo = __ldg(reinterpret_cast<const uint4*>(p.A + (long long)r0 * p.lda + kb * 32 + t * 8));
      if (r1 < p.M) hi = __ldg(reinterpret_cast<const uint4*>(p.A + (long long)r1 * p.lda + kb * 32 + t * 8));
      mma_bf16_16816(acc[m], lo.x, hi.x, lo.y, hi.y, wv.x, wv.y);
      mma_bf16_16816(acc[m], lo.z, hi.z, lo.w, hi.w, wv.z, wv.w);
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    red[warp][m * 16 + g][2 * t] = acc[m][0];
    red[warp][m * 16 + g][2 * t + 1] = acc[m][1];
    red[warp][m * 16 + g + 8][2 * t] = acc[m][2];
    red[warp][m * 16 + g + 8][2 * t + 1] = acc[m][3];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MT * 16 * 8; i += SK_THREADS) {
    const int row = i >> 3, col = i & 7;
    const int n = n0 + col;
    if (row >= p.M || n >= p.N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += red[w][row][col];
    if (p.bias != nullptr) v += __ldg(p.bias + n);
    long long o = (long long)row * p.ldo + n;
    if (p.pos != nullptr) o += (long long)(*p.pos) * p.pos_stride;
    switch (p.epilogue) {
      case DICOW_EPI_BIAS_BF16: reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v); break;
      case DICOW_EPI_BIAS_GELU_BF16:
        reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(gelu_erf_fast(v));
        break;
      case DICOW_EPI_RESIDUAL_F32:
        reinterpret_cast<float*>(p.out)[o] = p.resid[(long long)row * p.ldr + n] + v;
        break;
      default: reinterpret_cast<float*>(p.out)[o] = v; break;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// decode attention: softmax(q . K^T) V for one query row per (batch, head), head_dim 64
// ------------------------------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 4;

struct DecAttnParams {
  const __nv_bfloat16* Q;
  long long q_bs;
  const __nv_bfloat16* K;
  const __nv_bfloat16* V;
  long long kv_rs, kv_bs;
  __nv_bfloat16* out;
  long long o_bs;
  int Tk;
  const int* pos;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x, f[2 * i + 1] = v.y;
  }
}

// grid (H, B), 4 warps.  A group of 8 lanes owns one key at a time (16 B = 8 dims per lane, so every K / V row is
// read as one full 128-byte line); each group keeps an online-softmax state (m, l, o[8 dims per lane]); groups and
// warps are merged at the end.
__global__ void __launch_bounds__(DA_WARPS * 32) decode_attention_kernel(const DecAttnParams p) {
  __shared__ float sm_m[DA_WARPS], sm_l[DA_WARPS], sm_o[DA_WARPS][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int Tk = p.pos != nullptr ? (*p.pos + 1) : p.Tk;
  float q[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(p.Q + (long long)b * p.q_bs + h * 64 + sub * 8)), q);
  const __nv_bfloat16* kb = p.K + (long long)b * p.kv_bs + h * 64 + sub * 8;
  const __nv_bfloat16* vb = p.V + (long long)b * p.kv_bs + h * 64 + sub * 8;
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  constexpr int UNROLL = 4;
  const int stride = DA_WARPS * 4;
  for (int kw = warp * 4; kw < Tk; kw += stride * UNROLL) {  // warp-uniform trip count (shuffles below)
    const int k0 = kw + grp;
    uint4 kv[UNROLL], vv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int k = k0 + u * stride;
      if (k < Tk) {
        kv[u] = ldg_stream_128(kb + (long long)k * p.kv_rs);
        vv[u] = ldg_stream_128(vb + (long long)k * p.kv_rs);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int k = k0 + u * stride;
      const bool ok = k < Tk;  // uniform within the 8-lane group
      float kf[8];
      float s = 0.f;
      if (ok) {
        unpack8(kv[u], kf);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok) {
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn);  // 0 for the first key (m = -inf)
        const float pw = __expf(s - mn);
        float vf[8];
        unpack8(vv[u], vf);
        l = fmaf(l, corr, pw);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(o[i], corr, pw * vf[i]);
        m = mn;
      }
    }
  }
  // merge the 4 groups of the warp (lanes with the same `sub`)
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, x);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, x);
    const float mn = fmaxf(m, m2);
    const float c1 = (m == -INFINITY) ? 0.f : __expf(m - mn);
    const float c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[i], x);
      o[i] = o[i] * c1 + o2 * c2;
    }
    m = mn;
  }
  if (grp == 0) {
    if (sub == 0) sm_m[warp] = m, sm_l[warp] = l;
#pragma unroll
    for (int i = 0; i < 8; ++i) sm_o[warp][sub * 8 + i] = o[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float mt = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) mt = fmaxf(mt, sm_m[w]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float c = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - mt);
      lt = fmaf(sm_l[w], c, lt);
      ot = fmaf(sm_o[w][threadIdx.x], c, ot);
    }
    p.out[(long long)b * p.o_bs + h * 64 + threadIdx.x] = __float2bfloat16_rn(ot / lt);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// x[b, s, :] = embed_tokens[ids[b, past + s], :] + embed_positions[past + s, :]      (fp32)
// ------------------------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const long long* __restrict__ ids, long long ids_rs, const float* __restrict__ tok,
                             const float* __restrict__ posw, float* __restrict__ x, int S, int d, int past,
                             const int* pos, int vocab) {
  const int s = blockIdx.x, b = blockIdx.y;
  const int pp = (pos != nullptr ? *pos : past) + s;
  long long id = ids[(long long)b * ids_rs + pp];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* tr = reinterpret_cast<const float4*>(tok + id * d);
  const float4* pr = reinterpret_cast<const float4*>(posw + (long long)pp * d);
  float4* xr = reinterpret_cast<float4*>(x + ((long long)b * S + s) * d);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
    const float4 a = __ldg(tr + i), c = __ldg(pr + i);
    xr[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

__global__ void advance_kernel(int* pos, int by) { *pos += by; }

// ------------------------------------------------------------------------------------------------------------------
// logits rules + argmax
// ------------------------------------------------------------------------------------------------------------------
struct RulesParams {
  const float* logits;
  long long ld;
  int V;
  long long* ids;
  long long ids_rs;
  const int* pos;  // column of the token that produced these logits; the new token goes to column *pos + 1
  int cur_len;     // used when pos == NULL: number of tokens already in ids
  int begin_index, eos, pad, no_timestamps, ts_begin, max_initial_ts, ts_rules;
  const unsigned* suppress;  // bitmap, ceil(V / 32) words, or NULL
  int* unfinished;           // [B] 1 = still decoding
  float* processed;          // optional [B, V] fp32 copy of the processed scores (tests)
};

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ Best best_of(Best a, Best b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = best_of(x, y);
  }
  return x;
}

// One CTA per batch row.  Restates, in one pass over the row:
//   SuppressTokensLogitsProcessor -> scores[suppress] = -inf
//   WhisperTimeStampLogitsProcessor (HF:generation/logits_process.py:1996-2043): no_timestamps = -inf; timestamps come
//   in pairs; no decreasing timestamps; first token must be a timestamp; "if the probability mass over timestamps
//   exceeds the most likely text token, sample a timestamp"
//   DiCoW: the EOS score survives at the first generated position (src/models/dicow/utils.py:10-12)
//   argmax (lowest index on ties), finished rows emit pad, unfinished &= token != eos (generation.py:756-779)
__global__ void __launch_bounds__(512) logits_rules_argmax_kernel(const RulesParams p) {
  __shared__ float s_max[16], s_sum[16];
  __shared__ Best s_text[16], s_ts[16];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int len = p.pos != nullptr ? (*p.pos + 1) : p.cur_len;  // tokens in the sequence so far
  long long* row_ids = p.ids + (long long)b * p.ids_rs;
  const int ngen = len - p.begin_index;
  const bool rules = p.ts_rules != 0;
  const bool at_begin = rules && ngen == 0;
  const long long last = ngen >= 1 ? row_ids[len - 1] : -1;
  const long long penult = ngen >= 2 ? row_ids[len - 2] : -1;
  const bool last_was_ts = rules && ngen >= 1 && last >= p.ts_begin;
  const bool penult_was_ts = ngen < 2 || penult >= p.ts_begin;
  // last timestamp token of the generated part (timestamps never decrease, so it is the maximum)
  int ts_last = -1;
  for (int i = len - 1; rules && i >= p.begin_index; --i) {
    const long long tk = row_ids[i];
    if (tk >= p.ts_begin) {
      ts_last = (int)tk;
      break;
    }
  }
  int ts_floor = p.ts_begin;  // timestamp ids below this are forbidden
  if (ts_last >= 0) ts_floor = (last_was_ts && !penult_was_ts) ? ts_last : ts_last + 1;
  const bool ts_all_masked = last_was_ts && penult_was_ts;
  const bool text_lt_eos_masked = last_was_ts && !penult_was_ts;
  const int ts_cap = (at_begin && p.max_initial_ts >= 0) ? p.ts_begin + p.max_initial_ts : p.V;  // ids > cap forbidden

  const float* lg = p.logits + (long long)b * p.ld;
  float* proc = p.processed != nullptr ? p.processed + (long long)b * p.V : nullptr;
  Best bt{-INFINITY, p.V}, bs{-INFINITY, p.V};
  float tmax = -INFINITY, tsum = 0.f;  // online logsumexp over the timestamp region
  for (int v = tid; v < p.V; v += blockDim.x) {
    float x = lg[v];
    bool masked = (p.suppress != nullptr && ((p.suppress[v >> 5] >> (v & 31)) & 1u)) || (rules && v == p.no_timestamps);
    if (v < p.ts_begin) {
      masked = masked || at_begin || (text_lt_eos_masked && v < p.eos);
      if (masked) x = -INFINITY;
      bt = best_of(bt, Best{x, v});
    } else {
      masked = masked || ts_all_masked || v < ts_floor || v > ts_cap;
      if (masked) x = -INFINITY;
      bs = best_of(bs, Best{x, v});
      if (x > -INFINITY) {
        const float mn = fmaxf(tmax, x);
        tsum = tsum * __expf(tmax - mn) + __expf(x - mn);
        tmax = mn;
      }
    }
    if (proc != nullptr) proc[v] = x;
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, tmax, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, tsum, o);
    const float mn = fmaxf(tmax, m2);
    const float c1 = tmax == -INFINITY ? 0.f : __expf(tmax - mn);
    const float c2 = m2 == -INFINITY ? 0.f : __expf(m2 - mn);
    tsum = tsum * c1 + s2 * c2;
    tmax = mn;
  }
  if (lane == 0) s_text[warp] = bt, s_ts[warp] = bs, s_max[warp] = tmax, s_sum[warp] = tsum;
  __syncthreads();
  __shared__ int s_text_off;
  if (tid == 0) {
    for (int w = 1; w < nw; ++w) {
      bt = best_of(bt, s_text[w]);
      bs = best_of(bs, s_ts[w]);
      const float mn = fmaxf(tmax, s_max[w]);
      const float c1 = tmax == -INFINITY ? 0.f : __expf(tmax - mn);
      const float c2 = s_max[w] == -INFINITY ? 0.f : __expf(s_max[w] - mn);
      tsum = tsum * c1 + s_sum[w] * c2;
      tmax = mn;
    }
    // logsumexp(timestamp log-probs) > max(text log-probs)  <=>  lse(ts scores) > max(text scores)
    const float ts_lse = tsum > 0.f ? tmax + __logf(tsum) : -INFINITY;
    const bool text_off = rules && ts_lse > bt.v;
    Best win = text_off ? bs : best_of(bt, bs);
    if (at_begin) {  // DiCoW: EOS keeps the score it had before the timestamp rules
      float e = lg[p.eos];
      if ((p.suppress != nullptr && ((p.suppress[p.eos >> 5] >> (p.eos & 31)) & 1u)) || p.eos == p.no_timestamps)
        e = -INFINITY;
      win = best_of(win, Best{e, p.eos});
    }
    if (win.i >= p.V) win.i = p.eos;  // every score -inf: cannot happen with finite logits; stay defined
    long long tok = win.i;
    const int unf = p.unfinished[b];
    if (!unf) tok = p.pad;
    row_ids[len] = tok;
    p.unfinished[b] = unf && (tok != p.eos);
    s_text_off = text_off ? 1 : 0;
  }
  if (proc != nullptr) {  // materialise the processed scores exactly like the reference processors (tests only)
    __syncthreads();
    if (s_text_off)
      for (int v = tid; v < p.ts_begin; v += blockDim.x) proc[v] = -INFINITY;
    __syncthreads();
    if (tid == 0 && at_begin) {
      float e = lg[p.eos];
      if ((p.suppress != nullptr && ((p.suppress[p.eos >> 5] >> (p.eos & 31)) & 1u)) || p.eos == p.no_timestamps)
        e = -INFINITY;
      proc[p.eos] = e;
    }
  }
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_gemm_skinny_bf16(dicow_handle_t h, const dicow_gemm_skinny_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_gemm_skinny_args_t), "dicow_gemm_skinny_bf16: bad args struct");
  DICOW_REQUIRE(ctx, a->A && a->W && a->out, "dicow_gemm_skinny_bf16: null operand");
  DICOW_REQUIRE(ctx, a->M >= 1 && a->M <= 64 && a->N >= 1 && a->K >= 32 && (a->K % 32) == 0,
                "dicow_gemm_skinny_bf16: need 1 <= M <= 64, K %% 32 == 0 (got M=%d N=%d K=%d)", a->M, a->N, a->K);
  DICOW_REQUIRE(ctx, (a->lda % 8) == 0 && (a->ldw % 8) == 0 && (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->W) % 16) == 0,
                "dicow_gemm_skinny_bf16: A / W rows must be 16-byte aligned");
  DICOW_REQUIRE(ctx, a->epilogue >= DICOW_EPI_BIAS_BF16 && a->epilogue <= DICOW_EPI_BIAS_F32,
                "dicow_gemm_skinny_bf16: unsupported epilogue %d", a->epilogue);
  if (a->epilogue == DICOW_EPI_RESIDUAL_F32) DICOW_REQUIRE(ctx, a->resid != nullptr, "dicow_gemm_skinny_bf16: resid is NULL");
  SkinnyParams p{};
  p.A = reinterpret_cast<const __nv_bfloat16*>(a->A), p.lda = a->lda;
  p.W = reinterpret_cast<const __nv_bfloat16*>(a->W), p.ldw = a->ldw;
  p.M = a->M, p.N = a->N, p.K = a->K, p.bias = a->bias, p.out = a->out, p.ldo = a->ldo, p.epilogue = a->epilogue;
  p.resid = a->resid, p.ldr = a->ldr, p.pos = a->pos, p.pos_stride = a->pos_stride;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = ceil_div(a->N, 8);
  switch (ceil_div(a->M, 16)) {
    case 1: gemm_skinny_kernel<1><<<grid, SK_THREADS, 0, stream>>>(p); break;
    case 2: gemm_skinny_kernel<2><<<grid, SK_THREADS, 0, stream>>>(p); break;
    case 3: gemm_skinny_kernel<3><<<grid, SK_THREADS, 0, stream>>>(p); break;
    default: gemm_skinny_kernel<4><<<grid, SK_THREADS, 0, stream>>>(p); break;
  }
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_decode_attention_bf16(dicow_handle_t h, const dicow_decode_attention_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_decode_attention_args_t),
                "dicow_decode_attention_bf16: bad args struct");
  DICOW_REQUIRE(ctx, a->Q && a->K && a->V && a->out, "dicow_decode_attention_bf16: null operand");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->B <= 65535 && a->H >= 1 && (a->Tk >= 1 || a->pos != nullptr),
                "dicow_decode_attention_bf16: bad shape B=%d H=%d Tk=%d", a->B, a->H, a->Tk);
  DICOW_REQUIRE(ctx, (a->q_batch_stride % 8) == 0 && (a->kv_row_stride % 8) == 0 && (a->kv_batch_stride % 8) == 0 &&
                         ((reinterpret_cast<uintptr_t>(a->Q) | reinterpret_cast<uintptr_t>(a->K) |
                           reinterpret_cast<uintptr_t>(a->V)) % 16) == 0,
                "dicow_decode_attention_bf16: strides must be multiples of 8 elements, bases 16-byte aligned");
  DecAttnParams p{};
  p.Q = reinterpret_cast<const __nv_bfloat16*>(a->Q), p.q_bs = a->q_batch_stride;
  p.K = reinterpret_cast<const __nv_bfloat16*>(a->K), p.V = reinterpret_cast<const __nv_bfloat16*>(a->V);
  p.kv_rs = a->kv_row_stride, p.kv_bs = a->kv_batch_stride;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out), p.o_bs = a->o_batch_stride, p.Tk = a->Tk, p.pos = a->pos;
  dim3 grid(a->H, a->B);
  decode_attention_kernel<<<grid, DA_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_embed_tokens(dicow_handle_t h, const int64_t* ids, int64_t ids_row_stride, const float* embed_tokens,
                                  const float* embed_positions, float* x, int B, int S, int d, int vocab, int past,
                                  const int32_t* pos, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, ids && embed_tokens && embed_positions && x && B >= 1 && B <= 65535 && S >= 1 && d >= 4 && (d % 4) == 0,
                "dicow_embed_tokens: bad args");
  dim3 grid(S, B);
  embed_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const long long*>(ids), ids_row_stride, embed_tokens, embed_positions, x, S, d, past, pos, vocab);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_advance(dicow_handle_t h, int32_t* pos, int by, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, pos != nullptr, "dicow_advance: null pos");
  advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(pos, by);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_logits_rules_argmax(dicow_handle_t h, const dicow_logits_rules_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_logits_rules_args_t),
                "dicow_logits_rules_argmax: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->ids && a->unfinished && a->B >= 1 && a->V >= 2, "dicow_logits_rules_argmax: bad args");
  DICOW_REQUIRE(ctx, a->eos >= 0 && a->eos < a->V && a->ts_begin > a->eos && a->ts_begin <= a->V && a->begin_index >= 0,
                "dicow_logits_rules_argmax: need 0 <= eos < ts_begin <= V");
  RulesParams p{};
  p.logits = a->logits, p.ld = a->ld, p.V = a->V;
  p.ids = reinterpret_cast<long long*>(a->ids), p.ids_rs = a->ids_row_stride, p.pos = a->pos, p.cur_len = a->cur_len;
  p.begin_index = a->begin_index, p.eos = a->eos, p.pad = a->pad, p.no_timestamps = a->no_timestamps;
  p.ts_begin = a->ts_begin, p.max_initial_ts = a->max_initial_timestamp_index, p.ts_rules = a->timestamp_rules;
  p.suppress = a->suppress_bitmap, p.unfinished = a->unfinished, p.processed = a->processed_scores;
  logits_rules_argmax_kernel<<<a->B, 512, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
