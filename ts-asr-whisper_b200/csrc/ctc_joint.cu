// ctc_joint.cu -- joint CTC / attention decoding on the device.
//
// Replaces, per generated token, CTCRescorerLogitsProcessor.__call__ / update_state and CTCPrefixScore.__call__
// (src/models/dicow/decoding.py:8-159, 166-338), which the reference runs as ~40 eager tensor ops plus a Python loop over
// the T' = 375 CTC frames, with a dense [hypotheses, V, T', 2] state table (155 MB per hypothesis at Whisper's vocabulary):
//
//   ctc_prepare_kernel        prefix bookkeeping on the token ids of every hypothesis (decoding.py:263-291): decoded
//                             length, last label (with the reference's count-indexed replacement of a trailing timestamp),
//                             "still to be decoded", and the first frame of the forward recursion (the minimum over the
//                             scored hypotheses, decoding.py:98-100)
//   ctc_topk_kernel           log-softmax normaliser of the processed attention scores + the top-k text candidates
//                             (decoding.py:293-297) by a 4-pass radix select on the order-preserving integer image of the
//                             scores; EOS is forced into the set
//   ctc_prefix_score_kernel   forward variables r^n, r^b and prefix scores log psi of the k candidates of one hypothesis:
//                             one thread per candidate, the T' frames in sequence, log phi / blank posteriors of the
//                             hypothesis staged in shared memory (decoding.py:57-119); states are kept for the k
//                             candidates only ([T', 2, k] per hypothesis)
//   ctc_combine_select_kernel (1 - w) attention + w (ctc - ctc_prev) over candidates and timestamp ids (timestamps take the
//                             row maximum of the CTC scores, decoding.py:325-328), argmax, finished-row bookkeeping
//                             (generation.py:756-779) and update_state (decoding.py:253-260) in one pass
//   log_softmax_rows_kernel   CTC posteriors of the window, once per window (decoding.py:185)
//
// LOGZERO = -1e10 enters logaddexp as a number, exactly as in the reference.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr float CTC_LOGZERO = -1e10f;

__device__ __forceinline__ float lae(float a, float b) {  // torch.logaddexp for finite inputs
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

__device__ __forceinline__ unsigned ordered_key(float x) {  // larger float <=> larger unsigned; -inf is the smallest
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct CtcJointParams {
  const long long* ids;  // [B, ids_rs] token ids (prompt + generated)
  long long ids_rs;
  const int* pos;        // device scalar: the sequence holds *pos + 1 tokens (or cur_len when NULL)
  int cur_len;
  int B, V, T, V1, K;
  int bos, eos, pad, blank, first_ts, prefix_len;
  float w;
  const float* x;          // [B, T, V1] CTC log-posteriors of the window
  const float* proc;       // [B, V] attention scores after the other logits processors (-inf = masked)
  const float* raw;        // optional [B, V]: rows the log-softmax normaliser is taken over (NULL: proc)
  int* meta;               // [B, 4]: decoded_len, last label, to_be_decoded, unused;  meta[4 * B] = loop start
  float* lse;              // [B] log-sum-exp of the processed attention scores
  int* cs;                 // [B, K] candidate ids
  float* att;              // [B, K] attention log-probs of the candidates
  float* psi;              // [B, K] CTC prefix scores of the candidates
  float* states;           // [B, T, 2, K] forward variables of the candidates
  float* r_prev;           // [B, T, 2] forward variables of the hypotheses
  float* score_prev;       // [B]
  int* unfinished;         // [B]
  long long* ids_out;      // == ids (written at column len)
};

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) ctc_prepare_kernel(const CtcJointParams p) {
  __shared__ int s_start[64];
  const int b = threadIdx.x;
  int start = 0x7fffffff;
  if (b < p.B) {
    const int len = p.pos != nullptr ? (*p.pos + 1) : p.cur_len;
    const long long* row = p.ids + (long long)b * p.ids_rs;
    int first = 0;  // decoding.py:266-268: drop everything before the decoder start token
    if (row[0] != p.bos)
      for (int i = 0; i < len; ++i)
        if (row[i] == p.bos) {
          first = i;
          break;
        }
    const int g0 = first + (p.prefix_len > 1 ? p.prefix_len - 1 : 0) + 1;  // first generated label (slot 0 = sos -> blank)
    const int ngen = len - g0 > 0 ? len - g0 : 0;
    int decoded_len = 0, n_text = 1;
    for (int i = 0; i < ngen; ++i) {
      const long long v = row[g0 + i];
      decoded_len += (v <= p.first_ts && v != p.blank) ? 1 : 0;
      n_text += (v < p.first_ts || v == p.blank) ? 1 : 0;
    }
    long long last = ngen > 0 ? row[g0 + ngen - 1] : (long long)p.blank;
    if (last >= p.first_ts && last != p.blank) last = (n_text - 1 == 0) ? (long long)p.blank : row[g0 + n_text - 2];
    const int todo = last != p.eos ? 1 : 0;
    p.meta[4 * b + 0] = decoded_len, p.meta[4 * b + 1] = (int)last, p.meta[4 * b + 2] = todo, p.meta[4 * b + 3] = len;
    if (todo) start = decoded_len > 1 ? decoded_len : 1;
  }
  s_start[threadIdx.x] = start;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0x7fffffff;
    for (int i = 0; i < 64; ++i) m = min(m, s_start[i]);
    p.meta[4 * p.B] = m == 0x7fffffff ? 1 : m;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// one CTA per hypothesis: lse over the processed row, then radix select of the K largest text scores
// ------------------------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 1024;

__global__ void __launch_bounds__(TK_THREADS) ctc_topk_kernel(const CtcJointParams p) {
  __shared__ float s_m[32], s_s[32];
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_need, s_count_gt, s_count_eq;
  __shared__ int s_has_eos;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* row = p.proc + (long long)b * p.V;
  const float* nrow = p.raw != nullptr ? p.raw + (long long)b * p.V : row;
  // ---- log-sum-exp of the whole row (greedy: LogSoftmaxProcessor after the other processors, generation.py:252;
  // ---- beam search: log_softmax of the raw logits before them, generation.py:1003) ----
  float m = -INFINITY, s = 0.f;
  for (int v = tid; v < p.V; v += TK_THREADS) {
    const float x = nrow[v];
    if (x > -INFINITY) {
      const float mn = fmaxf(m, x);
      s = s * __expf(m - mn) + __expf(x - mn);
      m = mn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) s_m[warp] = m, s_s[warp] = s;
  if (tid == 0) s_prefix = 0u, s_need = (unsigned)p.K, s_count_gt = 0u, s_count_eq = 0u, s_has_eos = 0;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < TK_THREADS / 32; ++w) {
      const float mn = fmaxf(m, s_m[w]);
      s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (s_m[w] == -INFINITY ? 0.f : s_s[w] * __expf(s_m[w] - mn));
      m = mn;
    }
    p.lse[b] = s > 0.f ? m + logf(s) : -INFINITY;
  }
  // ---- radix select: the K-th largest key among v < first_ts, most significant byte first ----
  const int n = p.first_ts;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += TK_THREADS) hist[i] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix, pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int v = tid; v < n; v += TK_THREADS) {
      const unsigned k = ordered_key(row[v]);
      if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_need, acc = 0u;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (acc + hist[bin] >= need) break;
        acc += hist[bin];
      }
      s_need = need - acc;  // how many of the chosen bin are still needed
      s_prefix = prefix | ((unsigned)bin << shift);
    }
    __syncthreads();
  }
  const unsigned kth = s_prefix;  // key of the K-th largest; s_need = how many entries equal to it belong to the set
  const unsigned need_eq = s_need;
  int* cs = p.cs + (long long)b * p.K;
  float* att = p.att + (long long)b * p.K;
  const float lse = p.lse[b];
  __syncthreads();
  // entries above the threshold, in any order
  for (int v = tid; v < n; v += TK_THREADS) {
    const float x = row[v];
    const unsigned k = ordered_key(x);
    if (k > kth) {
      const unsigned slot = atomicAdd(&s_count_gt, 1u);
      cs[slot] = v, att[slot] = x - lse;
      if (v == p.eos) s_has_eos = 1;
    }
  }
  __syncthreads();
  const unsigned base = s_count_gt;
  // entries equal to the threshold fill the remaining need_eq slots (which of several equal scores get in is as
  // unspecified as in torch.topk; with continuous scores there is exactly one).  All threads scan: a single thread walking
  // the 50 365 text ids cost 1.5-2 ms per step.
  for (int v = tid; v < n; v += TK_THREADS) {
    const float x = row[v];
    if (ordered_key(x) == kth) {
      const unsigned e = atomicAdd(&s_count_eq, 1u);
      if (e < need_eq) {
        cs[base + e] = v, att[base + e] = x - lse;
        if (v == p.eos) s_has_eos = 1;
      }
    }
  }
  __syncthreads();
  // decoding.py:296-297: EOS is always scored; it takes the place of the weakest candidate
  if (tid == 0 && !s_has_eos && p.eos < n) {
    const unsigned slot = base + need_eq - 1;
    cs[slot] = p.eos, att[slot] = row[p.eos] - lse;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward variables and prefix scores: one CTA per hypothesis, one thread per candidate
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ctc_prefix_score_kernel(const CtcJointParams p) {
  extern __shared__ float cps_smem[];  // r_sum [T] | r_prev blank branch [T] | blank posterior [T]
  float* r_sum = cps_smem;
  float* r_pb = cps_smem + p.T;
  float* xb = cps_smem + 2 * p.T;
  const int b = blockIdx.x, j = blockIdx.y * blockDim.x + threadIdx.x;  // candidates of a hypothesis spread over CTAs
  const int decoded_len = p.meta[4 * b + 0], last = p.meta[4 * b + 1], todo = p.meta[4 * b + 2];
  const int loop_start = p.meta[4 * p.B];
  float* psi_out = p.psi + (long long)b * p.K;
  if (!todo) {  // finished hypotheses are not scored (decoding.py:288-291): every CTC score stays LOGZERO
    if (j < p.K) psi_out[j] = CTC_LOGZERO;
    return;
  }
  const float* xrow = p.x + (long long)b * p.T * p.V1;
  const float* rp = p.r_prev + (long long)b * p.T * 2;
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
    const float rn = rp[2 * t], rb = rp[2 * t + 1];
    r_sum[t] = lae(rn, rb);
    r_pb[t] = rb;
    xb[t] = xrow[(long long)t * p.V1 + p.blank];
  }
  __syncthreads();
  if (j >= p.K) return;
  const int c = p.cs[(long long)b * p.K + j];
  const float* xc = xrow + c;
  const float* phi = (decoded_len > 0 && c == last) ? r_pb : r_sum;
  float* st = p.states + (long long)b * p.T * 2 * p.K + j;  // st[(t * 2 + s) * K]
  const long long sK = p.K;
  // frame 0; r[start - 1, 0] of log psi is frame 0 when the prefix is empty or one label long, a LOGZERO row otherwise
  float rn = decoded_len == 0 ? __ldg(xc) : CTC_LOGZERO, rbk = CTC_LOGZERO;
  st[0] = rn, st[sK] = rbk;
  float acc = (decoded_len <= 1) ? rn : CTC_LOGZERO;
  if (loop_start > 1) rn = CTC_LOGZERO;  // r[loop_start - 1] is an untouched LOGZERO row
  // One pass over the frames: the running logsumexp of log psi (decoding.py:84-95) and the forward recursion
  // (decoding.py:98-107).  The candidate's posteriors x[t, c] are 207 KB apart: 16 frames are requested at a time so
  // that the dependent recursion waits for one memory round trip per 16 frames, not per frame.
  constexpr int PF = 16;
  float lm = -INFINITY, ls = 0.f;
  for (int t0 = 1; t0 < p.T; t0 += PF) {
    float xv[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) xv[u] = (t0 + u < p.T) ? __ldg(xc + (long long)(t0 + u) * p.V1) : 0.f;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int t = t0 + u;
      if (t < p.T) {
        const float ph = phi[t - 1];
        const float term = t >= decoded_len ? ph + xv[u] : CTC_LOGZERO;
        const float mn = fmaxf(lm, term);
        ls = ls * expf(lm - mn) + expf(term - mn);
        lm = mn;
        if (t >= loop_start) {
          const float nn = lae(rn, ph) + xv[u];
          const float nb = lae(rn, rbk) + xb[t];
          rn = nn, rbk = nb;
          st[(2LL * t) * sK] = rn, st[(2LL * t + 1) * sK] = rbk;
        } else {
          st[(2LL * t) * sK] = CTC_LOGZERO, st[(2LL * t + 1) * sK] = CTC_LOGZERO;
        }
      }
    }
  }
  if (p.T > 1) acc = lae(acc, lm + logf(ls));
  if (c == p.eos) acc = r_sum[p.T - 1];
  else if (c == p.blank) acc = CTC_LOGZERO;
  psi_out[j] = acc;
}

// ------------------------------------------------------------------------------------------------------------------
// combine, select, bookkeeping, update_state: one CTA per hypothesis
// ------------------------------------------------------------------------------------------------------------------
struct Pick {
  float v;
  int i;
  int slot;  // candidate slot, -1 for a timestamp id
};
__device__ __forceinline__ Pick better(Pick a, Pick b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__global__ void __launch_bounds__(512) ctc_combine_select_kernel(const CtcJointParams p) {
  __shared__ float s_f[16];
  __shared__ Pick s_p[16];
  __shared__ Pick s_win;
  __shared__ float s_maxpsi;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int todo = p.meta[4 * b + 2], len = p.meta[4 * b + 3];
  const float* psi = p.psi + (long long)b * p.K;
  const float* att = p.att + (long long)b * p.K;
  const int* cs = p.cs + (long long)b * p.K;
  const float prev = p.score_prev[b];
  // row maximum of the CTC scores: scored candidates, LOGZERO everywhere else (decoding.py:325)
  float mx = CTC_LOGZERO;
  if (todo)
    for (int j = tid; j < p.K; j += blockDim.x) mx = fmaxf(mx, psi[j]);
  mx = warp_max(mx);
  if (lane == 0) s_f[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_f[w]);
    s_maxpsi = mx;
  }
  __syncthreads();
  const float maxpsi = s_maxpsi;
  const float lse = p.lse[b];
  Pick best{-INFINITY, p.V, -1};
  if (todo)
    for (int j = tid; j < p.K; j += blockDim.x)
      best = better(best, Pick{(1.f - p.w) * att[j] + p.w * (psi[j] - prev), cs[j], j});
  const float* row = p.proc + (long long)b * p.V;
  const float ts_ctc = p.w * (maxpsi - prev);
  for (int v = p.first_ts + tid; v < p.V; v += blockDim.x)
    best = better(best, Pick{(1.f - p.w) * (row[v] - lse) + ts_ctc, v, -1});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Pick q;
    q.v = __shfl_xor_sync(0xffffffffu, best.v, o);
    q.i = __shfl_xor_sync(0xffffffffu, best.i, o);
    q.slot = __shfl_xor_sync(0xffffffffu, best.slot, o);
    best = better(best, q);
  }
  if (lane == 0) s_p[warp] = best;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = better(best, s_p[w]);
    if (best.i >= p.V) best.i = p.eos, best.slot = -1;  // nothing finite: stay defined
    const int unf = p.unfinished[b];
    if (!unf) best.i = p.pad, best.slot = -1;
    p.ids_out[(long long)b * p.ids_rs + len] = best.i;
    p.unfinished[b] = unf && (best.i != p.eos);
    s_win = best;
  }
  __syncthreads();
  // update_state (decoding.py:253-260): a text token moves the hypothesis to its candidate's forward variables / score
  const Pick win = s_win;
  if (win.i < p.first_ts) {
    if (win.slot >= 0) {
      const float* st = p.states + (long long)b * p.T * 2 * p.K + win.slot;
      float* rp = p.r_prev + (long long)b * p.T * 2;
      for (int i = tid; i < 2 * p.T; i += blockDim.x) rp[i] = st[(long long)i * p.K];
      if (tid == 0) p.score_prev[b] = psi[win.slot];
    } else if (tid == 0) {
      p.score_prev[b] = CTC_LOGZERO;  // pad on a finished row: the reference reads the reset score table
    }
  }
}

// out[r, :] = in[r, :] - logsumexp(in[r, :])   (in place allowed)
__global__ void __launch_bounds__(512) log_softmax_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int V) {
  __shared__ float s_m[16], s_s[16];
  __shared__ float s_lse;
  const long long r = blockIdx.x;
  const float* x = in + r * V;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float m = -INFINITY, s = 0.f;
  for (int v = tid; v < V; v += blockDim.x) {
    const float a = x[v];
    if (a > -INFINITY) {
      const float mn = fmaxf(m, a);
      s = s * __expf(m - mn) + __expf(a - mn);
      m = mn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) s_m[warp] = m, s_s[warp] = s;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const float mn = fmaxf(m, s_m[w]);
      s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (s_m[w] == -INFINITY ? 0.f : s_s[w] * __expf(s_m[w] - mn));
      m = mn;
    }
    s_lse = m + logf(s);
  }
  __syncthreads();
  const float lse = s_lse;
  float* o = out + r * V;
  for (int v = tid; v < V; v += blockDim.x) o[v] = x[v] - lse;
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_log_softmax_rows(dicow_handle_t h, const float* in, float* out, int64_t rows, int V, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, in && out && rows >= 1 && rows <= 0x7fffffff && V >= 1, "dicow_log_softmax_rows: bad args");
  log_softmax_rows_kernel<<<(unsigned)rows, 512, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(in, out, V);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_ctc_joint_step(dicow_handle_t h, const dicow_ctc_joint_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_ctc_joint_args_t), "dicow_ctc_joint_step: bad args struct");
  const bool cand_only = a->score_only && a->ctc_weight == 0.f;  // attention-only beam search: normaliser + candidates
  DICOW_REQUIRE(ctx, a->ids && a->processed_scores && a->workspace_i32 && a->workspace_f32 &&
                         (cand_only || (a->ctc_logp && a->states && a->r_prev && a->score_prev && a->unfinished)),
                "dicow_ctc_joint_step: null argument");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->B <= 64 && (cand_only || (a->T >= 2 && a->T <= 4096)) && a->K >= 1 && a->K <= 512 &&
                         a->first_timestamp >= a->K && a->first_timestamp <= a->V && (cand_only || a->V1 > a->blank) && a->eos < a->first_timestamp,
                "dicow_ctc_joint_step: need 1 <= B <= 64, 2 <= T <= 4096, 1 <= K <= 512 <= first_timestamp <= V (B=%d T=%d K=%d)",
                a->B, a->T, a->K);
  CtcJointParams p{};
  p.ids = reinterpret_cast<const long long*>(a->ids), p.ids_out = reinterpret_cast<long long*>(a->ids), p.ids_rs = a->ids_row_stride;
  p.pos = a->pos, p.cur_len = a->cur_len;
  p.B = a->B, p.V = a->V, p.T = a->T, p.V1 = a->V1, p.K = a->K;
  p.bos = a->bos, p.eos = a->eos, p.pad = a->pad, p.blank = a->blank, p.first_ts = a->first_timestamp, p.prefix_len = a->prefix_len;
  p.w = a->ctc_weight;
  p.x = a->ctc_logp, p.proc = a->processed_scores, p.raw = a->raw_logits;
  p.meta = a->workspace_i32;                      // 4 B + 1
  p.cs = a->workspace_i32 + 4 * a->B + 4;         // B K
  p.lse = a->workspace_f32;                       // B
  p.att = a->workspace_f32 + a->B;                // B K
  p.psi = a->workspace_f32 + a->B + (size_t)a->B * a->K;  // B K
  p.states = a->states, p.r_prev = a->r_prev, p.score_prev = a->score_prev, p.unfinished = a->unfinished;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ctc_prepare_kernel<<<1, 64, 0, stream>>>(p);
  ctc_topk_kernel<<<a->B, TK_THREADS, 0, stream>>>(p);
  // one thread per candidate, 128 per CTA: the per-frame recursion is issue-bound on one SM with 512 candidates per CTA
  if (!cand_only) ctc_prefix_score_kernel<<<dim3(a->B, (a->K + 127) / 128), 128, 3 * a->T * sizeof(float), stream>>>(p);
  if (!a->score_only) ctc_combine_select_kernel<<<a->B, 512, 0, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
