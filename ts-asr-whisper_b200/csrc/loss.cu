// loss.cu -- loss kernels of the training / evaluation forward (HBM-bound: every logit is read once).
//   * soft-label cross entropy of the decoder: src/models/dicow/modeling_dicow.py:95-144 (SoftLabelCreator.compute_loss)
//     and the hard-label fallback of modeling_dicow.py:312-323.  The reference materialises two dense one-hot
//     [B*S, V] fp32 target tensors; the targets are sparse (one id, or a Gaussian over the 1501 timestamp ids), so the
//     loss per token is  lse(logits) - sum_k w_k logit[k]  with at most 1501 non-zero w_k.
//   * CTC loss of the encoder head: src/models/dicow/encoder.py:108-135 (fp32 log-softmax, blank = last class, all
//     frames valid, zero_infinity=True); alpha recursion in log space, one CTA per utterance, one thread per state of
//     the blank-extended label sequence.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int LT = 512;  // threads per row

__device__ __forceinline__ float block_lse(const float* __restrict__ row, int V, float* sm_a, float* sm_b) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  float m = -INFINITY, s = 0.f;
  for (int v = tid; v < V; v += blockDim.x) {
    const float x = row[v];
    if (x > m) {
      s = s * __expf(m - x) + 1.0f;
      m = x;
    } else if (x > -INFINITY) {
      s += __expf(x - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) sm_a[warp] = m, sm_b[warp] = s;
  __syncthreads();
  float mt = -INFINITY;
  for (int w = 0; w < nw; ++w) mt = fmaxf(mt, sm_a[w]);
  float st = 0.f;
  for (int w = 0; w < nw; ++w) st += sm_a[w] == -INFINITY ? 0.f : sm_b[w] * __expf(sm_a[w] - mt);
  __syncthreads();
  return mt + logf(st);
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += sm[w];
  __syncthreads();
  return t;
}

// lse[r] = logsumexp(logits[r, :])
__global__ void __launch_bounds__(LT) row_lse_kernel(const float* __restrict__ logits, long long ld, int V,
                                                     float* __restrict__ lse) {
  __shared__ float sa[LT / 32], sb[LT / 32];
  const float v = block_lse(logits + (long long)blockIdx.x * ld, V, sa, sb);
  if (threadIdx.x == 0) lse[blockIdx.x] = v;
}

struct CeParams {
  const float* logits;
  long long ld;
  int R, V;
  const long long* labels;
  const long long* upp;  // or NULL
  int ts_begin, n_ts;    // n_ts == 0: hard labels only
  const float* smooth;   // [n_ts, n_ts] row-normalised Gaussian
  int soft_mode;         // 1: mask = labels != -100, mean over masked rows; 0: hard fallback, mean over all rows
  float* row_loss;       // [R]
  float* row_mask;       // [R]
  float* out;            // [1]
};

// -sum_v target_v log_softmax(logits)_v for the target distribution of one label
__device__ __forceinline__ float ce_one(const CeParams& p, const float* row, float lse, long long lab, float* sm) {
  if (lab < 0) lab = 0;  // the reference clamps -100 to 0 before one_hot; such rows are masked afterwards
  if (lab >= p.V) lab = p.V - 1;
  if (p.n_ts > 0 && lab >= p.ts_begin && lab < p.ts_begin + p.n_ts) {
    const float* w = p.smooth + (lab - p.ts_begin) * p.n_ts;
    float acc = 0.f;
    for (int k = threadIdx.x; k < p.n_ts; k += blockDim.x) acc = fmaf(__ldg(w + k), row[p.ts_begin + k] - lse, acc);
    return -block_sum(acc, sm);
  }
  return lse - row[lab];
}

__global__ void __launch_bounds__(LT) softlabel_ce_rows_kernel(const CeParams p) {
  __shared__ float sa[LT / 32], sb[LT / 32];
  const int r = blockIdx.x;
  const float* row = p.logits + (long long)r * p.ld;
  const long long lab = p.labels[r];
  const long long ulab = p.upp != nullptr ? p.upp[r] : lab;
  const float lse = block_lse(row, p.V, sa, sb);
  float lo, up;
  if (p.soft_mode) {
    lo = ce_one(p, row, lse, lab, sa);
    up = (p.upp != nullptr) ? ce_one(p, row, lse, ulab, sa) : lo;
    const float mk = lab != -100 ? 1.f : 0.f;
    lo *= mk, up *= mk;
    if (threadIdx.x == 0) p.row_mask[r] = mk;
  } else {  // CrossEntropyLoss(reduction='none') with ignore_index=-100: ignored rows contribute 0 to the mean
    lo = lab == -100 ? 0.f : ce_one(p, row, lse, lab, sa);
    up = (p.upp != nullptr) ? (ulab == -100 ? 0.f : ce_one(p, row, lse, ulab, sa)) : lo;
    if (threadIdx.x == 0) p.row_mask[r] = 1.f;
  }
  if (threadIdx.x == 0) p.row_loss[r] = fminf(lo, up);
}

// out = sum(row_loss) / max(sum(row_mask), 1)   (deterministic single-CTA reduction)
__global__ void __launch_bounds__(LT) masked_mean_kernel(const float* __restrict__ v, const float* __restrict__ mk, int n,
                                                         float* __restrict__ out) {
  __shared__ float sm[LT / 32];
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i], c += mk[i];
  a = block_sum(a, sm);
  c = block_sum(c, sm);
  if (threadIdx.x == 0) out[0] = a / fmaxf(c, 1.0f);
}

// ------------------------------------------------------------------------------------------------------------------
// CTC
// ------------------------------------------------------------------------------------------------------------------
struct CtcParams {
  const float* logits;  // [B, T, V1], rows ld elements apart
  long long ld;
  const float* lse;     // [B * T]
  int T, V1, Lmax;
  const long long* labels;  // [B, Lmax], negative = padding (a prefix of each row is valid)
  float* nll;               // [B] per-utterance negative log-likelihood (0 where infinite: zero_infinity)
  float* scaled;            // [B] nll / max(len, 1) for reduction "mean", nll for "sum"
  int mean;
};

__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(__expf(-fabsf(a - b)));
}

__global__ void __launch_bounds__(1024) ctc_alpha_kernel(const CtcParams p) {
  extern __shared__ float alpha[];  // [2][S]
  const int b = blockIdx.x, s = threadIdx.x;
  const long long* lab = p.labels + (long long)b * p.Lmax;
  __shared__ int s_len;
  if (s == 0) {
    int n = 0;
    for (int i = 0; i < p.Lmax; ++i) n += lab[i] >= 0 ? 1 : 0;
    s_len = n;
  }
  __syncthreads();
  const int L = s_len, S = 2 * L + 1, blank = p.V1 - 1;
  const bool active = s < S;
  int cls = blank;
  bool skip = false;
  if (active && (s & 1)) {
    cls = (int)lab[s >> 1];
    skip = s >= 3 && lab[(s >> 1) - 1] != cls;  // s-2 holds a different non-blank label
  }
  const float* lg = p.logits + (long long)b * p.T * p.ld;
  const float* ls = p.lse + (long long)b * p.T;
  float* cur = alpha;
  float* nxt = alpha + 2 * p.Lmax + 1;
  if (active) cur[s] = (s <= 1) ? lg[cls] - ls[0] : -INFINITY;
  __syncthreads();
  for (int t = 1; t < p.T; ++t) {
    if (active) {
      float a = cur[s];
      if (s >= 1) a = log_add(a, cur[s - 1]);
      if (skip) a = log_add(a, cur[s - 2]);
      nxt[s] = a + (lg[(long long)t * p.ld + cls] - ls[t]);
    }
    __syncthreads();
    float* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (s == 0) {
    float ll = cur[S - 1];
    if (S >= 2) ll = log_add(ll, cur[S - 2]);
    float nll = -ll;
    if (!(nll < INFINITY)) nll = 0.f;  // zero_infinity=True (also catches NaN)
    p.nll[b] = nll;
    p.scaled[b] = p.mean ? nll / (float)max(L, 1) : nll;
  }
}

__global__ void __launch_bounds__(LT) mean_kernel(const float* __restrict__ v, int n, int divide, float* __restrict__ out) {
  __shared__ float sm[LT / 32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
  a = block_sum(a, sm);
  if (threadIdx.x == 0) out[0] = divide ? a / (float)n : a;
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_softlabel_ce(dicow_handle_t h, const dicow_softlabel_ce_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_softlabel_ce_args_t), "dicow_softlabel_ce: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->labels && a->workspace && a->loss && a->rows >= 1 && a->V >= 1,
                "dicow_softlabel_ce: bad args");
  DICOW_REQUIRE(ctx, a->n_ts == 0 || (a->smoothing != nullptr && a->ts_begin >= 0 && a->ts_begin + a->n_ts <= a->V),
                "dicow_softlabel_ce: timestamp smoothing needs smoothing[n_ts, n_ts] and ts_begin + n_ts <= V");
  CeParams p{};
  p.logits = a->logits, p.ld = a->ld, p.R = a->rows, p.V = a->V;
  p.labels = reinterpret_cast<const long long*>(a->labels), p.upp = reinterpret_cast<const long long*>(a->upp_labels);
  p.ts_begin = a->ts_begin, p.n_ts = a->n_ts, p.smooth = a->smoothing, p.soft_mode = a->soft_mode;
  p.row_loss = a->workspace, p.row_mask = a->workspace + a->rows, p.out = a->loss;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  softlabel_ce_rows_kernel<<<a->rows, LT, 0, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  masked_mean_kernel<<<1, LT, 0, stream>>>(p.row_loss, p.row_mask, a->rows, a->loss);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_ctc_loss(dicow_handle_t h, const dicow_ctc_loss_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_ctc_loss_args_t), "dicow_ctc_loss: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->labels && a->workspace && a->loss && a->B >= 1 && a->T >= 1 && a->V1 >= 2,
                "dicow_ctc_loss: bad args");
  DICOW_REQUIRE(ctx, a->Lmax >= 0 && 2 * a->Lmax + 1 <= 1024, "dicow_ctc_loss: at most 511 labels per utterance (got %d)",
                a->Lmax);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  float* lse = a->workspace;                   // [B * T]
  float* nll = lse + (long long)a->B * a->T;   // [B]
  float* scaled = nll + a->B;                  // [B]
  const long long ld = a->ld > 0 ? a->ld : a->V1;
  DICOW_REQUIRE(ctx, ld >= a->V1, "dicow_ctc_loss: ld < V1");
  row_lse_kernel<<<a->B * a->T, LT, 0, stream>>>(a->logits, ld, a->V1, lse);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  CtcParams p{};
  p.logits = a->logits, p.ld = ld, p.lse = lse, p.T = a->T, p.V1 = a->V1, p.Lmax = a->Lmax;
  p.labels = reinterpret_cast<const long long*>(a->labels), p.nll = nll, p.scaled = scaled, p.mean = a->reduction_mean;
  const int S = 2 * a->Lmax + 1;
  const int threads = ((S + 31) / 32) * 32;
  ctc_alpha_kernel<<<a->B, threads, 2 * S * sizeof(float), stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  mean_kernel<<<1, LT, 0, stream>>>(scaled, a->B, a->reduction_mean, a->loss);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
