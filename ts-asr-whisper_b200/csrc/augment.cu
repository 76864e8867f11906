// augment.cu -- training-time augmentations of a padded batch on the GPU: soft STNO segment changes, Gaussian STNO noise
// with rescaling, and SpecAug (bicubic time warp + frequency / time masks) on [mel || STNO].
//
// Replaces the tensor code DataCollator.__call__ runs on the CPU in a DataLoader worker (src/data/collators.py:184-210:
// soft_segment_augmentation :77-136, add_gaussian_noise_and_rescale :50-75, SpecAug src/data/augmentations.py:70-98
// (time_warp = F.interpolate bicubic), :20-66 (mask_along_axis), :414-434 (SpecAug.forward)).  Every random number of the
// reference is data independent, so the host draws them in the reference's order (same torch seed -> same plan,
// ts_asr_whisper_b200/augment.py) and the kernels only apply the plan.
//
// All three kernels are HBM-bound element-wise passes (a recipe batch: 8 x 132 x 3000 fp32 = 12.7 MB in, the same out);
// SpecAug is ONE pass: the reference materialises the concatenated [B, 3000, 132] tensor, two interpolated halves, two
// masked copies and the split / stacked STNO, ~8 passes over the batch.
//
// Arithmetic is pinned with explicit round-to-nearest intrinsics (no compiler contraction) to the operation order of
// oracle/augment.py: bit-exact on the STNO augmentations and the masks, and the bicubic taps use the fused multiply-adds
// that reproduce the reference's PyTorch CPU build to <= 7e-7 (see the oracle's _bicubic_rows).
#include <math.h>

#include "common.h"

namespace dicow {
namespace {

// ---- soft segment augmentation: one CTA per changed segment -------------------------------------------------------
// seg[i] = (batch row, start, end, index into the classes other than the dominant one); soft[i] = (softness, 1 - softness)
__global__ void __launch_bounds__(128) stno_segment_kernel(float* __restrict__ stno, int C, int T, const int* __restrict__ seg,
                                                           const float* __restrict__ soft) {
  __shared__ float s_mean[8];
  __shared__ int s_target;
  const int i = blockIdx.x;
  const int b = seg[4 * i], start = seg[4 * i + 1], end = seg[4 * i + 2], which = seg[4 * i + 3];
  float* base = stno + (long long)b * C * T;
  if ((int)threadIdx.x < C) {  // mean of the class over the segment (collators.py:110)
    float acc = 0.f;
    for (int t = start; t < end; ++t) acc = __fadd_rn(acc, base[(long long)threadIdx.x * T + t]);
    s_mean[threadIdx.x] = __fdiv_rn(acc, (float)(end - start));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int dom = 0;
    for (int c = 1; c < C; ++c)
      if (s_mean[c] > s_mean[dom]) dom = c;          // argmax, first maximum
    s_target = which < dom ? which : which + 1;      // available_classes.remove(dominant)[which] (collators.py:113-118)
  }
  __syncthreads();
  const int target = s_target;
  const float softness = soft[2 * i], keep = soft[2 * i + 1];
  for (int t = start + (int)threadIdx.x; t < end; t += blockDim.x) {
    float v[8], tot = 0.f;
    for (int c = 0; c < C; ++c) {
      v[c] = __fadd_rn(__fmul_rn(keep, base[(long long)c * T + t]), __fmul_rn(softness, c == target ? 1.f : 0.f));
      tot = c == 0 ? v[0] : __fadd_rn(tot, v[c]);
    }
    for (int c = 0; c < C; ++c) base[(long long)c * T + t] = __fdiv_rn(v[c], tot);
  }
}

// ---- Gaussian noise + shift + renormalise (collators.py:63-75): one thread per (selected row, frame) ---------------
__global__ void __launch_bounds__(256) stno_noise_kernel(float* __restrict__ stno, int C, int T, const int* __restrict__ rows,
                                                         const float* __restrict__ noise, int n) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * T) return;
  const int j = (int)(gid / T), t = (int)(gid - (long long)j * T);
  float* base = stno + (long long)rows[j] * C * T + t;
  const float* nz = noise + (long long)j * C * T + t;
  float v[8], lo = 0.f;
  for (int c = 0; c < C; ++c) {
    v[c] = __fadd_rn(base[(long long)c * T], __ldg(nz + (long long)c * T));
    lo = fminf(lo, v[c]);
  }
  float tot = 0.f;
  for (int c = 0; c < C; ++c) {
    v[c] = __fsub_rn(v[c], lo);
    tot = c == 0 ? v[0] : __fadd_rn(tot, v[c]);
  }
  for (int c = 0; c < C; ++c) base[(long long)c * T] = __fdiv_rn(v[c], tot);
}

// ---- SpecAug -------------------------------------------------------------------------------------------------------
struct SpecParams {
  const float* feats;  // [B, M, Tf]
  const float* stno;   // [B, C, Ts], Tf = factor * Ts
  float* feats_out;    // [B, M, Tf]
  float* stno_out;     // [B, C, Ts]
  int B, M, C, Tf, Ts, factor;
  int center, warped;  // center < 0: no time warp
  const int* freq_masks;  // [B, nf, 2] (pos, length) over the first D channels of [mel || STNO]
  const int* time_masks;  // [B, nt, 2] or NULL
  int nf, nt, D;
};

__device__ __forceinline__ float cubic1(float x) {  // ((A + 2) x - (A + 3)) x x + 1, A = -0.75
  return __fmaf_rn(__fmul_rn(__fmaf_rn(1.25f, x, -2.25f), x), x, 1.f);
}
__device__ __forceinline__ float cubic2(float x) {  // ((A x - 5 A) x + 8 A) x - 4 A
  return __fmaf_rn(__fmaf_rn(__fmaf_rn(-0.75f, x, 3.75f), x, -6.f), x, 3.f);
}

// taps of output frame t of the warped signal: source slice [s0, s0 + L) resampled to `out_len` frames (ATen
// area_pixel_compute_source_index(cubic) + guard_index_and_lambda + get_cubic_upsample_coefficients)
__device__ __forceinline__ void warp_taps(const SpecParams& p, int t, int idx[4], float w[4]) {
  if (p.center < 0) {
    idx[0] = idx[1] = idx[2] = idx[3] = t;
    w[0] = 1.f, w[1] = w[2] = w[3] = 0.f;
    return;
  }
  const bool left = t < p.warped;
  const int s0 = left ? 0 : p.center, L = left ? p.center : p.Tf - p.center;
  const int out_len = left ? p.warped : p.Tf - p.warped, i = left ? t : t - p.warped;
  const float scale = __fdiv_rn((float)L, (float)out_len);
  const float real = __fmaf_rn(scale, __fadd_rn((float)i, 0.5f), -0.5f);
  const int i0 = min((int)floorf(real), L - 1);
  const float lam = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
  const float x2 = __fsub_rn(1.f, lam);
  w[0] = cubic2(__fadd_rn(lam, 1.f)), w[1] = cubic1(lam), w[2] = cubic1(x2), w[3] = cubic2(__fadd_rn(x2, 1.f));
#pragma unroll
  for (int j = 0; j < 4; ++j) idx[j] = s0 + max(min(i0 - 1 + j, L - 1), 0);
}

__device__ __forceinline__ bool masked(const int* __restrict__ m, int n, int x) {
  bool hit = false;
  for (int k = 0; k < n; ++k) {
    const int pos = __ldg(m + 2 * k), len = __ldg(m + 2 * k + 1);
    hit |= (pos <= x) && (x < pos + len);
  }
  return hit;
}

// grid (ceil(Tf / 256), M + C, B): thread = one output frame of one channel of [mel || STNO repeated x factor]
__global__ void __launch_bounds__(256) spec_augment_kernel(const SpecParams p) {
  const int t = blockIdx.x * 256 + threadIdx.x, ch = blockIdx.y, b = blockIdx.z;
  if (t >= p.Tf) return;
  const bool chan_masked = ch < p.D && masked(p.freq_masks + (long long)b * p.nf * 2, p.nf, ch);
  const bool time_zone = ch < p.D && p.time_masks != nullptr;
  if (ch < p.M) {
    float out = 0.f;
    if (!chan_masked && !(time_zone && masked(p.time_masks + (long long)b * p.nt * 2, p.nt, t))) {
      const float* src = p.feats + ((long long)b * p.M + ch) * p.Tf;
      int idx[4];
      float w[4];
      warp_taps(p, t, idx, w);
      out = __fmul_rn(__ldg(src + idx[0]), w[0]);
#pragma unroll
      for (int j = 1; j < 4; ++j) out = __fmaf_rn(__ldg(src + idx[j]), w[j], out);
    }
    p.feats_out[((long long)b * p.M + ch) * p.Tf + t] = out;
    return;
  }
  // STNO channel: the thread of the first frame of each group of `factor` frames averages the group (collators.py:210)
  if (t % p.factor != 0) return;
  const float* src = p.stno + ((long long)b * p.C + (ch - p.M)) * p.Ts;
  float acc = 0.f;
  for (int k = 0; k < p.factor; ++k) {
    const int tt = t + k;
    float v = 0.f;
    if (!chan_masked && !(time_zone && masked(p.time_masks + (long long)b * p.nt * 2, p.nt, tt))) {
      int idx[4];
      float w[4];
      warp_taps(p, tt, idx, w);
      v = __fmul_rn(__ldg(src + idx[0] / p.factor), w[0]);
#pragma unroll
      for (int j = 1; j < 4; ++j) v = __fmaf_rn(__ldg(src + idx[j] / p.factor), w[j], v);
    }
    acc = k == 0 ? v : __fadd_rn(acc, v);
  }
  p.stno_out[((long long)b * p.C + (ch - p.M)) * p.Ts + t / p.factor] = __fdiv_rn(acc, (float)p.factor);
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_augment_batch(dicow_handle_t h, const dicow_augment_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_augment_args_t), "dicow_augment_batch: bad args struct");
  DICOW_REQUIRE(ctx, a->stno != nullptr && a->B >= 1 && a->C >= 2 && a->C <= 8 && a->Ts >= 1,
                "dicow_augment_batch: need stno [B, C <= 8, Ts]");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a->n_seg > 0) {
    DICOW_REQUIRE(ctx, a->seg != nullptr && a->seg_soft != nullptr, "dicow_augment_batch: segments without tables");
    stno_segment_kernel<<<a->n_seg, 128, 0, stream>>>(a->stno, a->C, a->Ts, a->seg, a->seg_soft);
    DICOW_CUDA_OK(ctx, cudaGetLastError());
  }
  if (a->n_noise > 0) {
    DICOW_REQUIRE(ctx, a->noise_rows != nullptr && a->noise != nullptr, "dicow_augment_batch: noise rows without noise");
    const long long n = (long long)a->n_noise * a->Ts;
    stno_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a->stno, a->C, a->Ts, a->noise_rows, a->noise, a->n_noise);
    DICOW_CUDA_OK(ctx, cudaGetLastError());
  }
  if (a->spec) {
    DICOW_REQUIRE(ctx, a->feats != nullptr && a->feats_out != nullptr && a->stno_out != nullptr && a->feats != a->feats_out &&
                           a->stno != a->stno_out && a->M >= 1 && a->factor >= 1 && a->Tf == a->Ts * a->factor,
                  "dicow_augment_batch: SpecAug needs feats [B, M, Tf = factor * Ts] and distinct output buffers");
    DICOW_REQUIRE(ctx, a->n_freq_masks == 0 || a->freq_masks != nullptr, "dicow_augment_batch: n_freq_masks without table");
    DICOW_REQUIRE(ctx, a->center < 0 || (a->center >= 1 && a->center < a->Tf && a->warped >= 1 && a->warped < a->Tf),
                  "dicow_augment_batch: time warp needs 1 <= center, warped < Tf");
    DICOW_REQUIRE(ctx, a->M + a->C <= 65535 && a->B <= 65535, "dicow_augment_batch: batch / channel count exceeds the grid");
    SpecParams p;
    p.feats = a->feats, p.stno = a->stno, p.feats_out = a->feats_out, p.stno_out = a->stno_out;
    p.B = a->B, p.M = a->M, p.C = a->C, p.Tf = a->Tf, p.Ts = a->Ts, p.factor = a->factor;
    p.center = a->center, p.warped = a->warped;
    p.freq_masks = a->freq_masks, p.time_masks = a->n_time_masks > 0 ? a->time_masks : nullptr;
    p.nf = a->n_freq_masks, p.nt = a->n_time_masks;
    p.D = a->mask_channels < a->M + a->C ? a->mask_channels : a->M + a->C;
    dim3 grid((unsigned)((a->Tf + 255) / 256), (unsigned)(a->M + a->C), (unsigned)a->B);
    spec_augment_kernel<<<grid, 256, 0, stream>>>(p);
    DICOW_CUDA_OK(ctx, cudaGetLastError());
  }
  return DICOW_OK;
}
