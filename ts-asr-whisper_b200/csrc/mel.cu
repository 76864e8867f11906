// mel.cu -- Whisper log-mel front-end, fused: framing (reflect-padded, centred) + periodic Hann window + 400-point real
// DFT + power + slaney mel filterbank + log10, then a second tiny pass for the per-recording (max - 8) floor and the
// (x + 4) / 4 normalisation.
//
// Replaces WhisperFeatureExtractor._torch_extract_fbank_features (HF:models/whisper/feature_extraction_whisper.py:
// 135-164; torch.stft on the CPU in a DataLoader worker) as called by the reference at
// src/data/local_datasets.py:208-214 (padding="longest", pad_to_multiple_of=480000, return_attention_mask=True).
//
// One CTA = 64 consecutive frames of one recording.  The DFT is a direct real transform that exploits the even/odd
// symmetry of the basis:  Re X[f] = sum_{k=0..200} yE[k] cos(2 pi f k / 400),  Im X[f] = -sum_{k=1..199} yO[k] sin(..)
// with yE[k] = y[k] + y[400-k], yO[k] = y[k] - y[400-k], y = window * samples -- half the MACs of the plain DFT.  It is
// computed as a register-tiled fp32 mini-GEMM [64 frames x 201 k] x [201 k x 201 f] (thread tile 4 frames x 13
// frequencies x {re, im}); the basis streams from a 346 KB L2-resident table through shared memory.  fp32 throughout:
// the log10 of near-cancelling bins is too sensitive for bf16 tensor-core operands (estimated ~1.5e-3 output error).
// Cost: ~0.5 GFLOP per 30 s window, i.e. FP32-FMA bound (~8 us / window), ~0.3 % of an encoder forward.
#include <math.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int NFFT = 400;
constexpr int HOP = 160;
constexpr int NFREQ = 201;
constexpr int NFP = 208;  // frequencies / k rows padded to 16 x 13
constexpr int FR = 64;    // frames per CTA
constexpr int KT = 8;     // k rows per basis tile
constexpr int kMelThreads = 256;
constexpr int RAW = (FR - 1) * HOP + NFFT;        // samples spanned by the CTA's frames
constexpr int TILE_FLOATS = KT * 2 * NFP;         // 3328 = 13 per thread
constexpr int TABLE_WINDOW = 512;                 // window[400] padded
constexpr int TABLE_FLOATS = TABLE_WINDOW + NFP * 2 * NFP;

struct MelParams {
  const float* audio;  // [B, n_pad]
  long long audio_bs;
  long long n_pad;
  int frames;  // n_pad / 160
  int n_mels;
  const float* filters;  // [201, n_mels]
  const float* tables;   // window + basis
  float* out;            // [B, n_mels, frames]: log10(max(mel, 1e-10)) after pass 1
  unsigned* gmax;        // [B] order-preserving encoding of the running max
  const long long* lengths;
  int* mask;
};

__device__ __forceinline__ unsigned encode_ordered(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float decode_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(kMelThreads, 1) logmel_frames_kernel(const MelParams p) {
  extern __shared__ __align__(16) float sm[];
  float* yE = sm;                  // [NFP][FR]  (later aliased by the power spectrum P[f][fr])
  float* yO = yE + NFP * FR;       // [NFP][FR]
  float* stage = yO + NFP * FR;    // raw samples [RAW], later basis tiles [2][TILE_FLOATS]... (RAW > TILE_FLOATS)
  int* lo = reinterpret_cast<int*>(stage + RAW);  // [n_mels] first / one-past-last non-zero frequency of each filter
  int* hi = lo + 128;
  __shared__ float red[kMelThreads / 32];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const float* x = p.audio + (long long)b * p.audio_bs;
  const float* window = p.tables;
  const float* basis = p.tables + TABLE_WINDOW;

  // ---- raw samples of the CTA's frames, reflect-padded like torch.stft(center=True, pad_mode="reflect") ----
  const long long s0 = (long long)t0 * HOP - NFFT / 2;
  for (int i = tid; i < RAW; i += kMelThreads) {
    long long g = s0 + i;
    if (g < 0) g = -g;
    if (g >= p.n_pad) g = 2 * (p.n_pad - 1) - g;
    g = g < 0 ? 0 : (g >= p.n_pad ? p.n_pad - 1 : g);  // frames past the end of the recording: never stored
    stage[i] = __ldg(x + g);
  }
  // ---- non-zero range of every mel filter ----
  if (tid < p.n_mels) {
    int l = NFREQ, h = 0;
    for (int f = 0; f < NFREQ; ++f) {
      if (__ldg(p.filters + (long long)f * p.n_mels + tid) != 0.0f) {
        l = min(l, f);
        h = f + 1;
      }
    }
    lo[tid] = min(l, h);
    hi[tid] = h;
  }
  __syncthreads();
  // ---- windowed even / odd folds ----
  for (int i = tid; i < NFP * FR; i += kMelThreads) {
    const int k = i / FR, fr = i - k * FR;
    float e = 0.f, o = 0.f;
    if (k <= NFFT / 2) {
      const float w = __ldg(window + k);
      const float a = w * stage[fr * HOP + k];
      if (k == 0 || k == NFFT / 2) {
        e = a;
      } else {
        const float c = w * stage[fr * HOP + NFFT - k];  // Hann is symmetric: w[400 - k] == w[k]
        e = a + c;
        o = a - c;
      }
    }
    yE[i] = e;
    yO[i] = o;
  }
  __syncthreads();  // stage[] is free from here on: it becomes the basis tile buffer

  // ---- DFT: thread tile = 4 frames x 13 frequencies x {re, im} ----
  const int fg = tid >> 4;  // frames fg*4 .. +3
  const int qg = tid & 15;  // frequencies qg + 16 i
  float re[4][13], im[4][13];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 13; ++i) re[a][i] = 0.f, im[a][i] = 0.f;
  float pre[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) pre[j] = __ldg(basis + tid + kMelThreads * j);
  constexpr int NTILES = NFP / KT;  // 26
  for (int tile = 0; tile < NTILES; ++tile) {
    float* cur = stage + (tile & 1) * TILE_FLOATS;
#pragma unroll
    for (int j = 0; j < 13; ++j) cur[tid + kMelThreads * j] = pre[j];
    __syncthreads();  // tile visible; the other buffer (read two iterations ago) is free to overwrite next time
    if (tile + 1 < NTILES) {
#pragma unroll
      for (int j = 0; j < 13; ++j) pre[j] = __ldg(basis + (long long)(tile + 1) * TILE_FLOATS + tid + kMelThreads * j);
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int k = tile * KT + kk;
      const float4 e4 = *reinterpret_cast<const float4*>(yE + k * FR + fg * 4);
      const float4 o4 = *reinterpret_cast<const float4*>(yO + k * FR + fg * 4);
      const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
      const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
      const float* cs = cur + kk * 2 * NFP + qg;
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        const float c = cs[16 * i];
        const float s = cs[NFP + 16 * i];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          re[a][i] = fmaf(ev[a], c, re[a][i]);
          im[a][i] = fmaf(ov[a], s, im[a][i]);
        }
      }
    }
  }
  __syncthreads();  // everyone is done reading yE / yO
  // ---- power spectrum P[f][fr] over yE ----
  float* P = yE;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const int f = qg + 16 * i;
    float4 v;
    v.x = fmaf(re[0][i], re[0][i], im[0][i] * im[0][i]);
    v.y = fmaf(re[1][i], re[1][i], im[1][i] * im[1][i]);
    v.z = fmaf(re[2][i], re[2][i], im[2][i] * im[2][i]);
    v.w = fmaf(re[3][i], re[3][i], im[3][i] * im[3][i]);
    *reinterpret_cast<float4*>(P + f * FR + fg * 4) = v;
  }
  __syncthreads();
  // ---- mel filterbank + log10; thread = (frame, mel bin m = mg + 4 i); output is frame-contiguous ----
  const int fr = tid & (FR - 1);
  const int mg = tid >> 6;
  const int t = t0 + fr;
  float vmax = -INFINITY;
  for (int m = mg; m < p.n_mels; m += kMelThreads / FR) {
    float acc = 0.f;
    const int h = hi[m];
    for (int f = lo[m]; f < h; ++f) acc = fmaf(P[f * FR + fr], __ldg(p.filters + (long long)f * p.n_mels + m), acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    if (t < p.frames) {
      p.out[((long long)b * p.n_mels + m) * p.frames + t] = v;
      vmax = fmaxf(vmax, v);
    }
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0) red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
#pragma unroll
    for (int i = 1; i < kMelThreads / 32; ++i) m = fmaxf(m, red[i]);
    if (m > -INFINITY) atomicMax(p.gmax + b, encode_ordered(m));
  }
}

// pass 2: x = max(x, recording max - 8); (x + 4) / 4; attention_mask[t] = (t * 160 < length)
__global__ void __launch_bounds__(256) logmel_finalize_kernel(const MelParams p) {
  const int b = blockIdx.y;
  const float floor_v = decode_ordered(p.gmax[b]) - 8.0f;
  const long long n = (long long)p.n_mels * p.frames;
  float* o = p.out + (long long)b * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = (fmaxf(o[i], floor_v) + 4.0f) * 0.25f;
  if (p.mask != nullptr) {
    const long long len = p.lengths != nullptr ? p.lengths[b] : p.n_pad;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.frames; t += gridDim.x * blockDim.x)
      p.mask[(long long)b * p.frames + t] = ((long long)t * HOP < len) ? 1 : 0;
  }
}

constexpr size_t kMelSmem = (size_t)(2 * NFP * FR + RAW) * sizeof(float) + 256 * sizeof(int);

// ------------------------------------------------------------------------------------------------------------------
// STNO mask from per-speaker sample-level activity (src/data/local_datasets.py:162-196): one thread per encoder frame;
// a_i = mean of speaker i's 0/1 activity over the frame's 320 samples (zero past the end of the recording), then
//   S = prod_i (1 - a_i), T = a_s prod_{i != s} (1 - a_i), N = (1 - a_s)(1 - prod_{i != s} (1 - a_i)), O = a_s - T
// in fp32 with the speakers multiplied in row order (bit-exact against numpy).  HBM-bound: n_speakers bytes per sample.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stno_mask_kernel(const unsigned char* __restrict__ act, long long ld, int n_spk,
                                                        long long n_samples, int target, int frame_samples, long long frames,
                                                        float* __restrict__ out, long long frame_stride, long long class_stride) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= frames) return;
  const long long s0 = t * frame_samples;
  float sil = 1.f, others = 1.f, a_s = 0.f;
  for (int i = 0; i < n_spk; ++i) {
    const unsigned char* row = act + (long long)i * ld + s0;
    int cnt = 0;
    for (int k = 0; k < frame_samples; ++k)
      if (s0 + k < n_samples) cnt += row[k] != 0;
    const float a = __fdiv_rn((float)cnt, (float)frame_samples);
    const float na = __fsub_rn(1.f, a);
    sil = __fmul_rn(sil, na);
    if (i != target) others = __fmul_rn(others, na);
    else a_s = a;
  }
  const float tgt = __fmul_rn(a_s, others);
  float* o = out + t * frame_stride;
  o[0] = sil;
  o[class_stride] = tgt;
  o[2 * class_stride] = __fmul_rn(__fsub_rn(1.f, a_s), __fsub_rn(1.f, others));
  o[3 * class_stride] = __fsub_rn(a_s, tgt);
}

}  // namespace

// window + DFT basis, built in double on the host once per handle (dicow_create)
int mel_tables_create(dicow_ctx* ctx) {
  std::vector<float> t(TABLE_FLOATS, 0.0f);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < NFFT; ++k) t[k] = (float)(0.5 - 0.5 * cos(two_pi * k / NFFT));  // torch.hann_window(400), periodic
  for (int k = 0; k < NFREQ; ++k) {
    for (int f = 0; f < NFREQ; ++f) {
      const int j = (int)(((long long)f * k) % NFFT);  // exact argument reduction
      t[TABLE_WINDOW + (k * 2 + 0) * NFP + f] = (float)cos(two_pi * j / NFFT);
      t[TABLE_WINDOW + (k * 2 + 1) * NFP + f] = (float)sin(two_pi * j / NFFT);
    }
  }
  DICOW_CUDA_OK(ctx, cudaMalloc(&ctx->mel_tables, TABLE_FLOATS * sizeof(float)));
  DICOW_CUDA_OK(ctx, cudaMemcpy(ctx->mel_tables, t.data(), TABLE_FLOATS * sizeof(float), cudaMemcpyHostToDevice));
  return DICOW_OK;
}

}  // namespace dicow

using namespace dicow;

extern "C" int dicow_logmel(dicow_handle_t h, const dicow_logmel_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_logmel_args_t), "dicow_logmel: bad args struct");
  DICOW_REQUIRE(ctx, a->audio && a->mel_filters && a->out && a->workspace, "dicow_logmel: null operand");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->B <= 65535 && a->n_pad >= NFFT && (a->n_pad % HOP) == 0,
                "dicow_logmel: need 1 <= B <= 65535 and n_pad a multiple of 160, >= 400 (got B=%d n_pad=%lld)", a->B,
                (long long)a->n_pad);
  DICOW_REQUIRE(ctx, a->n_mels >= 1 && a->n_mels <= 128, "dicow_logmel: n_mels must be in [1, 128]");
  DICOW_REQUIRE(ctx, a->attention_mask == nullptr || a->lengths != nullptr, "dicow_logmel: attention_mask needs lengths");
  DICOW_REQUIRE(ctx, ctx->mel_tables != nullptr, "dicow_logmel: tables not initialised");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MelParams p{};
  p.audio = a->audio, p.audio_bs = a->audio_batch_stride, p.n_pad = a->n_pad;
  p.frames = (int)(a->n_pad / HOP);
  p.n_mels = a->n_mels, p.filters = a->mel_filters, p.tables = ctx->mel_tables;
  p.out = a->out, p.gmax = reinterpret_cast<unsigned*>(a->workspace);
  p.lengths = reinterpret_cast<const long long*>(a->lengths), p.mask = a->attention_mask;
  static DeviceOnce attr_once;
  if (attr_once.first(ctx)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)kMelSmem));
  }
  DICOW_CUDA_OK(ctx, cudaMemsetAsync(a->workspace, 0, sizeof(unsigned) * a->B, stream));
  dim3 grid(ceil_div(p.frames, FR), a->B);
  logmel_frames_kernel<<<grid, kMelThreads, kMelSmem, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  const long long n = (long long)p.n_mels * p.frames;
  dim3 grid2((unsigned)((n + 256 * 8 - 1) / (256 * 8)), a->B);
  logmel_finalize_kernel<<<grid2, 256, 0, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_stno_mask(dicow_handle_t h, const uint8_t* activity, int64_t ld, int n_speakers, int64_t n_samples,
                               int target, int frame_samples, int64_t frames, float* out, int64_t frame_stride,
                               int64_t class_stride, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, activity && out && n_speakers >= 1 && n_samples >= 1 && target >= -1 && target < n_speakers &&
                         frame_samples >= 1 && frames >= 1 && ld >= n_samples,
                "dicow_stno_mask: bad args");
  stno_mask_kernel<<<(unsigned)((frames + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      activity, ld, n_speakers, n_samples, target, frame_samples, frames, out, frame_stride, class_stride);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
