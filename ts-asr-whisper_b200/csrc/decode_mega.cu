// decode_mega.cu -- the decoder LAYERS of one greedy token step as ONE persistent kernel (sm_100a).
//
// The step is HBM-bound on paper (344 MB of decoder weights + B x 30.7 MB of cross K/V per token: 0.128 ms at B = 16) but
// as a sequence of ~45 small kernels it is bound by their latency chains: every linear layer pays a launch ramp, one DRAM
// round trip and a reduction for 0.5-2 us worth of weight streaming (5.7-16 us each, 255 of the 430 us of a step; PDL and
// LayerNorm prologues were measured in round 1 and did not help).  Here one CTA per SM runs ALL phases of the layers with a
// grid-wide barrier between them (one 64-bit arrival counter in global memory):
//
//   embed | per layer: [LN1 -> q|k,v (+ cache append)] | self-attention | [out_proj + residual] | [LN2 -> q] |
//                      cross-attention | [out_proj + residual] | [LN3 -> fc1 + GELU] | [fc2 + residual]
//
// Linear phases (M <= 16 rows): the 8-column tiles of the weight matrix are dealt round-robin to the CTAs; a CTA requests the
// weights of ALL its tiles at once (16-byte loads straight into registers, no L1 allocation: the whole matrix is in flight
// GPU-wide after one instruction burst), stages the A operand once in shared memory -- LayerNorm(x) recomputed per CTA from
// the L2-resident fp32 stream (two rows per warp), or a copy of a bf16 activation -- multiplies with mma.sync m16n8k16 (each
// warp owns a K slice of every tile), reduces the eight partial tiles through shared memory and applies the epilogue
// (bias, GELU, residual add, q | k,v split with the cache append at the device-side position).
// Attention phases: self-attention = one warp per (row, head) over the <= 448 cached keys; cross-attention = one CTA per
// (row, head) streaming the head-major [T, k | v] slab (the decode_attention_kernel algorithm of decode.cu).
//
// Everything the phases exchange goes through L2 (ld.global.cg / plain stores + __threadfence before each barrier arrival).
// proj_out, the logits rules and the position advance stay separate kernels (decode.cu).
//
// Replaces, per generated token, the layer loop of HF WhisperDecoder.forward with a KV cache
// (HF:models/whisper/modeling_whisper.py:449-506, 691-796) inside DiCoWGenerationMixin._sample (generation.py:707-782).
#include <math.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int MK_WARPS = 8;
constexpr int MK_THREADS = MK_WARPS * 32;
constexpr int MK_ROWS = 16;       // one m16 tile of rows (decode batch <= 16)
constexpr int MK_MAX_T = 5;       // tiles a CTA multiplies at once
constexpr int MK_NBAR_LAYER = 8;  // grid barriers per layer (+ 1 after the embedding)

enum { MK_EPI_QKV = 0, MK_EPI_BF16 = 1, MK_EPI_GELU = 2, MK_EPI_RESID = 3 };

struct MegaParams {
  int B, d, H, ffn, L, T, S_max, vocab;
  const dicow_decode_layer_t* layers;  // [L] device table
  const long long* ids;
  long long ids_rs;
  const float* tok;
  const float* pos_emb;
  const int* pos;
  float* x;               // [B, d] fp32 residual stream of the step
  __nv_bfloat16* q;       // [B, d]
  __nv_bfloat16* ctx;     // [B, d]
  __nv_bfloat16* h;       // [B, ffn]
  float* attn_ws;         // [B * H][2][68]: cross-attention partials (o[64], m, l, pad) of the <= 2 CTAs that share a (row, head)
  unsigned long long* bar;  // arrival counter (monotonic; every launch adds gridDim.x * nbar)
  float eps;
  int a_pitch;  // shared-memory row pitch of the A operand, elements
  int flags;        // bit 0: stage the LayerNorm source rows with cp.async too (bf16 A rows always are); bit 1: L2-prefetch
                    // the next phase's weights before the barrier (both measured slower: DESIGN.md section 4.2)
  long long* prof;  // debug (dicow_debug_set_attention_profile): %globaltimer of CTA 0 at every phase boundary, else NULL
};

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {  // weights: read once, do not pollute L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {  // L2 -> shared, no registers, no L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x, f[2 * i + 1] = v.y;
  }
}

// grid-wide barrier: every CTA adds one to a monotonic counter and waits until the counter reaches `target`
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's global writes of the phase are visible before its arrival is
    atomicAdd(bar, 1ull);
    const long long t0 = clock64();
    while (ld_acquire_u64(bar) < target) {
      if (clock64() - t0 > DICOW_WAIT_TIMEOUT_CYCLES) __trap();  // a missing CTA traps instead of hanging the GPU
    }
  }
  __syncthreads();
}

struct LinearArgs {
  const float* x;                // LN source (fp32 [B, K]) or NULL
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* A;        // bf16 source [B, K] when x == NULL
  const __nv_bfloat16* W;        // [N, K]
  const float* bias;             // [N] or NULL
  int N, K;
  // outputs
  __nv_bfloat16* out_bf16;       // QKV: q (columns < n_split); BF16 / GELU: the output
  long long ldo;
  __nv_bfloat16* out2;           // QKV: cache base [B, S_max, 2d], columns >= n_split at row offset pos
  long long ldo2, pos_off;
  int n_split;
  float* xres;                   // RESID: x[row, n] += value
};

// One linear phase.  T tiles at once, KBW k32-blocks per warp (K / 32 <= 8 * KBW).
template <int T, int KBW, int EPI>
__device__ __forceinline__ void linear_phase(const MegaParams& p, const LinearArgs& a, __nv_bfloat16* sA, float* red,
                                          float* xstage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (a.N + 7) >> 3;
  const int kblocks = a.K >> 5;
  const int G = (int)gridDim.x;
  const int pitch = a.K + 32;  // A row pitch in elements: 64 (mod 128) bytes, conflict-free fragment loads
  bool staged = false;
  for (int tile0 = (int)blockIdx.x; tile0 < ntiles; tile0 += G * T) {
    // ---- 1. request the weights of this round's tiles (in flight while A is staged) ----
    uint4 w[T][KBW];
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int tile = tile0 + j * G;
      const int nrow = tile * 8 + g;
      const bool ok = tile < ntiles && nrow < a.N;
      const __nv_bfloat16* wrow = a.W + (long long)(ok ? nrow : 0) * a.K + t * 8;
#pragma unroll
      for (int i = 0; i < KBW; ++i) {
        const int kb = warp + i * MK_WARPS;
        w[j][i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok && kb < kblocks) w[j][i] = ldg_stream(wrow + kb * 32);
      }
    }
    // ---- 2. stage A (once per phase) ----
    if (!staged) {
      staged = true;
      if (a.x != nullptr) {
        // LayerNorm(x): the B fp32 rows come into shared memory with cp.async (every 16-byte piece in flight at once: one
        // L2 round trip), then two rows per warp are normalised out of shared memory (K <= 1280: <= 10 float4 per lane)
        const int nvec = a.K >> 2;
        const bool async_stage = (p.flags & 1) != 0;
        const float* xs = a.x;
        if (async_stage) {
          for (int i = threadIdx.x; i < p.B * nvec; i += MK_THREADS) cp_async16(xstage + (size_t)i * 4, a.x + (size_t)i * 4);
          cp_async_wait_all();
          __syncthreads();
          xs = xstage;
        }
        for (int r = warp; r < MK_ROWS; r += MK_WARPS) {
          __nv_bfloat16* srow = sA + (size_t)r * pitch;
          if (r >= p.B) {
            for (int c = lane * 4; c < a.K; c += 128) *reinterpret_cast<uint2*>(srow + c) = make_uint2(0u, 0u);
            continue;
          }
          const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)r * a.K);
          float4 v[10];
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < nvec ? (async_stage ? xr[c] : __ldcg(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 10; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
          const float mean = warp_sum_f(s) / (float)a.K;
          float qq = 0.f;
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            if (lane + 32 * i < nvec) {
              const float e0 = v[i].x - mean, e1 = v[i].y - mean, e2 = v[i].z - mean, e3 = v[i].w - mean;
              qq += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
            }
          }
          const float rstd = rsqrtf(warp_sum_f(qq) / (float)a.K + p.eps);
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < nvec) {
              const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma) + c4);
              const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta) + c4);
              const float y0 = fmaf((v[i].x - mean) * rstd, ga.x, be.x), y1 = fmaf((v[i].y - mean) * rstd, ga.y, be.y);
              const float y2 = fmaf((v[i].z - mean) * rstd, ga.z, be.z), y3 = fmaf((v[i].w - mean) * rstd, ga.w, be.w);
              *reinterpret_cast<uint2*>(srow + c4 * 4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
            }
          }
        }
      } else {
        // bf16 activation rows: cp.async straight into the padded A slab (160 KB for fc2: all of it in flight at once)
        const int vec_per_row = a.K >> 3;
        for (int i = threadIdx.x; i < MK_ROWS * vec_per_row; i += MK_THREADS) {
          const int r = i / vec_per_row, c = (i - r * vec_per_row) * 8;
          __nv_bfloat16* dst = sA + (size_t)r * pitch + c;
          if (r >= p.B)
            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
          else
            cp_async16(dst, a.A + (long long)r * a.K + c);
        }
        cp_async_wait_all();
      }
      __syncthreads();
    }
    // ---- 3. multiply: this warp's K slice of every tile ----
    float acc[T][4];
#pragma unroll
    for (int j = 0; j < T; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int i = 0; i < KBW; ++i) {
      const int kb = warp + i * MK_WARPS;
      if (kb < kblocks) {
        const __nv_bfloat16* a0 = sA + (size_t)g * pitch + kb * 32 + t * 8;
        const uint4 lo = *reinterpret_cast<const uint4*>(a0);
        const uint4 hi = *reinterpret_cast<const uint4*>(a0 + 8 * (size_t)pitch);
#pragma unroll
        for (int j = 0; j < T; ++j) {
          mma_16816(acc[j], lo.x, hi.x, lo.y, hi.y, w[j][i].x, w[j][i].y);
          mma_16816(acc[j], lo.z, hi.z, lo.w, hi.w, w[j][i].z, w[j][i].w);
        }
      }
    }
    // ---- 4. cross-warp sums + epilogue ----
#pragma unroll
    for (int j = 0; j < T; ++j) {
      float* r0 = red + ((size_t)(warp * T + j) * MK_ROWS + g) * 8 + 2 * t;
      r0[0] = acc[j][0], r0[1] = acc[j][1];
      r0[64] = acc[j][2], r0[65] = acc[j][3];  // row g + 8
    }
    __syncthreads();
    for (int e = threadIdx.x; e < T * MK_ROWS * 8; e += MK_THREADS) {
      const int j = e >> 7, row = (e >> 3) & 15, col = e & 7;
      const int tile = tile0 + j * G;
      const int n = tile * 8 + col;
      if (tile >= ntiles || n >= a.N || row >= p.B) continue;
      float v = 0.f;
#pragma unroll
      for (int wq = 0; wq < MK_WARPS; ++wq) v += red[(size_t)(wq * T + j) * MK_ROWS * 8 + (e & 127)];
      if (a.bias != nullptr) v += __ldg(a.bias + n);
      if (EPI == MK_EPI_QKV) {
        if (n < a.n_split)
          a.out_bf16[(long long)row * a.ldo + n] = __float2bfloat16_rn(v);
        else
          a.out2[(long long)row * a.ldo2 + a.pos_off + (n - a.n_split)] = __float2bfloat16_rn(v);
      } else if (EPI == MK_EPI_BF16) {
        a.out_bf16[(long long)row * a.ldo + n] = __float2bfloat16_rn(v);
      } else if (EPI == MK_EPI_GELU) {
        a.out_bf16[(long long)row * a.ldo + n] = __float2bfloat16_rn(gelu_erf_fast(v));
      } else {
        float* xp = a.xres + (long long)row * a.N + n;
        *xp = __ldcg(xp) + v;
      }
    }
    __syncthreads();  // red (and, in a second round, nothing else) is reused
  }
}

// L2 prefetch of the weights a CTA will request in its next linear phase (round 0 of linear_phase<T, KBW>): issued before the
// grid barrier in front of that phase, so that the requests after the barrier are L2 hits instead of DRAM round trips
template <int T, int KBW>
__device__ __noinline__ void prefetch_phase(const __nv_bfloat16* W, int N, int K, int flags) {
  if ((flags & 2) == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (N + 7) >> 3, kblocks = K >> 5;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const int tile = (int)blockIdx.x + j * (int)gridDim.x;
    const int nrow = tile * 8 + g;
    if (tile < ntiles && nrow < N) {
      const __nv_bfloat16* wrow = W + (long long)nrow * K + t * 8;
#pragma unroll
      for (int i = 0; i < KBW; ++i) {
        const int kb = warp + i * MK_WARPS;
        if (kb < kblocks) prefetch_l2(wrow + kb * 32);
      }
    }
  }
}

// self-attention of the step: one warp per (row, head); an 8-lane group owns a key (decode_attention_kernel of decode.cu)
__device__ __forceinline__ void self_attention_phase(const MegaParams& p, const __nv_bfloat16* kv, int Tk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int d = p.d;
  const long long kv_bs = (long long)p.S_max * 2 * d, kv_rs = 2 * d;
  for (int item = (int)blockIdx.x * MK_WARPS + warp; item < p.B * p.H; item += (int)gridDim.x * MK_WARPS) {
    const int b = item / p.H, h = item - b * p.H;
    float q[8];
    unpack8f(__ldcg(reinterpret_cast<const uint4*>(p.q + (long long)b * d + h * 64 + sub * 8)), q);
    const __nv_bfloat16* kb = kv + (long long)b * kv_bs + h * 64 + sub * 8;
    const __nv_bfloat16* vb = kb + d;
    float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    constexpr int UNROLL = 4;
    for (int kw = 0; kw < Tk; kw += 4 * UNROLL) {  // warp-uniform trip count (shuffles below)
      uint4 kk[UNROLL], vv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = kw + grp + u * 4;
        if (k < Tk) {
          kk[u] = __ldcg(reinterpret_cast<const uint4*>(kb + (long long)k * kv_rs));
          vv[u] = __ldcg(reinterpret_cast<const uint4*>(vb + (long long)k * kv_rs));
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = kw + grp + u * 4;
        const bool ok = k < Tk;
        float kf[8];
        float s = 0.f;
        if (ok) {
          unpack8f(kk[u], kf);
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ok) {
          const float mn = fmaxf(m, s);
          const float corr = __expf(m - mn);
          const float pw = __expf(s - mn);
          float vf[8];
          unpack8f(vv[u], vf);
          l = fmaf(l, corr, pw);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(o[i], corr, pw * vf[i]);
          m = mn;
        }
      }
    }
#pragma unroll
    for (int x = 8; x <= 16; x <<= 1) {  // merge the 4 groups of the warp
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
      const float inv = 1.0f / l;
      uint4 pk;
      pk.x = pack_bf16(o[0] * inv, o[1] * inv), pk.y = pack_bf16(o[2] * inv, o[3] * inv);
      pk.z = pack_bf16(o[4] * inv, o[5] * inv), pk.w = pack_bf16(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(p.ctx + (long long)b * d + h * 64 + sub * 8) = pk;
    }
  }
}

// cross-attention of the step.  The B x H (row, head) slabs of T keys are ONE key space of B * H * T keys dealt to the CTAs
// in equal contiguous ranges (one CTA per slab left 24 of 148 SMs with a third slab: 47 us per layer against 19 us of HBM
// time); a CTA walks the <= 3 slab segments of its range.  Per segment all 8 warps stream keys -- an 8-lane group owns a
// key (one 128-byte line each for k and v), two buffers of 6 keys per group -- and a slab that lies inside one range is
// finished by its CTA; a slab whose keys straddle a range boundary has two partials (o[64], m, l) and the CTA that arrives
// second (an arrival flag per slab) merges them, always in slab order.
constexpr int MK_PART = 68;  // floats per partial: o[64], m, l, arrival flag (slot 0 only), padding

__device__ __forceinline__ void cross_attention_phase(const MegaParams& p, const __nv_bfloat16* ckv, float* sm, int chunk) {
  float* sm_m = sm;                 // [8]
  float* sm_l = sm + MK_WARPS;      // [8]
  float* sm_o = sm + 2 * MK_WARPS;  // [8][64]
  int* sm_flag = reinterpret_cast<int*>(sm + 2 * MK_WARPS + MK_WARPS * 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int d = p.d, T = p.T;
  const long long total = (long long)p.B * p.H * T;
  long long k_lo = (long long)blockIdx.x * chunk;
  const long long k_end = min(k_lo + chunk, total);
  while (k_lo < k_end) {
    const long long item = k_lo / T;
    const int t_lo = (int)(k_lo - item * T);
    const int t_hi = (int)min((long long)T, k_end - item * T);
    const int b = (int)(item / p.H), h = (int)(item - (long long)b * p.H);
    float q[8];
    unpack8f(__ldcg(reinterpret_cast<const uint4*>(p.q + (long long)b * d + h * 64 + sub * 8)), q);
    const __nv_bfloat16* kb = ckv + item * T * 128 + sub * 8;
    const __nv_bfloat16* vb = kb + 64;
    float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    constexpr int U = 6;                    // keys in flight per 8-lane group and buffer
    constexpr int stride = MK_WARPS * 4;    // keys per step of all groups
    auto load = [&](int kw, uint4(&kk)[U], uint4(&vv)[U]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kw + grp + u * stride;
        if (k < t_hi) {
          kk[u] = ldg_stream(kb + (long long)k * 128);  // written by earlier LAUNCHES (the window's K/V projection)
          vv[u] = ldg_stream(vb + (long long)k * 128);
        }
      }
    };
    auto consume = [&](int kw, const uint4(&kk)[U], const uint4(&vv)[U]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kw + grp + u * stride;
        const bool ok = k < t_hi;  // uniform within the 8-lane group
        float kf[8];
        float s = 0.f;
        if (ok) {
          unpack8f(kk[u], kf);
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ok) {
          const float mn = fmaxf(m, s);
          const float corr = __expf(m - mn);
          const float pw = __expf(s - mn);
          float vf[8];
          unpack8f(vv[u], vf);
          l = fmaf(l, corr, pw);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(o[i], corr, pw * vf[i]);
          m = mn;
        }
      }
    };
    // two buffers: the loads of step i + 1 are in flight while step i is multiplied (2 x 6 keys x (k, v) x 16 B per lane:
    // 96 KB of loads in flight per SM)
    uint4 ka[U], va[U], kc[U], vc[U];
    int kw = t_lo + warp * 4;  // warp-uniform loop bounds (shuffles inside)
    if (kw < t_hi) load(kw, ka, va);
    while (kw < t_hi) {
      int kn = kw + stride * U;
      if (kn < t_hi) load(kn, kc, vc);
      consume(kw, ka, va);
      kw = kn;
      if (kw >= t_hi) break;
      kn = kw + stride * U;
      if (kn < t_hi) load(kn, ka, va);
      consume(kw, kc, vc);
      kw = kn;
    }
#pragma unroll
    for (int x = 8; x <= 16; x <<= 1) {  // merge the 4 groups of the warp
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
      for (int i = 0; i < 8; ++i) sm_o[warp * 64 + sub * 8 + i] = o[i];
    }
    __syncthreads();
    // the CTA's partial of this slab: (mt, lt, ot) in threads 0..63
    const bool whole = t_lo == 0 && t_hi == T;
    const int slot = t_lo == 0 ? 0 : 1;
    float* parts = p.attn_ws + item * 2 * MK_PART;
    float mt = -INFINITY, lt = 0.f, ot = 0.f;
    if (threadIdx.x < 64) {
#pragma unroll
      for (int w = 0; w < MK_WARPS; ++w) mt = fmaxf(mt, sm_m[w]);
#pragma unroll
      for (int w = 0; w < MK_WARPS; ++w) {
        const float c = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - mt);
        lt = fmaf(sm_l[w], c, lt);
        ot = fmaf(sm_o[w * 64 + threadIdx.x], c, ot);
      }
      if (whole) {
        p.ctx[(long long)b * d + h * 64 + threadIdx.x] = __float2bfloat16_rn(ot / lt);
      } else {
        float* mine = parts + slot * MK_PART;
        if (threadIdx.x == 0) mine[64] = mt, mine[65] = lt;
        mine[threadIdx.x] = ot;
      }
    }
    if (!whole) {
      // a slab shared by two CTAs: whoever arrives second merges (always in slot order: the result does not depend on who)
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        *sm_flag = atomicAdd(reinterpret_cast<int*>(parts + 66), 1);
      }
      __syncthreads();
      if (*sm_flag == 1) {
        if (threadIdx.x < 64) {
          __threadfence();
          const float* other = parts + (1 - slot) * MK_PART;
          const float mo = __ldcg(other + 64), lo = __ldcg(other + 65), oo = __ldcg(other + threadIdx.x);
          const float m0 = slot == 0 ? mt : mo, l0 = slot == 0 ? lt : lo, o0 = slot == 0 ? ot : oo;
          const float m1 = slot == 0 ? mo : mt, l1 = slot == 0 ? lo : lt, o1 = slot == 0 ? oo : ot;
          const float mm = fmaxf(m0, m1);
          const float c0 = __expf(m0 - mm), c1 = __expf(m1 - mm);
          const float ll = l0 * c0 + l1 * c1;
          p.ctx[(long long)b * d + h * 64 + threadIdx.x] = __float2bfloat16_rn((o0 * c0 + o1 * c1) / ll);
        }
        if (threadIdx.x == 0) *reinterpret_cast<int*>(parts + 66) = 0;  // ready for the next step
      }
    }
    __syncthreads();  // sm_* are rewritten by the next segment
    k_lo = item * T + t_hi;
  }
}

__global__ void __launch_bounds__(MK_THREADS, 1) decode_layers_kernel(const MegaParams p) {
  extern __shared__ __align__(16) unsigned char mk_smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(mk_smem);
  float* red = reinterpret_cast<float*>(mk_smem + (size_t)MK_ROWS * p.a_pitch * sizeof(__nv_bfloat16));
  // fp32 staging of the LayerNorm source rows [B, d]: the upper part of the A slab (LayerNorm phases have K = d <= 1280: their
  // bf16 A rows end below 16 * (1280 + 32) * 2 = 42 KB, the slab is sized for fc2's K = ffn)
  float* xstage = reinterpret_cast<float*>(mk_smem + (size_t)MK_ROWS * (p.d + 32) * sizeof(__nv_bfloat16));
  const int d = p.d;
  const int pos = __ldcg(p.pos);
  // cross-attention: keys per CTA (the key space of all (row, head) slabs split evenly)
  // (at least one whole slab, so that a slab has at most two parts)
  const int xchunk = max((int)(((long long)p.B * p.H * p.T + gridDim.x - 1) / gridDim.x), p.T);
  // barrier targets: every launch adds gridDim.x * nbar arrivals; the counter only ever grows
  const unsigned long long per_launch = (unsigned long long)gridDim.x * (unsigned long long)(1 + MK_NBAR_LAYER * p.L);
  __shared__ unsigned long long base_s;
  if (threadIdx.x == 0) base_s = (ld_acquire_u64(p.bar) / per_launch) * per_launch;
  __syncthreads();
  unsigned long long target = base_s;
  int stamp = 0;
  auto mark = [&]() {
    if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      p.prof[stamp] = tns;
    }
    ++stamp;
  };
  auto sync_grid = [&]() {
    mark();  // end of the phase's work on CTA 0
    target += gridDim.x;
    grid_barrier(p.bar, target);
    mark();  // barrier passed
  };
  mark();

  // ---- embedding: x[b] = embed_tokens[ids[b, pos]] + embed_positions[pos] ----
  for (int b = (int)blockIdx.x; b < p.B; b += (int)gridDim.x) {
    long long id = __ldcg(p.ids + (long long)b * p.ids_rs + pos);
    id = id < 0 ? 0 : (id >= p.vocab ? p.vocab - 1 : id);
    const float4* tr = reinterpret_cast<const float4*>(p.tok + id * d);
    const float4* pr = reinterpret_cast<const float4*>(p.pos_emb + (long long)pos * d);
    float4* xr = reinterpret_cast<float4*>(p.x + (long long)b * d);
    for (int i = threadIdx.x; i < d / 4; i += MK_THREADS) {
      const float4 a = __ldg(tr + i), c = __ldg(pr + i);
      xr[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
  }
  prefetch_phase<4, 5>(reinterpret_cast<const __nv_bfloat16*>(p.layers[0].wqkv), 3 * d, d, p.flags);
  sync_grid();

  for (int li = 0; li < p.L; ++li) {
    const dicow_decode_layer_t& w = p.layers[li];
    __nv_bfloat16* self_kv = reinterpret_cast<__nv_bfloat16*>(w.self_kv);
    LinearArgs a{};
    // q | k,v of the self-attention (k,v appended to the cache at the step's position)
    a.x = p.x, a.gamma = w.ln1_g, a.beta = w.ln1_b, a.A = nullptr;
    a.W = reinterpret_cast<const __nv_bfloat16*>(w.wqkv), a.bias = w.bqkv, a.N = 3 * d, a.K = d;
    a.out_bf16 = p.q, a.ldo = d, a.out2 = self_kv, a.ldo2 = (long long)p.S_max * 2 * d, a.pos_off = (long long)pos * 2 * d;
    a.n_split = d;
    linear_phase<4, 5, MK_EPI_QKV>(p, a, sA, red, xstage);
    sync_grid();
    self_attention_phase(p, self_kv, pos + 1);
    prefetch_phase<2, 5>(reinterpret_cast<const __nv_bfloat16*>(w.wo_self), d, d, p.flags);
    sync_grid();
    a = LinearArgs{};
    a.A = p.ctx, a.W = reinterpret_cast<const __nv_bfloat16*>(w.wo_self), a.bias = w.bo_self, a.N = d, a.K = d, a.xres = p.x;
    linear_phase<2, 5, MK_EPI_RESID>(p, a, sA, red, xstage);
    prefetch_phase<2, 5>(reinterpret_cast<const __nv_bfloat16*>(w.wq_cross), d, d, p.flags);
    sync_grid();
    a = LinearArgs{};
    a.x = p.x, a.gamma = w.ln2_g, a.beta = w.ln2_b;
    a.W = reinterpret_cast<const __nv_bfloat16*>(w.wq_cross), a.bias = w.bq_cross, a.N = d, a.K = d, a.out_bf16 = p.q, a.ldo = d;
    linear_phase<2, 5, MK_EPI_BF16>(p, a, sA, red, xstage);
    sync_grid();
    cross_attention_phase(p, reinterpret_cast<const __nv_bfloat16*>(w.cross_kv), red, xchunk);
    prefetch_phase<2, 5>(reinterpret_cast<const __nv_bfloat16*>(w.wo_cross), d, d, p.flags);
    sync_grid();
    a = LinearArgs{};
    a.A = p.ctx, a.W = reinterpret_cast<const __nv_bfloat16*>(w.wo_cross), a.bias = w.bo_cross, a.N = d, a.K = d, a.xres = p.x;
    linear_phase<2, 5, MK_EPI_RESID>(p, a, sA, red, xstage);
    prefetch_phase<MK_MAX_T, 5>(reinterpret_cast<const __nv_bfloat16*>(w.w1), p.ffn, d, p.flags);
    sync_grid();
    a = LinearArgs{};
    a.x = p.x, a.gamma = w.ln3_g, a.beta = w.ln3_b;
    a.W = reinterpret_cast<const __nv_bfloat16*>(w.w1), a.bias = w.b1, a.N = p.ffn, a.K = d, a.out_bf16 = p.h, a.ldo = p.ffn;
    linear_phase<MK_MAX_T, 5, MK_EPI_GELU>(p, a, sA, red, xstage);
    prefetch_phase<1, 20>(reinterpret_cast<const __nv_bfloat16*>(w.w2), d, p.ffn, p.flags);
    sync_grid();
    a = LinearArgs{};
    a.A = p.h, a.W = reinterpret_cast<const __nv_bfloat16*>(w.w2), a.bias = w.b2, a.N = d, a.K = p.ffn, a.xres = p.x;
    linear_phase<1, 20, MK_EPI_RESID>(p, a, sA, red, xstage);
    if (li + 1 < p.L) prefetch_phase<4, 5>(reinterpret_cast<const __nv_bfloat16*>(p.layers[li + 1].wqkv), 3 * d, d, p.flags);
    sync_grid();
  }
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_decode_layers(dicow_handle_t h, const dicow_decode_layers_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_decode_layers_args_t), "dicow_decode_layers: bad args struct");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->B <= MK_ROWS, "dicow_decode_layers: 1 <= B <= %d rows (got %d)", MK_ROWS, a->B);
  DICOW_REQUIRE(ctx, a->d >= 64 && a->d <= 1280 && (a->d % 64) == 0 && a->H * 64 == a->d,
                "dicow_decode_layers: d_model must be a multiple of 64 up to 1280 with head_dim 64 (d=%d, H=%d)", a->d, a->H);
  DICOW_REQUIRE(ctx, a->ffn >= 64 && a->ffn <= 5120 && (a->ffn % 32) == 0, "dicow_decode_layers: ffn must be a multiple of 32 up to 5120");
  DICOW_REQUIRE(ctx, a->L >= 1 && a->layers && a->ids && a->embed_tokens && a->embed_positions && a->pos && a->x && a->q && a->ctx &&
                         a->hidden && a->barrier && a->attn_workspace,
                "dicow_decode_layers: null argument");
  DICOW_REQUIRE(ctx, a->T >= 1 && a->S_max >= 1, "dicow_decode_layers: bad cache sizes");
  MegaParams p{};
  p.B = a->B, p.d = a->d, p.H = a->H, p.ffn = a->ffn, p.L = a->L, p.T = a->T, p.S_max = a->S_max, p.vocab = a->vocab;
  p.layers = a->layers;
  p.ids = reinterpret_cast<const long long*>(a->ids), p.ids_rs = a->ids_row_stride;
  p.tok = a->embed_tokens, p.pos_emb = a->embed_positions, p.pos = a->pos;
  p.x = a->x, p.q = reinterpret_cast<__nv_bfloat16*>(a->q), p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx);
  p.h = reinterpret_cast<__nv_bfloat16*>(a->hidden);
  p.attn_ws = a->attn_workspace;
  p.bar = reinterpret_cast<unsigned long long*>(a->barrier);
  p.eps = a->eps;
  p.flags = a->flags;
  p.prof = reinterpret_cast<long long*>(ctx->attn_prof);
  const int kmax = a->ffn > a->d ? a->ffn : a->d;
  p.a_pitch = kmax + 32;  // row pitch = 64 (mod 128) bytes: conflict-free fragment loads
  size_t slab = (size_t)MK_ROWS * p.a_pitch * 2;  // A slab (bf16, K = ffn) ...
  const size_t ln_need = (size_t)MK_ROWS * (a->d + 32) * 2 + (size_t)MK_ROWS * a->d * 4;  // ... or LayerNorm: bf16 rows + fp32 stage
  if (ln_need > slab) slab = ln_need, p.a_pitch = (int)(slab / (MK_ROWS * 2));
  const size_t smem = (size_t)MK_ROWS * p.a_pitch * 2 + (size_t)MK_WARPS * MK_MAX_T * MK_ROWS * 8 * sizeof(float);
  DICOW_REQUIRE(ctx, smem <= (size_t)ctx->max_smem_optin, "dicow_decode_layers: %zu bytes of shared memory needed", smem);
  static DeviceHighWater attr_smem;
  if (attr_smem.raise(ctx, smem))
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(decode_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // One CTA per SM; the grid barrier needs every CTA running.  The shared-memory footprint admits exactly one CTA per SM and
  // the grid is the SM count, so all CTAs become resident as soon as the kernels ahead of this one in the stream drain: a plain
  // launch is enough.  DICOW_MEGA_COOPERATIVE=1 asks the driver to verify co-residency (cooperative launch).
  // KNOWN ISSUE (driver 580.159, unresolved): in ONE process history -- all of tests/test_gpu_training.py followed by
  // tests/test_gpu_turbo_parity.py, no smaller subset of the training tests reproduces it -- the first launch of this kernel at
  // large-v3-turbo dimensions dies with SIGSEGV inside libcuda's cuLaunchKernelEx (with and without the cooperative attribute,
  // with eager module loading, with the kernel pre-loaded).  The whole GPU suite in its normal order, and either file alone, pass.
  // The kernel is experimental and opt-in (DiCoW.decode_megakernel / DICOW_DECODE_MEGA=1), see DESIGN.md section 4.2.
  static const int cooperative = [] {
    const char* e = getenv("DICOW_MEGA_COOPERATIVE");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctx->num_sms), cfg.blockDim = dim3(MK_THREADS), cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = cooperative ? 1 : 0;
  DICOW_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, decode_layers_kernel, p));
  return DICOW_OK;
}
