// elementwise.cu -- HBM-bound kernels of the encoder: fused FDDT + LayerNorm (warp-shuffle reductions),
// mel-feature re-layout for the implicit-GEMM conv stem.
//
// FDDT (Frame-level Diarization-Dependent Transformation), reference src/models/dicow/FDDT.py:41-63:
//     x'[b,t,:] = sum_{c in S,T,N,O} stno[b,c,t] * (w_c (.) x[b,t,:] + b_c)
// In the reference this is ~15 eager element-wise launches per layer, each a full read+write of the residual
// stream (encoder.py:205-206), followed by nn.LayerNorm (HF:modeling_whisper.py:393).  Here one kernel reads the
// fp32 residual row once, applies FDDT, writes it back, and emits LayerNorm(x') in bf16 for the QKV GEMM.
#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

struct FddtLnParams {
  float* x;  // [rows, d] fp32 residual stream (updated in place when FDDT is applied)
  int rows, d, T;
  // FDDT (nullable => skipped)
  const float* stno;  // [B, 4, T]
  long long stno_bs;
  const float* fddt_w;  // [4, d] (S,T,N,O) or NULL for bias-only FDDT
  const float* fddt_b;  // [4, d]
  // LayerNorm (nullable gamma => no LN outputs)
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* ln_bf16;  // [rows, d] or NULL
  float* ln_f32;           // [rows, d] or NULL
  __nv_bfloat16* x_bf16;   // [rows, d] bf16 copy of x' or NULL
  // pending residual updates (bf16 GEMM outputs: out_proj / fc2), added before FDDT: x' = FDDT(x + delta1 + delta2)
  const __nv_bfloat16* delta1;
  const __nv_bfloat16* delta2;
  int store_x;  // write x' back (to xs)
  float* xs;    // where x' goes: x itself, or a separate [rows, d] buffer (the training forward keeps x as a saved activation)
};

// one warp per row; VPL float4 per lane (d <= 128 * VPL)
template <int VPL>
__global__ void __launch_bounds__(256) fddt_ln_kernel(const FddtLnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= p.rows) return;
  const int nvec = p.d >> 2;
  float4 v[VPL];
  float4* xrow = reinterpret_cast<float4*>(p.x + (long long)row * p.d);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < nvec) ? xrow[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int dd = 0; dd < 2; ++dd) {
    const __nv_bfloat16* dl = dd == 0 ? p.delta1 : p.delta2;
    if (dl == nullptr) continue;
    const uint2* drow = reinterpret_cast<const uint2*>(dl + (long long)row * p.d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const uint2 u = __ldg(drow + c);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        v[i].x += a.x, v[i].y += a.y, v[i].z += b2.x, v[i].w += b2.y;
      }
    }
  }
  if (p.stno != nullptr) {
    const int b = row / p.T, t = row - b * p.T;
    float m[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) m[c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.T + t);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nvec) {
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f), bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.fddt_w != nullptr) w = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (p.fddt_w != nullptr) {
            const float4 wc = __ldg(reinterpret_cast<const float4*>(p.fddt_w + (long long)c * p.d) + c4);
            w.x = fmaf(m[c], wc.x, w.x), w.y = fmaf(m[c], wc.y, w.y), w.z = fmaf(m[c], wc.z, w.z),
            w.w = fmaf(m[c], wc.w, w.w);
          }
          const float4 bc = __ldg(reinterpret_cast<const float4*>(p.fddt_b + (long long)c * p.d) + c4);
          bb.x = fmaf(m[c], bc.x, bb.x), bb.y = fmaf(m[c], bc.y, bb.y), bb.z = fmaf(m[c], bc.z, bb.z),
          bb.w = fmaf(m[c], bc.w, bb.w);
        }
        v[i].x = fmaf(v[i].x, w.x, bb.x), v[i].y = fmaf(v[i].y, w.y, bb.y);
        v[i].z = fmaf(v[i].z, w.z, bb.z), v[i].w = fmaf(v[i].w, w.w, bb.w);
      }
    }
  }
  if (p.store_x) {
    float4* xsrow = reinterpret_cast<float4*>(p.xs + (long long)row * p.d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nvec) xsrow[c4] = v[i];
    }
  }
  if (p.x_bf16 != nullptr) {
    uint2* o = reinterpret_cast<uint2*>(p.x_bf16 + (long long)row * p.d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nvec) o[c4] = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
    }
  }
  if (p.gamma == nullptr) return;
  // LayerNorm over d: two-pass (mean, then centred variance) in fp32, eps inside the sqrt (nn.LayerNorm)
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);  // padded lanes hold zeros
  const float mean = warp_sum(s) / (float)p.d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + 32 * i < nvec) {
      const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b2 * b2) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.d + p.eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    if (c4 < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + c4);
      const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta) + c4);
      float4 y;
      y.x = fmaf((v[i].x - mean) * rstd, g.x, be.x);
      y.y = fmaf((v[i].y - mean) * rstd, g.y, be.y);
      y.z = fmaf((v[i].z - mean) * rstd, g.z, be.z);
      y.w = fmaf((v[i].w - mean) * rstd, g.w, be.w);
      if (p.ln_bf16 != nullptr)
        reinterpret_cast<uint2*>(p.ln_bf16 + (long long)row * p.d)[c4] =
            make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      if (p.ln_f32 != nullptr) reinterpret_cast<float4*>(p.ln_f32 + (long long)row * p.d)[c4] = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm of a handful of rows (the decode step: one row per sequence).  One CTA of 128 threads per row, at most
// three float4 per thread (d <= 1536), a loop-free body of ~150 instructions: the kernel runs once per launch per
// warp, so a 10x unrolled body (fddt_ln_kernel<10>) is bound by instruction fetch -- measured 7.3 us for 16 rows.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_128(float v, float* sm) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  const float t = (sm[0] + sm[1]) + (sm[2] + sm[3]);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(128) ln_rows_kernel(const float* __restrict__ x, int d, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps,
                                                      __nv_bfloat16* __restrict__ ln_bf16, float* __restrict__ ln_f32) {
  __shared__ float sm[4];
  const int row = blockIdx.x, tid = threadIdx.x, nvec = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * d);
  // gamma / beta are requested before the row (loads are not moved across the barriers of the block sums by the
  // compiler: after them they were a third memory round trip in a kernel that is nothing but latency) -- and before the
  // dependency wait of a decode-step launch: they are immutable
  float4 gv[3], bv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c4 = tid + 128 * i;
    gv[i] = bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < nvec) {
      gv[i] = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      bv[i] = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    }
  }
  griddep_launch();
  griddep_wait();
  float4 v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 128 * i;
    v[i] = c < nvec ? __ldcg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = block_sum_128(s, sm) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (tid + 128 * i < nvec) {
      const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b2 * b2) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(block_sum_128(q, sm) / (float)d + eps);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c4 = tid + 128 * i;
    if (c4 < nvec) {
      const float4 g = gv[i];
      const float4 be = bv[i];
      float4 y;
      y.x = fmaf((v[i].x - mean) * rstd, g.x, be.x);
      y.y = fmaf((v[i].y - mean) * rstd, g.y, be.y);
      y.z = fmaf((v[i].z - mean) * rstd, g.z, be.z);
      y.w = fmaf((v[i].w - mean) * rstd, g.w, be.w);
      if (ln_bf16 != nullptr)
        reinterpret_cast<uint2*>(ln_bf16 + (long long)row * d)[c4] = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      if (ln_f32 != nullptr) reinterpret_cast<float4*>(ln_f32 + (long long)row * d)[c4] = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// full-matrix FDDT (CustomLinear per class, src/models/dicow/layers.py:7-47 + FDDT.py:52-62): the four class transforms are
// ONE GEMM y = x [W_S; W_T; W_N; W_O]^T + [b_S; ...] (bf16 [rows, 4 d], like the reference's autocast Linear outputs);
// this kernel forms the mask-weighted sum in fp32:  x'[r, :] = sum_c stno[b, c, t] * y[r, c d : (c + 1) d]  (+ pos[t, :]).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fddt_full_combine_kernel(const __nv_bfloat16* __restrict__ y, long long ldy,
                                                                const float* __restrict__ stno, long long stno_bs, int T,
                                                                int rows, int d, const float* __restrict__ pos,
                                                                float* __restrict__ x) {
  const int nvec = d >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / nvec), c4 = (int)(i - (long long)r * nvec);
    const int b = r / T, t = r - b * T;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float m = __ldg(stno + (long long)b * stno_bs + (long long)c * T + t);
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(y + (long long)r * ldy + (long long)c * d) + c4);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 e = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      acc.x = fmaf(m, a.x, acc.x), acc.y = fmaf(m, a.y, acc.y), acc.z = fmaf(m, e.x, acc.z), acc.w = fmaf(m, e.y, acc.w);
    }
    if (pos != nullptr) {
      const float4 pv = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * d) + c4);
      acc.x += pv.x, acc.y += pv.y, acc.z += pv.z, acc.w += pv.w;
    }
    reinterpret_cast<float4*>(x + (long long)r * d)[c4] = acc;
  }
}

// TMA-pipelined variant (default): one producer warp streams rows (x fp32 + up to two bf16 deltas) into a 16-stage
// shared-memory ring with cp.async.bulk + mbarrier transaction counts; 16 consumer warps each take a row from the ring,
// so the HBM latency is hidden by the ring depth (160 KB in flight per SM) instead of by occupancy -- the register-
// resident warp-per-row kernel runs at 16 warps / SM and 40-60 % of the HBM roofline (ncu r01).  The FDDT tables sit in
// shared memory (read 8 values per element from there instead of through L1 tags).
constexpr int LT_NW = 16;   // consumer warps
constexpr int LT_NST = 16;  // ring stages (one row each), a multiple of LT_NW so a stage always belongs to one warp

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int VPL>
__global__ void __launch_bounds__((LT_NW + 1) * 32, 1) fddt_ln_tma_kernel(const FddtLnParams p) {
  extern __shared__ __align__(128) uint8_t lsm[];
  const int d = p.d;
  const uint32_t xb = d * 4, db = d * 2, stage_bytes = xb + 2 * db;
  uint8_t* ring = lsm;                                                   // [LT_NST][x | d1 | d2]
  float* tab = reinterpret_cast<float*>(lsm + LT_NST * stage_bytes);     // [8][d]: w S,T,N,O then b S,T,N,O
  float* gb = tab + (p.stno != nullptr ? 8 * d : 0);                     // [2][d]: LayerNorm gamma, beta
  uint64_t* full = reinterpret_cast<uint64_t*>(gb + (p.gamma != nullptr ? 2 * d : 0));
  uint64_t* empty = full + LT_NST;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LT_NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  if (p.stno != nullptr) {
    for (int i = threadIdx.x; i < 4 * nvec; i += blockDim.x) {
      reinterpret_cast<float4*>(tab)[i] =
          p.fddt_w != nullptr ? __ldg(reinterpret_cast<const float4*>(p.fddt_w) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      reinterpret_cast<float4*>(tab)[4 * nvec + i] = __ldg(reinterpret_cast<const float4*>(p.fddt_b) + i);
    }
  }
  if (p.gamma != nullptr) {  // the affine parameters once per CTA: re-reading them per row through L1 cost as many L1 bytes
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {  // as the row itself brings in from HBM (ncu r01: l1tex 56 %)
      reinterpret_cast<float4*>(gb)[i] = __ldg(reinterpret_cast<const float4*>(p.gamma) + i);
      reinterpret_cast<float4*>(gb)[nvec + i] = __ldg(reinterpret_cast<const float4*>(p.beta) + i);
    }
  }
  __syncthreads();
  // rows of this CTA: row = i * gridDim.x + blockIdx.x
  const int n_local = (p.rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t tx = xb + (p.delta1 != nullptr ? db : 0) + (p.delta2 != nullptr ? db : 0);

  if (warp == LT_NW) {
    // ===================== producer =====================
    for (int i = 0; i < n_local; ++i) {
      const int st = i % LT_NST;
      mbar_wait(&empty[st], ((i / LT_NST) & 1) ^ 1);
      if (elect_one()) {
        const long long row = (long long)i * gridDim.x + blockIdx.x;
        uint8_t* dst = ring + st * stage_bytes;
        mbar_arrive_expect_tx(&full[st], tx);
        bulk_load(dst, p.x + row * d, xb, &full[st]);
        if (p.delta1 != nullptr) bulk_load(dst + xb, p.delta1 + row * d, db, &full[st]);
        if (p.delta2 != nullptr) bulk_load(dst + xb + db, p.delta2 + row * d, db, &full[st]);
      }
      __syncwarp();
    }
    return;
  }
  // ===================== consumers: one row per warp per turn =====================
  // the STNO mask of a row (4 scattered scalars, an L2 round trip) is requested one turn ahead
  float m_next[4] = {0.f, 0.f, 0.f, 0.f};
  auto load_mask = [&](int i, float(&m)[4]) {
    if (p.stno != nullptr && i < n_local) {
      const int row = i * (int)gridDim.x + (int)blockIdx.x;
      const int b = row / p.T, t = row - b * p.T;
#pragma unroll
      for (int c = 0; c < 4; ++c) m[c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.T + t);
    }
  };
  load_mask(warp, m_next);
  for (int i = warp; i < n_local; i += LT_NW) {
    const int st = i % LT_NST;
    const int row = i * (int)gridDim.x + (int)blockIdx.x;
    float m[4] = {m_next[0], m_next[1], m_next[2], m_next[3]};
    load_mask(i + LT_NW, m_next);
    mbar_wait(&full[st], (i / LT_NST) & 1);
    const uint8_t* src = ring + st * stage_bytes;
    float4 v[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = lane + 32 * k;
      v[k] = (c < nvec) ? reinterpret_cast<const float4*>(src)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int dd = 0; dd < 2; ++dd) {
      if ((dd == 0 ? p.delta1 : p.delta2) == nullptr) continue;
      const uint2* dr = reinterpret_cast<const uint2*>(src + xb + dd * db);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c = lane + 32 * k;
        if (c < nvec) {
          const uint2 u = dr[c];
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
          const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
          v[k].x += a.x, v[k].y += a.y, v[k].z += b2.x, v[k].w += b2.y;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // the row is in registers: the stage may be refilled
    if (p.stno != nullptr) {
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < nvec) {
          float4 w = make_float4(0.f, 0.f, 0.f, 0.f), bb = w;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 wc = reinterpret_cast<const float4*>(tab + c * d)[c4];
            const float4 bc = reinterpret_cast<const float4*>(tab + (4 + c) * d)[c4];
            w.x = fmaf(m[c], wc.x, w.x), w.y = fmaf(m[c], wc.y, w.y), w.z = fmaf(m[c], wc.z, w.z), w.w = fmaf(m[c], wc.w, w.w);
            bb.x = fmaf(m[c], bc.x, bb.x), bb.y = fmaf(m[c], bc.y, bb.y), bb.z = fmaf(m[c], bc.z, bb.z),
            bb.w = fmaf(m[c], bc.w, bb.w);
          }
          if (p.fddt_w == nullptr) w = make_float4(1.f, 1.f, 1.f, 1.f);  // bias-only FDDT (FDDT.py:43-51): h += sum_c m_c b_c
          v[k].x = fmaf(v[k].x, w.x, bb.x), v[k].y = fmaf(v[k].y, w.y, bb.y);
          v[k].z = fmaf(v[k].z, w.z, bb.z), v[k].w = fmaf(v[k].w, w.w, bb.w);
        }
      }
    }
    if (p.store_x) {
      float4* xrow = reinterpret_cast<float4*>(p.xs + (long long)row * d);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < nvec) xrow[c4] = v[k];
      }
    }
    if (p.x_bf16 != nullptr) {
      uint2* o = reinterpret_cast<uint2*>(p.x_bf16 + (long long)row * d);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < nvec) o[c4] = make_uint2(pack_bf16(v[k].x, v[k].y), pack_bf16(v[k].z, v[k].w));
      }
    }
    if (p.gamma == nullptr) continue;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);  // padded lanes hold zeros
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < nvec) {
        const float a = v[k].x - mean, b2 = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
        q += (a * a + b2 * b2) + (c * c + e * e);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + p.eps);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < nvec) {
        const float4 g = reinterpret_cast<const float4*>(gb)[c4];
        const float4 be = reinterpret_cast<const float4*>(gb)[nvec + c4];
        float4 y;
        y.x = fmaf((v[k].x - mean) * rstd, g.x, be.x);
        y.y = fmaf((v[k].y - mean) * rstd, g.y, be.y);
        y.z = fmaf((v[k].z - mean) * rstd, g.z, be.z);
        y.w = fmaf((v[k].w - mean) * rstd, g.w, be.w);
        if (p.ln_bf16 != nullptr)
          reinterpret_cast<uint2*>(p.ln_bf16 + (long long)row * d)[c4] =
              make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        if (p.ln_f32 != nullptr) reinterpret_cast<float4*>(p.ln_f32 + (long long)row * d)[c4] = y;
      }
    }
  }
}

template <int VPL>
int launch_fddt_ln_tma(dicow_ctx* ctx, const FddtLnParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)LT_NST * 8 * p.d + (p.stno != nullptr ? (size_t)32 * p.d : 0) +
                      (p.gamma != nullptr ? (size_t)8 * p.d : 0) + 2 * LT_NST * 8 + 128;
  auto kfn = fddt_ln_tma_kernel<VPL>;
  static DeviceHighWater attr_smem;
  if (attr_smem.raise(ctx, smem)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid = p.rows < ctx->num_sms ? p.rows : ctx->num_sms;
  kfn<<<grid, (LT_NW + 1) * 32, smem, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

// Column-owner variant (default): thread t owns the float4 column t of every row its CTA processes, so the FDDT tables
// (4 classes x {w, b}) and the LayerNorm affine parameters of that column live in registers for the whole kernel --
// the warp-per-row kernel above re-reads 8 table values per element through L1 (40 KB per row against 10 KB of HBM
// data; ncu r01: l1tex 60 %, DRAM 40 %).  A CTA handles ROWS rows per iteration; the two LayerNorm reductions go through
// shared memory once per iteration for all ROWS rows together.
constexpr int LN_ROWS = 4;

__global__ void __launch_bounds__(512) fddt_ln_cols_kernel(const FddtLnParams p, const int row_groups) {
  __shared__ float red[2][32][LN_ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int nvec = p.d >> 2;
  const bool own = tid < nvec;  // threads past d/4 only take part in the reductions
  float4 tw[4], tb[4], g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
  const bool has_fddt = p.stno != nullptr;
  if (own) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tw[c] = make_float4(1.f, 1.f, 1.f, 1.f);
      tb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_fddt) {
        if (p.fddt_w != nullptr) tw[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_w + (long long)c * p.d) + tid);
        tb[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_b + (long long)c * p.d) + tid);
      }
    }
    if (p.gamma != nullptr) {
      g4 = __ldg(reinterpret_cast<const float4*>(p.gamma) + tid);
      b4 = __ldg(reinterpret_cast<const float4*>(p.beta) + tid);
    }
  }
  const float inv_d = 1.0f / (float)p.d;
  for (int grp = blockIdx.x; grp < row_groups; grp += gridDim.x) {
    const int row0 = grp * LN_ROWS;
    float4 v[LN_ROWS];
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) {
      const int row = row0 + r;
      v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (own && row < p.rows) v[r] = reinterpret_cast<const float4*>(p.x + (long long)row * p.d)[tid];
    }
#pragma unroll
    for (int dd = 0; dd < 2; ++dd) {
      const __nv_bfloat16* dl = dd == 0 ? p.delta1 : p.delta2;
      if (dl == nullptr) continue;
#pragma unroll
      for (int r = 0; r < LN_ROWS; ++r) {
        const int row = row0 + r;
        if (own && row < p.rows) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(dl + (long long)row * p.d) + tid);
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
          const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
          v[r].x += a.x, v[r].y += a.y, v[r].z += c.x, v[r].w += c.y;
        }
      }
    }
    if (has_fddt) {
#pragma unroll
      for (int r = 0; r < LN_ROWS; ++r) {
        const int row = row0 + r;
        if (own && row < p.rows) {
          const int b = row / p.T, t = row - b * p.T;
          float4 w = make_float4(0.f, 0.f, 0.f, 0.f), bb = w;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float m = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.T + t);
            w.x = fmaf(m, tw[c].x, w.x), w.y = fmaf(m, tw[c].y, w.y), w.z = fmaf(m, tw[c].z, w.z), w.w = fmaf(m, tw[c].w, w.w);
            bb.x = fmaf(m, tb[c].x, bb.x), bb.y = fmaf(m, tb[c].y, bb.y), bb.z = fmaf(m, tb[c].z, bb.z),
            bb.w = fmaf(m, tb[c].w, bb.w);
          }
          if (p.fddt_w == nullptr) w = make_float4(1.f, 1.f, 1.f, 1.f);  // bias-only FDDT
          v[r].x = fmaf(v[r].x, w.x, bb.x), v[r].y = fmaf(v[r].y, w.y, bb.y);
          v[r].z = fmaf(v[r].z, w.z, bb.z), v[r].w = fmaf(v[r].w, w.w, bb.w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) {
      const int row = row0 + r;
      if (own && row < p.rows) {
        if (p.store_x) reinterpret_cast<float4*>(p.xs + (long long)row * p.d)[tid] = v[r];
        if (p.x_bf16 != nullptr)
          reinterpret_cast<uint2*>(p.x_bf16 + (long long)row * p.d)[tid] =
              make_uint2(pack_bf16(v[r].x, v[r].y), pack_bf16(v[r].z, v[r].w));
      }
    }
    if (p.gamma == nullptr) continue;  // uniform
    // LayerNorm over d: mean, then centred variance (two-pass in fp32, eps inside the sqrt like nn.LayerNorm)
    float s[LN_ROWS];
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) s[r] = warp_sum((v[r].x + v[r].y) + (v[r].z + v[r].w));
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < LN_ROWS; ++r) red[0][warp][r] = s[r];
    }
    __syncthreads();
    float mean[LN_ROWS], q[LN_ROWS];
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += red[0][w][r];
      mean[r] = t * inv_d;
      float a = 0.f;
      if (own) {
        const float e0 = v[r].x - mean[r], e1 = v[r].y - mean[r], e2 = v[r].z - mean[r], e3 = v[r].w - mean[r];
        a = (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      q[r] = warp_sum(a);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < LN_ROWS; ++r) red[1][warp][r] = q[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) {
      const int row = row0 + r;
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += red[1][w][r];
      const float rstd = rsqrtf(t * inv_d + p.eps);
      if (own && row < p.rows) {
        float4 y;
        y.x = fmaf((v[r].x - mean[r]) * rstd, g4.x, b4.x);
        y.y = fmaf((v[r].y - mean[r]) * rstd, g4.y, b4.y);
        y.z = fmaf((v[r].z - mean[r]) * rstd, g4.z, b4.z);
        y.w = fmaf((v[r].w - mean[r]) * rstd, g4.w, b4.w);
        if (p.ln_bf16 != nullptr)
          reinterpret_cast<uint2*>(p.ln_bf16 + (long long)row * p.d)[tid] =
              make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        if (p.ln_f32 != nullptr) reinterpret_cast<float4*>(p.ln_f32 + (long long)row * p.d)[tid] = y;
      }
    }
    // red[0] is rewritten only after the next iteration's loads and a full pass of arithmetic, but a fast warp could
    // still race a slow one reading red[1]: the first __syncthreads of the next iteration orders red[0] writes after
    // these reads only for red[0]; keep the two buffers disjoint and re-synchronise before reuse
    __syncthreads();
  }
}

// input_features fp32 [B, C, F] -> channels-last bf16 [B, F + 2, C] with zero rows 0 and F+1
// (the zero-padded buffer conv1's implicit GEMM reads; reference: encoder.py:167 nn.Conv1d(padding=1))
__global__ void __launch_bounds__(256) features_to_cl_kernel(const float* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int C, int F) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (long long)b * C * F;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, f = f0 + tx;
    tile[ty + j][tx] = (c < C && f < F) ? src[(long long)c * F + f] : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* dst = out + (long long)b * (F + 2) * C;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int f = f0 + ty + j, c = c0 + tx;
    if (f < F && c < C) dst[(long long)(f + 1) * C + c] = __float2bfloat16_rn(tile[tx][ty + j]);
  }
  // zero pad rows (done by the blocks of the first f-tile)
  if (blockIdx.x == 0) {
    const int c = c0 + tx;
    if (ty == 0 && c < C) {
      dst[c] = __float2bfloat16_rn(0.f);
      dst[(long long)(F + 1) * C + c] = __float2bfloat16_rn(0.f);
    }
  }
}

// zero the two pad rows of a channels-last [B, T + 2, C] bf16 buffer (conv1 writes rows 1..T)
__global__ void zero_pad_rows_kernel(__nv_bfloat16* buf, int T, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  __nv_bfloat16* base = buf + (long long)b * (T + 2) * C;
  base[c] = __float2bfloat16_rn(0.f);
  base[(long long)(T + 1) * C + c] = __float2bfloat16_rn(0.f);
}

// fp32 -> bf16 cast (weights preparation, small tensors)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_fddt_layernorm(dicow_handle_t h, const dicow_fddt_ln_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_fddt_ln_args_t), "dicow_fddt_layernorm: bad args struct");
  DICOW_REQUIRE(ctx, a->x != nullptr && a->rows >= 1 && a->d >= 4 && (a->d % 4) == 0 && a->d <= 2048 && (a->d <= 1280 || a->d >= 128),
                "dicow_fddt_layernorm: need x, rows>=1, d%%4==0, d<=2048 (got rows=%d d=%d)", a->rows, a->d);
  DICOW_REQUIRE(ctx, a->stno == nullptr || (a->fddt_b != nullptr && a->T >= 1 && (a->rows % a->T) == 0),
                "dicow_fddt_layernorm: FDDT needs fddt_b and rows %% T == 0");
  DICOW_REQUIRE(ctx, a->gamma == nullptr || a->beta != nullptr, "dicow_fddt_layernorm: gamma without beta");
  FddtLnParams p{};
  p.x = a->x, p.rows = a->rows, p.d = a->d, p.T = a->T > 0 ? a->T : a->rows;
  p.stno = a->stno, p.stno_bs = a->stno_batch_stride, p.fddt_w = a->fddt_w, p.fddt_b = a->fddt_b;
  p.gamma = a->gamma, p.beta = a->beta, p.eps = a->eps;
  p.ln_bf16 = reinterpret_cast<__nv_bfloat16*>(a->ln_out_bf16), p.ln_f32 = a->ln_out_f32;
  p.x_bf16 = reinterpret_cast<__nv_bfloat16*>(a->x_out_bf16);
  p.delta1 = reinterpret_cast<const __nv_bfloat16*>(a->delta1_bf16);
  p.delta2 = reinterpret_cast<const __nv_bfloat16*>(a->delta2_bf16);
  p.store_x = a->store_x;
  p.xs = a->x_out != nullptr ? a->x_out : a->x;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // default: TMA-pipelined kernel (needs 16-byte rows: d % 8 == 0); flags bit 0 -> warp-per-row, bit 1 -> column-owner.
  // A handful of rows (the decode step: one row per sequence) cannot amortise the ring's set-up (measured 7.1 us for
  // 16 rows): they take the warp-per-row kernel, two rows per CTA.
  const bool few_rows = a->rows <= 128;
  if (few_rows && a->gamma != nullptr && a->stno == nullptr && a->delta1_bf16 == nullptr && a->delta2_bf16 == nullptr &&
      a->x_out_bf16 == nullptr && a->d <= 1536 && !(a->flags & 3) && (reinterpret_cast<uintptr_t>(a->x) % 16) == 0) {
    DICOW_CUDA_OK(ctx, launch_step_kernel(ln_rows_kernel, dim3(a->rows), dim3(128), 0, stream, 1u, a->x, a->d, a->gamma, a->beta,
                                          a->eps, reinterpret_cast<__nv_bfloat16*>(a->ln_out_bf16), a->ln_out_f32));
    return DICOW_OK;
  }
  if (!few_rows && (a->d % 8) == 0 && a->d <= 1280 && !(a->flags & 3) && (reinterpret_cast<uintptr_t>(a->x) % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(a->delta1_bf16) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->delta2_bf16) % 16) == 0) {
    switch (ceil_div(a->d, 128)) {
#define DICOW_CASE(V) \
  case V: return launch_fddt_ln_tma<V>(ctx, p, stream);
      DICOW_CASE(1) DICOW_CASE(2) DICOW_CASE(3) DICOW_CASE(4) DICOW_CASE(5) DICOW_CASE(6) DICOW_CASE(7) DICOW_CASE(8)
      DICOW_CASE(9) DICOW_CASE(10)
#undef DICOW_CASE
      default: break;
    }
  }
  if (a->d >= 128 && a->d <= 2048 && (a->flags & 2)) {  // column-owner kernel (comparison)
    const int threads = ceil_div(a->d / 4, 32) * 32;
    const int groups = ceil_div(a->rows, LN_ROWS);
    const int per_sm = threads >= 512 ? 2 : (2048 / threads > 6 ? 6 : 2048 / threads);
    const int grid2 = groups < ctx->num_sms * per_sm ? groups : ctx->num_sms * per_sm;
    fddt_ln_cols_kernel<<<grid2, threads, 0, stream>>>(p, groups);
    DICOW_CUDA_OK(ctx, cudaGetLastError());
    return DICOW_OK;
  }
  const int rows_per_block = few_rows ? 2 : 8;
  const int grid = ceil_div(a->rows, rows_per_block);
  const int vpl = ceil_div(a->d, 128);
  switch (vpl) {
#define DICOW_CASE(V) \
  case V: fddt_ln_kernel<V><<<grid, rows_per_block * 32, 0, stream>>>(p); break;
    DICOW_CASE(1) DICOW_CASE(2) DICOW_CASE(3) DICOW_CASE(4) DICOW_CASE(5) DICOW_CASE(6) DICOW_CASE(7) DICOW_CASE(8)
    DICOW_CASE(9) DICOW_CASE(10)
#undef DICOW_CASE
    default: return set_error(ctx, DICOW_ERR_UNSUPPORTED, "dicow_fddt_layernorm: d=%d unsupported", a->d);
  }
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_fddt_full_combine(dicow_handle_t h, const void* y_bf16, int64_t ldy, const float* stno,
                                       int64_t stno_batch_stride, int T, int rows, int d, const float* pos, float* x,
                                       void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, y_bf16 && stno && x && rows >= 1 && T >= 1 && (rows % T) == 0 && d >= 4 && (d % 4) == 0 && (ldy % 4) == 0 &&
                         (reinterpret_cast<uintptr_t>(y_bf16) % 8) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0,
                "dicow_fddt_full_combine: bad args");
  const long long total = (long long)rows * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  fddt_full_combine_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const __nv_bfloat16*>(y_bf16), ldy, stno, stno_batch_stride, T, rows, d, pos, x);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_features_to_channels_last(dicow_handle_t h, const float* in, void* out_bf16, int B, int C, int F,
                                               void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, in && out_bf16 && B >= 1 && C >= 1 && F >= 1, "dicow_features_to_channels_last: bad args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid(ceil_div(F, 32), ceil_div(C, 32), B);
  features_to_cl_kernel<<<grid, 256, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), C, F);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_zero_pad_rows(dicow_handle_t h, void* buf_bf16, int B, int T, int C, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, buf_bf16 && B >= 1 && T >= 1 && C >= 1, "dicow_zero_pad_rows: bad args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid(ceil_div(C, 256), B);
  zero_pad_rows_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(buf_bf16), T, C);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_cast_f32_bf16(dicow_handle_t h, const float* in, void* out_bf16, int64_t n, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, in && out_bf16 && n >= 0, "dicow_cast_f32_bf16: bad args");
  if (n == 0) return DICOW_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
