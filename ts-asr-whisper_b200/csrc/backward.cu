// backward.cu -- HBM-bound backward kernels of the fine-tuning step (the GEMM / attention backward live in gemm.cu and
// attention_bwd.cu):
//   * LayerNorm backward fused with the FDDT backward, the residual-stream gradient add and the bf16 copy that the
//     next dgrad / wgrad GEMM consumes; parameter gradients (gamma, beta, the 4 x {w, b} FDDT tables) are column
//     reductions kept in registers by column-owner threads and flushed with one atomic per column per CTA;
//   * column sums (bias gradients);
//   * col2im of the stride-2 subsampling convolutions (dgrad of the implicit-GEMM conv);
//   * CTC backward (alpha / beta in log space -> d logits) and the decoder cross-entropy backward (softmax - target).
// Reference semantics: autograd through nn.LayerNorm (HF:modeling_whisper.py:393,403), FDDT.forward
// (src/models/dicow/FDDT.py:52-62), nn.functional.ctc_loss (src/models/dicow/encoder.py:123-134) and
// SoftLabelCreator.compute_loss (src/models/dicow/modeling_dicow.py:95-144).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm (+ FDDT) backward
// ------------------------------------------------------------------------------------------------------------------
constexpr int LB_ROWS = 2;

struct LnBwdParams {
  const float* x;  // [rows, d] value the forward kernel read (before pending deltas / FDDT)
  const __nv_bfloat16* delta1;
  const __nv_bfloat16* delta2;
  int rows, d, T;
  const float* stno;
  long long stno_bs;
  const float* fddt_w;
  const float* fddt_b;
  const float* gamma;          // NULL: no LayerNorm on this path (dy unused)
  float eps;
  const __nv_bfloat16* dy;     // [rows, d] gradient w.r.t. the LayerNorm output
  const float* g_in;           // [rows, d] gradient arriving through the residual stream (or NULL)
  float* g_out;                // [rows, d] gradient w.r.t. x (== w.r.t. delta1 and delta2)
  __nv_bfloat16* g_out_bf16;   // optional bf16 copy
  float* dgamma;               // [d] +=
  float* dbeta;                // [d] +=
  float* dfddt_w;              // [4, d] += (or NULL)
  float* dfddt_b;              // [4, d] +=
  float* gsum;                 // [d] += column sums of g_out_bf16 (the bias gradient of the Linear that produced a delta), or NULL
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float f4_sum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float4 bf16x4(const uint2 u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// thread t owns float4 column t; a CTA walks over its row pairs; two block reductions per pair
__global__ void __launch_bounds__(512) ln_fddt_bwd_kernel(const LnBwdParams p, const int row_groups) {
  __shared__ float red[2][16][LB_ROWS][2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int nvec = p.d >> 2;
  const bool own = tid < nvec;
  const bool has_fddt = p.stno != nullptr;
  const bool has_ln = p.gamma != nullptr;
  float4 tw[4], tb[4], g4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc_dg = g4, acc_db = g4, acc_w[4], acc_b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tw[c] = make_float4(1.f, 1.f, 1.f, 1.f), tb[c] = g4, acc_w[c] = g4, acc_b[c] = g4;
    if (own && has_fddt) {
      if (p.fddt_w != nullptr) tw[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_w + (long long)c * p.d) + tid);
      tb[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_b + (long long)c * p.d) + tid);
    }
  }
  if (own && has_ln) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma) + tid);
  const float inv_d = 1.0f / (float)p.d;
  for (int grp = blockIdx.x; grp < row_groups; grp += gridDim.x) {
    const int row0 = grp * LB_ROWS;
    float4 xs[LB_ROWS], xp[LB_ROWS], weff[LB_ROWS], dyg[LB_ROWS];
    float m[LB_ROWS][4];
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) {
      const int row = row0 + r;
      const bool ok = own && row < p.rows;
      xs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      dyg[r] = xs[r];
      weff[r] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (ok) {
        xs[r] = reinterpret_cast<const float4*>(p.x + (long long)row * p.d)[tid];
        if (p.delta1 != nullptr) xs[r] = f4_add(xs[r], bf16x4(__ldg(reinterpret_cast<const uint2*>(p.delta1 + (long long)row * p.d) + tid)));
        if (p.delta2 != nullptr) xs[r] = f4_add(xs[r], bf16x4(__ldg(reinterpret_cast<const uint2*>(p.delta2 + (long long)row * p.d) + tid)));
        if (has_ln) dyg[r] = f4_mul(bf16x4(__ldg(reinterpret_cast<const uint2*>(p.dy + (long long)row * p.d) + tid)), g4);
      }
      xp[r] = xs[r];
      if (has_fddt && row < p.rows) {
        const int b = row / p.T, t = row - b * p.T;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f), bb = w;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          m[r][c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.T + t);
          w = f4_add(w, f4_scale(tw[c], m[r][c]));
          bb = f4_add(bb, f4_scale(tb[c], m[r][c]));
        }
        if (p.fddt_w == nullptr) w = make_float4(1.f, 1.f, 1.f, 1.f);  // bias-only FDDT (FDDT.py:43-51): x' = x + sum_c m_c b_c
        weff[r] = w;
        if (ok) xp[r] = f4_add(f4_mul(xs[r], w), bb);
      }
    }
    float4 dxp[LB_ROWS];  // gradient w.r.t. x' (the LayerNorm input)
    if (has_ln) {
      // reduction 1: sum x', sum dy*gamma
      float s1[LB_ROWS], s2[LB_ROWS];
#pragma unroll
      for (int r = 0; r < LB_ROWS; ++r) {
        s1[r] = warp_sum(own ? f4_sum(xp[r]) : 0.f);
        s2[r] = warp_sum(f4_sum(dyg[r]));
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < LB_ROWS; ++r) red[0][warp][r][0] = s1[r], red[0][warp][r][1] = s2[r];
      }
      __syncthreads();
      float mean[LB_ROWS], c1[LB_ROWS];
      float4 xc[LB_ROWS];
#pragma unroll
      for (int r = 0; r < LB_ROWS; ++r) {
        float a = 0.f, b2 = 0.f;
        for (int w = 0; w < nwarps; ++w) a += red[0][w][r][0], b2 += red[0][w][r][1];
        mean[r] = a * inv_d, c1[r] = b2 * inv_d;
        xc[r] = own ? make_float4(xp[r].x - mean[r], xp[r].y - mean[r], xp[r].z - mean[r], xp[r].w - mean[r])
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        s1[r] = warp_sum(f4_sum(f4_mul(xc[r], xc[r])));
        s2[r] = warp_sum(f4_sum(f4_mul(dyg[r], xc[r])));
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < LB_ROWS; ++r) red[1][warp][r][0] = s1[r], red[1][warp][r][1] = s2[r];
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < LB_ROWS; ++r) {
        const int row = row0 + r;
        float a = 0.f, b2 = 0.f;
        for (int w = 0; w < nwarps; ++w) a += red[1][w][r][0], b2 += red[1][w][r][1];
        const float rstd = rsqrtf(a * inv_d + p.eps);
        const float c2 = b2 * inv_d * rstd * rstd;  // mean(dy*gamma * xhat) * rstd, with xhat = xc * rstd
        // dx' = rstd * (dyg - c1 - xhat * mean(dyg * xhat))
        dxp[r] = make_float4(rstd * (dyg[r].x - c1[r] - xc[r].x * c2), rstd * (dyg[r].y - c1[r] - xc[r].y * c2),
                             rstd * (dyg[r].z - c1[r] - xc[r].z * c2), rstd * (dyg[r].w - c1[r] - xc[r].w * c2));
        if (own && row < p.rows) {
          // dgamma += dy * xhat ; dbeta += dy   (dy = dyg / gamma is avoided: reload dy)
          const float4 dy = bf16x4(__ldg(reinterpret_cast<const uint2*>(p.dy + (long long)row * p.d) + tid));
          acc_dg = f4_add(acc_dg, f4_mul(dy, f4_scale(xc[r], rstd)));
          acc_db = f4_add(acc_db, dy);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < LB_ROWS; ++r) dxp[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) {
      const int row = row0 + r;
      if (!(own && row < p.rows)) continue;
      if (p.g_in != nullptr) dxp[r] = f4_add(dxp[r], reinterpret_cast<const float4*>(p.g_in + (long long)row * p.d)[tid]);
      float4 g = dxp[r];
      if (has_fddt) {
        const float4 dweff = f4_mul(dxp[r], xs[r]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc_w[c] = f4_add(acc_w[c], f4_scale(dweff, m[r][c]));
          acc_b[c] = f4_add(acc_b[c], f4_scale(dxp[r], m[r][c]));
        }
        g = f4_mul(dxp[r], weff[r]);
      }
      reinterpret_cast<float4*>(p.g_out + (long long)row * p.d)[tid] = g;
      if (p.g_out_bf16 != nullptr)
        reinterpret_cast<uint2*>(p.g_out_bf16 + (long long)row * p.d)[tid] =
            make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
    }
    __syncthreads();  // red[] reuse
  }
  if (own) {
    if (has_ln && p.dgamma != nullptr) {
      atomic_add4(p.dgamma + 4 * tid, acc_dg);
      atomic_add4(p.dbeta + 4 * tid, acc_db);
    }
    if (has_fddt && p.dfddt_b != nullptr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (p.dfddt_w != nullptr) atomic_add4(p.dfddt_w + (long long)c * p.d + 4 * tid, acc_w[c]);
        atomic_add4(p.dfddt_b + (long long)c * p.d + 4 * tid, acc_b[c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Ring variant (default when d % 8 == 0): the register-resident kernel above is latency bound (ncu / torch.profiler r01:
// 190 us per launch at B = 8 against a 47 us HBM floor -- three CTA-wide barriers per pair of rows and only two rows of
// loads in flight).  Here one producer warp streams whole rows (x fp32 | delta1 | delta2 | dy bf16 | g_in fp32) into a
// shared-memory ring with cp.async.bulk + mbarrier transaction counts, RB_NS "statistics" warps turn a staged row into
// its scalars (mean, rstd, mean(dy gamma), mean(dy gamma xhat) rstd, the 4 STNO weights) with warp shuffles only, and
// the column-owner warps (thread t = float4 column t, FDDT tables / gamma / the 10 column accumulators in registers)
// sweep the staged rows, write g_out / g_out_bf16 and release the stage.  No CTA-wide barrier in the loop; the HBM
// latency is covered by the ring depth.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RB_NS = 4;        // statistics warps
constexpr int RB_MAX_ST = 8;    // ring stages (one row each)

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// RB_VPL: float4 columns per statistics lane == number of column-owner warps (d <= 128 RB_VPL)
template <int RB_VPL>
__global__ void __launch_bounds__((1 + RB_NS + RB_VPL) * 32, 1) ln_fddt_bwd_ring_kernel(const LnBwdParams p, const int nst) {
  constexpr int ncw = RB_VPL;
  extern __shared__ __align__(128) uint8_t rsm[];
  const int d = p.d, nvec = d >> 2;
  const uint32_t xb = d * 4, hb = d * 2;
  const uint32_t off_d1 = xb, off_d2 = xb + hb, off_dy = xb + 2 * hb, off_gin = xb + 3 * hb, stage_bytes = 2 * xb + 3 * hb;
  const bool has_fddt = p.stno != nullptr, has_ln = p.gamma != nullptr;
  uint8_t* ring = rsm;
  float* tab = reinterpret_cast<float*>(rsm + (size_t)nst * stage_bytes);        // [8][d] FDDT w then b (statistics warps)
  float* scal = tab + (has_fddt ? 8 * d : 0);                                     // [nst][8]
  uint64_t* full = reinterpret_cast<uint64_t*>(scal + nst * 8);
  uint64_t* stat = full + nst;
  uint64_t* empty = stat + nst;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&stat[s], 1);
      mbar_init(&empty[s], ncw);
    }
    fence_barrier_init();
  }
  if (has_fddt) {
    for (int i = threadIdx.x; i < 4 * nvec; i += blockDim.x) {
      reinterpret_cast<float4*>(tab)[i] =
          p.fddt_w != nullptr ? __ldg(reinterpret_cast<const float4*>(p.fddt_w) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      reinterpret_cast<float4*>(tab)[4 * nvec + i] = __ldg(reinterpret_cast<const float4*>(p.fddt_b) + i);
    }
  }
  __syncthreads();
  const int n_local = (p.rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // row = i * grid + block
  const float inv_d = 1.0f / (float)d;

  if (warp == 0) {
    // ===================== producer =====================
    const uint32_t tx = xb + (p.delta1 ? hb : 0) + (p.delta2 ? hb : 0) + (has_ln ? hb : 0) + (p.g_in ? xb : 0);
    for (int i = 0; i < n_local; ++i) {
      const int st = i % nst;
      mbar_wait(&empty[st], ((i / nst) & 1) ^ 1);
      if (elect_one()) {
        const long long row = (long long)i * gridDim.x + blockIdx.x;
        uint8_t* dst = ring + (size_t)st * stage_bytes;
        mbar_arrive_expect_tx(&full[st], tx);
        bulk_g2s(dst, p.x + row * d, xb, &full[st]);
        if (p.delta1) bulk_g2s(dst + off_d1, p.delta1 + row * d, hb, &full[st]);
        if (p.delta2) bulk_g2s(dst + off_d2, p.delta2 + row * d, hb, &full[st]);
        if (has_ln) bulk_g2s(dst + off_dy, p.dy + row * d, hb, &full[st]);
        if (p.g_in) bulk_g2s(dst + off_gin, p.g_in + row * d, xb, &full[st]);
      }
      __syncwarp();
    }
    return;
  }
  if (warp <= RB_NS) {
    // ===================== statistics warps: one row each per turn =====================
    for (int i = warp - 1; i < n_local; i += RB_NS) {
      const int st = i % nst;
      const int row = i * (int)gridDim.x + (int)blockIdx.x;
      float m[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_fddt) {
        const int b = row / p.T, t = row - b * p.T;
#pragma unroll
        for (int c = 0; c < 4; ++c) m[c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.T + t);
      }
      mbar_wait(&full[st], (i / nst) & 1);
      float mean = 0.f, rstd = 0.f, c1 = 0.f, c2 = 0.f;
      if (has_ln) {
        const uint8_t* src = ring + (size_t)st * stage_bytes;
        float4 xp[RB_VPL], dg[RB_VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < RB_VPL; ++k) {
          const int c4 = lane + 32 * k;
          xp[k] = make_float4(0.f, 0.f, 0.f, 0.f), dg[k] = xp[k];
          if (c4 < nvec) {
            float4 v = reinterpret_cast<const float4*>(src)[c4];
            if (p.delta1) v = f4_add(v, bf16x4(reinterpret_cast<const uint2*>(src + off_d1)[c4]));
            if (p.delta2) v = f4_add(v, bf16x4(reinterpret_cast<const uint2*>(src + off_d2)[c4]));
            if (has_fddt) {
              float4 w = make_float4(0.f, 0.f, 0.f, 0.f), bb = w;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                w = f4_add(w, f4_scale(reinterpret_cast<const float4*>(tab + c * d)[c4], m[c]));
                bb = f4_add(bb, f4_scale(reinterpret_cast<const float4*>(tab + (4 + c) * d)[c4], m[c]));
              }
              if (p.fddt_w == nullptr) w = make_float4(1.f, 1.f, 1.f, 1.f);  // bias-only FDDT
              v = f4_add(f4_mul(v, w), bb);
            }
            xp[k] = v;
            dg[k] = f4_mul(bf16x4(reinterpret_cast<const uint2*>(src + off_dy)[c4]),
                           __ldg(reinterpret_cast<const float4*>(p.gamma) + c4));
            s1 += f4_sum(v), s2 += f4_sum(dg[k]);
          }
        }
        mean = warp_sum(s1) * inv_d, c1 = warp_sum(s2) * inv_d;
        float q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int k = 0; k < RB_VPL; ++k) {
          if (lane + 32 * k < nvec) {
            const float4 xc = make_float4(xp[k].x - mean, xp[k].y - mean, xp[k].z - mean, xp[k].w - mean);
            q1 += f4_sum(f4_mul(xc, xc)), q2 += f4_sum(f4_mul(dg[k], xc));
          }
        }
        rstd = rsqrtf(warp_sum(q1) * inv_d + p.eps);
        c2 = warp_sum(q2) * inv_d * rstd * rstd;
      }
      if (lane == 0) {
        float4* sc = reinterpret_cast<float4*>(scal + st * 8);
        sc[0] = make_float4(mean, rstd, c1, c2);
        sc[1] = make_float4(m[0], m[1], m[2], m[3]);
        mbar_arrive(&stat[st]);  // release: the scalars are visible to whoever observes this phase
      }
      __syncwarp();
    }
    return;
  }
  // ===================== column owners =====================
  const int tid = (warp - 1 - RB_NS) * 32 + lane;
  const bool own = tid < nvec;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 tw[4], tb[4], g4 = z4, acc_dg = z4, acc_db = z4, acc_gs = z4, acc_w[4], acc_b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tw[c] = make_float4(1.f, 1.f, 1.f, 1.f), tb[c] = z4, acc_w[c] = z4, acc_b[c] = z4;
    if (own && has_fddt) {
      if (p.fddt_w != nullptr) tw[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_w + (long long)c * d) + tid);
      tb[c] = __ldg(reinterpret_cast<const float4*>(p.fddt_b + (long long)c * d) + tid);
    }
  }
  if (own && has_ln) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma) + tid);
  for (int i = 0; i < n_local; ++i) {
    const int st = i % nst;
    const long long row = (long long)i * gridDim.x + blockIdx.x;
    mbar_wait(&stat[st], (i / nst) & 1);
    const uint8_t* src = ring + (size_t)st * stage_bytes;
    float4 g = z4;
    if (own) {
      const float4 s0 = reinterpret_cast<const float4*>(scal + st * 8)[0];  // mean, rstd, c1, c2
      const float4 mk = reinterpret_cast<const float4*>(scal + st * 8)[1];
      const float mm[4] = {mk.x, mk.y, mk.z, mk.w};
      float4 xs = reinterpret_cast<const float4*>(src)[tid];
      if (p.delta1) xs = f4_add(xs, bf16x4(reinterpret_cast<const uint2*>(src + off_d1)[tid]));
      if (p.delta2) xs = f4_add(xs, bf16x4(reinterpret_cast<const uint2*>(src + off_d2)[tid]));
      float4 weff = make_float4(1.f, 1.f, 1.f, 1.f), xp = xs;
      if (has_fddt) {
        float4 w = z4, bb = z4;
#pragma unroll
        for (int c = 0; c < 4; ++c) w = f4_add(w, f4_scale(tw[c], mm[c])), bb = f4_add(bb, f4_scale(tb[c], mm[c]));
        if (p.fddt_w == nullptr) w = make_float4(1.f, 1.f, 1.f, 1.f);  // bias-only FDDT
        weff = w;
        xp = f4_add(f4_mul(xs, w), bb);
      }
      float4 dxp = z4;
      if (has_ln) {
        const float4 dy = bf16x4(reinterpret_cast<const uint2*>(src + off_dy)[tid]);
        const float4 dyg = f4_mul(dy, g4);
        const float4 xc = make_float4(xp.x - s0.x, xp.y - s0.x, xp.z - s0.x, xp.w - s0.x);
        // dx' = rstd * (dyg - mean(dyg) - xc * mean(dyg xhat) rstd)
        dxp = make_float4(s0.y * (dyg.x - s0.z - xc.x * s0.w), s0.y * (dyg.y - s0.z - xc.y * s0.w),
                          s0.y * (dyg.z - s0.z - xc.z * s0.w), s0.y * (dyg.w - s0.z - xc.w * s0.w));
        acc_dg = f4_add(acc_dg, f4_mul(dy, f4_scale(xc, s0.y)));
        acc_db = f4_add(acc_db, dy);
      }
      if (p.g_in) dxp = f4_add(dxp, reinterpret_cast<const float4*>(src + off_gin)[tid]);
      g = dxp;
      if (has_fddt) {
        const float4 dweff = f4_mul(dxp, xs);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc_w[c] = f4_add(acc_w[c], f4_scale(dweff, mm[c]));
          acc_b[c] = f4_add(acc_b[c], f4_scale(dxp, mm[c]));
        }
        g = f4_mul(dxp, weff);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // this warp's smem reads of the stage are done
    if (own) {
      reinterpret_cast<float4*>(p.g_out + row * d)[tid] = g;
      if (p.g_out_bf16 != nullptr) {
        const uint2 pk = make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
        reinterpret_cast<uint2*>(p.g_out_bf16 + row * d)[tid] = pk;
        acc_gs = f4_add(acc_gs, bf16x4(pk));  // what a column sum over the stored bf16 rows would add
      }
    }
  }
  if (own) {
    if (p.gsum != nullptr && p.g_out_bf16 != nullptr) atomic_add4(p.gsum + 4 * tid, acc_gs);
    if (has_ln && p.dgamma != nullptr) {
      atomic_add4(p.dgamma + 4 * tid, acc_dg);
      atomic_add4(p.dbeta + 4 * tid, acc_db);
    }
    if (has_fddt && p.dfddt_b != nullptr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (p.dfddt_w != nullptr) atomic_add4(p.dfddt_w + (long long)c * d + 4 * tid, acc_w[c]);
        atomic_add4(p.dfddt_b + (long long)c * d + 4 * tid, acc_b[c]);
      }
    }
  }
}

template <int VPL>
int launch_ln_bwd_ring(dicow_ctx* ctx, const LnBwdParams& p, int nst, size_t smem, int grid, cudaStream_t stream) {
  auto kfn = ln_fddt_bwd_ring_kernel<VPL>;
  static DeviceHighWater attr_smem;
  if (attr_smem.raise(ctx, smem)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  kfn<<<grid, (1 + RB_NS + VPL) * 32, smem, stream>>>(p, nst);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// backward of fddt_full_combine (elementwise.cu): dY[r, c d : (c + 1) d] = stno[b, c, t] * G[r, :]  (bf16), the gradient
// of the four stacked class transforms of full-matrix FDDT (src/models/dicow/FDDT.py:52-62); dx and the CustomLinear
// weight / bias gradients follow from ordinary dgrad / wgrad GEMMs and column sums on dY.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fddt_full_scatter_kernel(const float* __restrict__ g, const float* __restrict__ stno,
                                                                long long stno_bs, int T, int rows, int d,
                                                                __nv_bfloat16* __restrict__ dy, long long ldy) {
  const int nvec = d >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / nvec), c4 = (int)(i - (long long)r * nvec);
    const int b = r / T, t = r - b * T;
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + (long long)r * d) + c4);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float m = __ldg(stno + (long long)b * stno_bs + (long long)c * T + t);
      reinterpret_cast<uint2*>(dy + (long long)r * ldy + (long long)c * d)[c4] =
          make_uint2(pack_bf16(m * v.x, m * v.y), pack_bf16(m * v.z, m * v.w));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out[n] += sum_rows X[row, n]     (bias gradients);  X bf16 or fp32
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, long long ld, int rows, int N,
                                                     float* __restrict__ out, float alpha, int rows_per_block) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (n >= N) return;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += (float)X[(long long)r * ld + n];
  atomicAdd(out + n, alpha * acc);
}

// bf16 rows read as 16-byte vectors: a block covers a slab of 256 columns x rows_per_block rows; thread (cg, rl) owns 8 columns
// and every 8th row, four loads in flight, the 8 row lanes of a column group meet in shared memory.  (The scalar kernel above
// moved 2 bytes per load instruction through a serial add chain: 2.6 TB/s on the fine-tune step's 165 launches.)
__global__ void __launch_bounds__(256) colsum_bf16_vec_kernel(const __nv_bfloat16* __restrict__ X, long long ld, int rows, int N,
                                                              float* __restrict__ out, float alpha, int rows_per_block) {
  __shared__ float part[8][32][8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const __nv_bfloat16* base = X + n;
    int r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (long long)(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ld));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[rl][cg][j] = acc[j];
  __syncthreads();
  // thread t sums column t of the slab over the 8 row lanes
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][c >> 3][c & 7];
    atomicAdd(out + blockIdx.x * 256 + c, alpha * t);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// col2im of Conv1d(k = 3, padding 1, stride s):  dX[b, t, c] = sum_k dcol[b, (t + 1 - k) / s, k C + c]
//   (t + 1 - k) % s == 0 and 0 <= (t + 1 - k) / s < T_out
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx,
                                                     int B, int T, int T_out, int C, int stride, long long dx_bs,
                                                     long long dx_rs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % T);
  const int b = (int)(i / ((long long)C * T));
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int u = t + 1 - k;
    if (u >= 0 && (u % stride) == 0 && u / stride < T_out)
      acc += __bfloat162float(dcol[((long long)b * T_out + u / stride) * (3LL * C) + (long long)k * C + c]);
  }
  dx[(long long)b * dx_bs + (long long)t * dx_rs + c] = __float2bfloat16_rn(acc);
}

// ------------------------------------------------------------------------------------------------------------------
// out_bf16[i] = g[i] * gelu'(pre[i])     (conv stem backward: the GELU sits between an fp32 gradient and the conv dgrad)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dgelu_mul_kernel(const T* __restrict__ g, long long ldg,
                                                        const __nv_bfloat16* __restrict__ pre, long long ldp,
                                                        __nv_bfloat16* __restrict__ out, long long ldo, int rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int c = (int)(i % cols);
  const long long r = i / cols;
  const float gv = (float)g[r * ldg + c];
  out[r * ldo + c] = __float2bfloat16_rn(gv * dgelu_erf_fast(__bfloat162float(pre[r * ldp + c])));
}

// ------------------------------------------------------------------------------------------------------------------
// SCB gate backward (src/models/dicow/layers.py:79-93,168: q + tanh(gate) * upd):
//   dupd_bf16 = tanh(gate) * G          dgate += (1 - tanh^2(gate)) * sum(G .* upd)
// HBM-bound streaming pass (6 B read + 2 B written per element); one fp32 atomic per CTA for the scalar gradient.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ G, long long ldg,
                                                       const __nv_bfloat16* __restrict__ upd, long long ldu,
                                                       const float* __restrict__ gate, __nv_bfloat16* __restrict__ dupd,
                                                       long long ldd, int rows, int cols, float* __restrict__ dgate) {
  const float th = tanhf(*gate);
  const int cp = cols >> 1;  // column pairs (cols is even)
  const long long total = (long long)rows * cp;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cp;
    const int c = (int)(i - r * cp) * 2;
    const float2 gv = *reinterpret_cast<const float2*>(G + r * ldg + c);
    const float2 uv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(upd + r * ldu + c));
    acc = fmaf(gv.x, uv.x, fmaf(gv.y, uv.y, acc));
    *reinterpret_cast<__nv_bfloat162*>(dupd + r * ldd + c) = __floats2bfloat162_rn(th * gv.x, th * gv.y);
  }
  if (dgate == nullptr) return;
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    acc = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(dgate, (1.f - th * th) * acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// decoder embedding backward: d_tok[ids[r], :] += g[r, :], d_pos[past + r % S, :] += g[r, :]   (fp32 atomics)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embedding_bwd_kernel(const float* __restrict__ g, const long long* __restrict__ ids,
                                                            int rows, int S, int d, int past, int vocab,
                                                            float* __restrict__ d_tok, float* __restrict__ d_pos) {
  const int r = blockIdx.x;
  const long long id = ids[r];
  const int s = past + r % S;
  const bool tok_ok = d_tok != nullptr && id >= 0 && id < vocab;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = g[(long long)r * d + c];
    if (tok_ok) atomicAdd(d_tok + id * d + c, v);
    if (d_pos != nullptr) atomicAdd(d_pos + (long long)s * d + c, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out_bf16[r, c] = c < cols ? in_f32[r, c] : 0, c < cols_out  (fp32 gradient -> padded bf16 GEMM operand)
// ------------------------------------------------------------------------------------------------------------------
// grid (column blocks, rows): a thread converts two adjacent columns (one 4-byte store; the fp32 rows may be unaligned -- V + 1 =
// 51 867 columns -- so the loads stay scalar, coalesced across the warp).  The one-element-per-thread form with a 64-bit
// division per element ran at 1.5 TB/s on the [6000, 51867] CTC logits gradient (1.2 ms of the CTC pre-train step).
__global__ void __launch_bounds__(256) cast_2d_kernel(const float* __restrict__ in, long long ldi,
                                                      __nv_bfloat16* __restrict__ out, long long ldo, int rows, int cols,
                                                      int cols_out) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= cols_out) return;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* src = in + r * ldi;
    const float a = c < cols ? __ldg(src + c) : 0.f;
    const float b = c + 1 < cols ? __ldg(src + c + 1) : 0.f;
    __nv_bfloat16* dst = out + r * ldo + c;
    if (c + 1 < cols_out && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(dst) = pack_bf16(a, b);
    } else {
      dst[0] = __float2bfloat16_rn(a);
      if (c + 1 < cols_out) dst[1] = __float2bfloat16_rn(b);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTC backward
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float mx = fmaxf(a, b);
  return mx + log1pf(__expf(-fabsf(a - b)));
}

struct CtcBwdParams {
  const float* logits;  // [B, T, V1], rows ld elements apart
  long long ld;
  const float* lse;     // [B * T] natural-log sum exp of every row
  int T, V1, Lmax;
  const long long* labels;
  float* alpha;  // [B, T, S] workspace, S = 2 Lmax + 1
  float* beta;   // [B, T, S]
  float* nll;    // [B]
  int mean, B;
  __nv_bfloat16* dlogits;  // [B, T, ldd]  (fp32 when out_f32)
  long long ldd;
  float loss_scale;         // upstream gradient * (ctc weight)
  const float* scale_dev;   // optional device-resident multiplier of loss_scale (the upstream gradient)
  int out_f32;
};

__device__ __forceinline__ void store_grad(void* base, bool f32, long long i, float v) {
  if (f32)
    reinterpret_cast<float*>(base)[i] = v;
  else
    reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
}

// alpha and beta lattices: one CTA per (utterance, direction); thread == extended-label state
__global__ void __launch_bounds__(1024) ctc_lattice_kernel(const CtcBwdParams p) {
  extern __shared__ float buf[];  // [2][S]
  const int b = blockIdx.x, dir = blockIdx.y, s = threadIdx.x;
  const long long* lab = p.labels + (long long)b * p.Lmax;
  __shared__ int s_len;
  if (s == 0) {
    int n = 0;
    for (int i = 0; i < p.Lmax; ++i) n += lab[i] >= 0 ? 1 : 0;
    s_len = n;
  }
  __syncthreads();
  const int L = s_len, S = 2 * L + 1, Sm = 2 * p.Lmax + 1, blank = p.V1 - 1;
  const bool active = s < S;
  int cls = blank;
  bool skip_prev = false, skip_next = false;  // may jump from s-2 (alpha) / to s+2 (beta)
  if (active && (s & 1)) {
    cls = (int)lab[s >> 1];
    skip_prev = s >= 3 && lab[(s >> 1) - 1] != cls;
    skip_next = s + 2 < S && lab[(s >> 1) + 1] != cls;
  }
  const float* lg = p.logits + (long long)b * p.T * p.ld;
  const float* ls = p.lse + (long long)b * p.T;
  float* lat = (dir == 0 ? p.alpha : p.beta) + (long long)b * p.T * Sm;
  float* cur = buf;
  float* nxt = buf + Sm;
  if (dir == 0) {
    if (active) {
      const float v = (s <= 1) ? lg[cls] - ls[0] : -INFINITY;
      cur[s] = v, lat[s] = v;
    }
    __syncthreads();
    for (int t = 1; t < p.T; ++t) {
      if (active) {
        float a = cur[s];
        if (s >= 1) a = log_add(a, cur[s - 1]);
        if (skip_prev) a = log_add(a, cur[s - 2]);
        a += lg[(long long)t * p.ld + cls] - ls[t];
        nxt[s] = a, lat[(long long)t * Sm + s] = a;
      }
      __syncthreads();
      float* tmp = cur;
      cur = nxt, nxt = tmp;
    }
    if (s == 0) {
      float ll = cur[S - 1];
      if (S >= 2) ll = log_add(ll, cur[S - 2]);
      p.nll[b] = -ll;
    }
  } else {
    const int tl = p.T - 1;
    if (active) {
      const float v = (s >= S - 2) ? lg[(long long)tl * p.ld + cls] - ls[tl] : -INFINITY;
      cur[s] = v, lat[(long long)tl * Sm + s] = v;
    }
    __syncthreads();
    for (int t = p.T - 2; t >= 0; --t) {
      if (active) {
        float a = cur[s];
        if (s + 1 < S) a = log_add(a, cur[s + 1]);
        if (skip_next) a = log_add(a, cur[s + 2]);
        a += lg[(long long)t * p.ld + cls] - ls[t];
        nxt[s] = a, lat[(long long)t * Sm + s] = a;
      }
      __syncthreads();
      float* tmp = cur;
      cur = nxt, nxt = tmp;
    }
  }
}

// d logits[b, t, v] = scale_b * (softmax(logits)[v] - occupancy[b, t, v]); one CTA per (b, t) row.
//   occupancy[v] = sum_{s: ext[s] = v} exp(alpha_t(s) + beta_t(s) - lp[t, v] + nll)
// scale_b = loss_scale / (B * max(L_b, 1)) for "mean", loss_scale for "sum"; 0 when the loss is infinite (zero_infinity).
__global__ void __launch_bounds__(512) ctc_grad_kernel(const CtcBwdParams p) {
  extern __shared__ float occ[];  // [S]: occupancy of the class of state s, accumulated at its first occurrence
  const int bt = blockIdx.x, b = bt / p.T, t = bt - b * p.T;
  const int tid = threadIdx.x;
  const long long* lab = p.labels + (long long)b * p.Lmax;
  __shared__ int s_len;
  __shared__ float s_blank;
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < p.Lmax; ++i) n += lab[i] >= 0 ? 1 : 0;
    s_len = n;
    s_blank = 0.f;
  }
  __syncthreads();
  const int L = s_len, S = 2 * L + 1, Sm = 2 * p.Lmax + 1, blank = p.V1 - 1;
  const float nll = p.nll[b];
  const bool finite = nll < INFINITY;  // also false for NaN
  const float ls = p.loss_scale * (p.scale_dev != nullptr ? __ldg(p.scale_dev) : 1.f);
  const float scale = !finite ? 0.f : (p.mean ? ls / ((float)p.B * (float)max(L, 1)) : ls);
  const float* lg = p.logits + (long long)bt * p.ld;
  const float lse = p.lse[bt];
  const bool f32 = p.out_f32 != 0;
  void* out = f32 ? static_cast<void*>(reinterpret_cast<float*>(p.dlogits) + (long long)bt * p.ldd)
                  : static_cast<void*>(p.dlogits + (long long)bt * p.ldd);
  // dense part: scale * softmax
  for (int v = tid; v < p.ldd; v += blockDim.x) store_grad(out, f32, v, v < p.V1 ? scale * __expf(lg[v] - lse) : 0.f);
  if (!finite) return;  // uniform
  for (int s = tid; s < S; s += blockDim.x) occ[s] = 0.f;
  __syncthreads();
  const float* al = p.alpha + (long long)bt * Sm;
  const float* be = p.beta + (long long)bt * Sm;
  float blank_acc = 0.f;
  for (int s = tid; s < S; s += blockDim.x) {
    const int cls = (s & 1) ? (int)lab[s >> 1] : blank;
    const float lp = lg[cls] - lse;
    const float w = __expf(al[s] + be[s] - lp + nll);  // alpha and beta both contain the emission at t
    if (!(s & 1)) {
      blank_acc += w;
    } else {
      int first = s;  // first state with the same class (repeated labels share one logit)
      for (int u = 1; u < s; u += 2)
        if (lab[u >> 1] == cls) {
          first = u;
          break;
        }
      atomicAdd(&occ[first], w);
    }
  }
  blank_acc = warp_sum(blank_acc);
  if ((tid & 31) == 0 && blank_acc != 0.f) atomicAdd(&s_blank, blank_acc);
  __syncthreads();
  // sparse part: subtract the occupancy at the label classes (each class exactly once) and at the blank
  for (int s = 1 + 2 * tid; s < S; s += 2 * blockDim.x) {
    const float o = occ[s];
    if (o != 0.f) {
      const int cls = (int)lab[s >> 1];
      store_grad(out, f32, cls, scale * (__expf(lg[cls] - lse) - o));
    }
  }
  if (tid == 0) store_grad(out, f32, blank, scale * (__expf(lg[blank] - lse) - s_blank));
}

// ------------------------------------------------------------------------------------------------------------------
// decoder cross-entropy backward: d logits = scale * (softmax - target_of_the_winning_stream), rows with label -100 -> 0
// ------------------------------------------------------------------------------------------------------------------
struct CeBwdParams {
  const float* logits;
  long long ld;
  int R, V;
  const long long* labels;
  const long long* upp;
  int ts_begin, n_ts;
  const float* smooth;
  int soft_mode;
  const float* row_lower;  // optional per-row losses of both streams from the forward (to pick the min); NULL: recompute
  float scale;             // upstream gradient / normaliser
  const float* scale_dev;  // optional device-resident multiplier of scale
  __nv_bfloat16* dlogits;
  long long ldd;
};

__device__ __forceinline__ float ce_target_dot(const CeBwdParams& p, const float* row, long long lab, float* sm) {
  if (lab < 0) lab = 0;
  if (lab >= p.V) lab = p.V - 1;
  if (p.n_ts > 0 && lab >= p.ts_begin && lab < p.ts_begin + p.n_ts) {
    const float* w = p.smooth + (lab - p.ts_begin) * p.n_ts;
    float acc = 0.f;
    for (int k = threadIdx.x; k < p.n_ts; k += blockDim.x) acc = fmaf(__ldg(w + k), row[p.ts_begin + k], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    float t = 0.f;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += sm[w2];
    __syncthreads();
    return t;
  }
  return row[lab];
}

__global__ void __launch_bounds__(512) ce_bwd_kernel(const CeBwdParams p) {
  __shared__ float sa[16], sb[16];
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const float* row = p.logits + (long long)r * p.ld;
  __nv_bfloat16* out = p.dlogits + (long long)r * p.ldd;
  const long long lab = p.labels[r];
  const long long ulab = p.upp != nullptr ? p.upp[r] : lab;
  // which stream wins the per-token min (ties -> the lower-case stream, like torch.min on equal values)
  bool use_upper = false;
  if (p.upp != nullptr && ulab != lab) {
    // loss = lse - <target, logits>: smaller loss == larger target dot
    const float dl = ce_target_dot(p, row, lab, sa), du = ce_target_dot(p, row, ulab, sa);
    const bool lo_ign = !p.soft_mode && lab == -100, up_ign = !p.soft_mode && ulab == -100;
    if (lo_ign || up_ign) {
      // hard fallback: an ignored stream has loss 0, the other a positive loss -> the ignored one is the min (no gradient)
      for (int v = tid; v < p.ldd; v += blockDim.x) out[v] = __float2bfloat16_rn(0.f);
      return;
    }
    use_upper = du > dl;
  }
  const long long tl = use_upper ? ulab : lab;
  const bool masked = p.soft_mode ? (lab == -100) : (tl == -100);
  if (masked) {
    for (int v = tid; v < p.ldd; v += blockDim.x) out[v] = __float2bfloat16_rn(0.f);
    return;
  }
  // logsumexp of the row
  float m = -INFINITY, s = 0.f;
  for (int v = tid; v < p.V; v += blockDim.x) {
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
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) sa[warp] = m, sb[warp] = s;
  __syncthreads();
  float mt = -INFINITY;
  for (int w = 0; w < nw; ++w) mt = fmaxf(mt, sa[w]);
  float st = 0.f;
  for (int w = 0; w < nw; ++w) st += sa[w] == -INFINITY ? 0.f : sb[w] * __expf(sa[w] - mt);
  const float lse = mt + logf(st);
  long long t = tl < 0 ? 0 : (tl >= p.V ? p.V - 1 : tl);
  const bool is_ts = p.n_ts > 0 && t >= p.ts_begin && t < p.ts_begin + p.n_ts;
  const float* w = is_ts ? p.smooth + (t - p.ts_begin) * p.n_ts : nullptr;
  const float scale = p.scale * (p.scale_dev != nullptr ? __ldg(p.scale_dev) : 1.f);
  for (int v = tid; v < p.ldd; v += blockDim.x) {
    float g = 0.f;
    if (v < p.V) {
      float tgt = 0.f;
      if (is_ts) {
        if (v >= p.ts_begin && v < p.ts_begin + p.n_ts) tgt = __ldg(w + (v - p.ts_begin));
      } else if (v == t) {
        tgt = 1.f;
      }
      g = scale * (__expf(row[v] - lse) - tgt);
    }
    out[v] = __float2bfloat16_rn(g);
  }
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_fddt_full_scatter(dicow_handle_t h, const float* g, const float* stno, int64_t stno_batch_stride, int T,
                                       int rows, int d, void* dy_bf16, int64_t ldy, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, g && stno && dy_bf16 && rows >= 1 && T >= 1 && (rows % T) == 0 && d >= 4 && (d % 4) == 0 && (ldy % 4) == 0 &&
                         (reinterpret_cast<uintptr_t>(dy_bf16) % 8) == 0 && (reinterpret_cast<uintptr_t>(g) % 16) == 0,
                "dicow_fddt_full_scatter: bad args");
  const long long total = (long long)rows * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  fddt_full_scatter_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      g, stno, stno_batch_stride, T, rows, d, reinterpret_cast<__nv_bfloat16*>(dy_bf16), ldy);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_colsum(dicow_handle_t h, const void* x, int is_bf16, int64_t ld, int rows, int N, float* out,
                            float alpha, void* stream_);

extern "C" int dicow_layernorm_fddt_bwd(dicow_handle_t h, const dicow_ln_bwd_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_ln_bwd_args_t), "dicow_layernorm_fddt_bwd: bad args struct");
  DICOW_REQUIRE(ctx, a->x && a->g_out && a->rows >= 1 && a->d >= 4 && (a->d % 4) == 0 && a->d <= 2048,
                "dicow_layernorm_fddt_bwd: need x, g_out, d %% 4 == 0, d <= 2048");
  DICOW_REQUIRE(ctx, a->gamma == nullptr || a->dy_bf16 != nullptr, "dicow_layernorm_fddt_bwd: gamma without dy");
  DICOW_REQUIRE(ctx, a->stno == nullptr || (a->fddt_b != nullptr && a->T >= 1 && (a->rows % a->T) == 0),
                "dicow_layernorm_fddt_bwd: FDDT needs fddt_b and rows %% T == 0");
  LnBwdParams p{};
  p.x = a->x, p.delta1 = reinterpret_cast<const __nv_bfloat16*>(a->delta1_bf16);
  p.delta2 = reinterpret_cast<const __nv_bfloat16*>(a->delta2_bf16);
  p.rows = a->rows, p.d = a->d, p.T = a->T > 0 ? a->T : a->rows;
  p.stno = a->stno, p.stno_bs = a->stno_batch_stride, p.fddt_w = a->fddt_w, p.fddt_b = a->fddt_b;
  p.gamma = a->gamma, p.eps = a->eps, p.dy = reinterpret_cast<const __nv_bfloat16*>(a->dy_bf16), p.g_in = a->g_in;
  p.g_out = a->g_out, p.g_out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->g_out_bf16);
  p.dgamma = a->dgamma, p.dbeta = a->dbeta, p.dfddt_w = a->dfddt_w, p.dfddt_b = a->dfddt_b;
  p.gsum = a->g_colsum;
  DICOW_REQUIRE(ctx, a->g_colsum == nullptr || a->g_out_bf16 != nullptr, "dicow_layernorm_fddt_bwd: g_colsum needs g_out_bf16");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool aligned = (a->d % 8) == 0 && (reinterpret_cast<uintptr_t>(a->x) % 16) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->delta1_bf16) % 16) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->delta2_bf16) % 16) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->dy_bf16) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->g_in) % 16) == 0;
  if (aligned && a->rows >= 64) {
    // ring kernel: stage = one row of every input stream; as many stages as fit next to the FDDT tables
    const size_t stage = (size_t)a->d * 14, fixed = (p.stno ? (size_t)32 * a->d : 0) + RB_MAX_ST * (32 + 24) + 256;
    int nst = (int)(((size_t)ctx->max_smem_optin - fixed) / stage);
    nst = nst > RB_MAX_ST ? RB_MAX_ST : nst;
    if (nst >= 3) {
      const int ncw = ceil_div(a->d / 4, 32);
      const size_t smem = (size_t)nst * stage + fixed;
      const int grid = a->rows < ctx->num_sms ? a->rows : ctx->num_sms;
      if (ncw <= 4) return launch_ln_bwd_ring<4>(ctx, p, nst, smem, grid, stream);
      if (ncw <= 8) return launch_ln_bwd_ring<8>(ctx, p, nst, smem, grid, stream);
      if (ncw <= 10) return launch_ln_bwd_ring<10>(ctx, p, nst, smem, grid, stream);
      return launch_ln_bwd_ring<16>(ctx, p, nst, smem, grid, stream);
    }
  }
  const int threads = ceil_div(a->d / 4, 32) * 32;
  const int groups = ceil_div(a->rows, LB_ROWS);
  const int grid = groups < 2 * ctx->num_sms ? groups : 2 * ctx->num_sms;
  ln_fddt_bwd_kernel<<<grid, threads, 0, stream>>>(p, groups);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  if (p.gsum != nullptr)  // this kernel does not carry the column sums: a separate pass over the bf16 rows
    return dicow_colsum(h, p.g_out_bf16, 1, a->d, a->rows, a->d, p.gsum, 1.0f, stream_);
  return DICOW_OK;
}

extern "C" int dicow_colsum(dicow_handle_t h, const void* x, int is_bf16, int64_t ld, int rows, int N, float* out,
                            float alpha, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, x && out && rows >= 1 && N >= 1, "dicow_colsum: bad args");
  const int rpb = 128;
  dim3 grid(ceil_div(N, 256), ceil_div(rows, rpb));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (is_bf16 && (N % 8) == 0 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) {
    // about four blocks per SM
    long long want = ((long long)rows * ceil_div(N, 256)) / (4LL * ctx->num_sms);
    want = ((want + 31) / 32) * 32;
    const int rpb_v = (int)(want < 64 ? 64 : (want > 1024 ? 1024 : want));
    dim3 gv(ceil_div(N, 256), ceil_div(rows, rpb_v));
    colsum_bf16_vec_kernel<<<gv, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, N, out, alpha, rpb_v);
  } else if (is_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, N, out, alpha, rpb);
  else
    colsum_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), ld, rows, N, out, alpha, rpb);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_conv1d_col2im(dicow_handle_t h, const void* dcol_bf16, void* dx_bf16, int B, int T, int T_out, int C,
                                   int stride, int64_t dx_batch_stride, int64_t dx_row_stride, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, dcol_bf16 && dx_bf16 && B >= 1 && T >= 1 && T_out >= 1 && C >= 1 && stride >= 1, "dicow_conv1d_col2im: bad args");
  const long long total = (long long)B * T * C;
  col2im_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dcol_bf16), reinterpret_cast<__nv_bfloat16*>(dx_bf16), B, T, T_out, C, stride,
      dx_batch_stride, dx_row_stride);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_ctc_loss_bwd(dicow_handle_t h, const dicow_ctc_bwd_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_ctc_bwd_args_t), "dicow_ctc_loss_bwd: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->lse && a->labels && a->workspace && a->dlogits_bf16 && a->B >= 1 && a->T >= 1 && a->V1 >= 2,
                "dicow_ctc_loss_bwd: bad args");
  DICOW_REQUIRE(ctx, a->Lmax >= 0 && 2 * a->Lmax + 1 <= 1024 && a->ldd >= a->V1, "dicow_ctc_loss_bwd: bad Lmax / ldd");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int S = 2 * a->Lmax + 1;
  CtcBwdParams p{};
  p.logits = a->logits, p.lse = a->lse, p.T = a->T, p.V1 = a->V1, p.Lmax = a->Lmax, p.B = a->B;
  p.ld = a->ld > 0 ? a->ld : a->V1;
  DICOW_REQUIRE(ctx, p.ld >= a->V1, "dicow_ctc_loss_bwd: ld < V1");
  p.labels = reinterpret_cast<const long long*>(a->labels);
  p.alpha = a->workspace;
  p.beta = p.alpha + (long long)a->B * a->T * S;
  p.nll = p.beta + (long long)a->B * a->T * S;
  p.mean = a->reduction_mean, p.dlogits = reinterpret_cast<__nv_bfloat16*>(a->dlogits_bf16), p.ldd = a->ldd;
  p.loss_scale = a->loss_scale, p.scale_dev = a->scale_dev, p.out_f32 = a->out_f32;
  const int threads = ((S + 31) / 32) * 32;
  ctc_lattice_kernel<<<dim3(a->B, 2), threads, 2 * S * sizeof(float), stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  ctc_grad_kernel<<<a->B * a->T, 512, S * sizeof(float), stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_softlabel_ce_bwd(dicow_handle_t h, const dicow_softlabel_ce_bwd_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_softlabel_ce_bwd_args_t), "dicow_softlabel_ce_bwd: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->labels && a->dlogits_bf16 && a->rows >= 1 && a->V >= 1 && a->ldd >= a->V,
                "dicow_softlabel_ce_bwd: bad args");
  CeBwdParams p{};
  p.logits = a->logits, p.ld = a->ld, p.R = a->rows, p.V = a->V;
  p.labels = reinterpret_cast<const long long*>(a->labels), p.upp = reinterpret_cast<const long long*>(a->upp_labels);
  p.ts_begin = a->ts_begin, p.n_ts = a->n_ts, p.smooth = a->smoothing, p.soft_mode = a->soft_mode;
  p.scale = a->scale, p.scale_dev = a->scale_dev;
  p.dlogits = reinterpret_cast<__nv_bfloat16*>(a->dlogits_bf16), p.ldd = a->ldd;
  ce_bwd_kernel<<<a->rows, 512, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_dgelu_mul(dicow_handle_t h, const void* g, int g_is_bf16, int64_t ldg, const void* pre_bf16, int64_t ldp,
                               void* out_bf16, int64_t ldo, int rows, int cols, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, g && pre_bf16 && out_bf16 && rows >= 1 && cols >= 1, "dicow_dgelu_mul: bad args");
  const long long total = (long long)rows * cols;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (g_is_bf16)
    dgelu_mul_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(g), ldg,
                                                              reinterpret_cast<const __nv_bfloat16*>(pre_bf16), ldp,
                                                              reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, rows, cols);
  else
    dgelu_mul_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(g), ldg,
                                                      reinterpret_cast<const __nv_bfloat16*>(pre_bf16), ldp,
                                                      reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, rows, cols);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_gate_bwd(dicow_handle_t h, const float* g, int64_t ldg, const void* upd_bf16, int64_t ldu, const float* gate,
                              void* dupd_bf16, int64_t ldd, int rows, int cols, float* dgate, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, g && upd_bf16 && gate && dupd_bf16 && rows >= 1 && cols >= 2 && (cols % 2) == 0 && (ldg % 2) == 0 &&
                         (ldu % 2) == 0 && (ldd % 2) == 0,
                "dicow_gate_bwd: bad args (cols and leading dimensions must be even)");
  const long long total = (long long)rows * (cols / 2);
  const long long want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148LL * 8 ? want : 148LL * 8);  // 8 resident CTAs of 256 threads per SM
  gate_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      g, ldg, reinterpret_cast<const __nv_bfloat16*>(upd_bf16), ldu, gate, reinterpret_cast<__nv_bfloat16*>(dupd_bf16), ldd,
      rows, cols, dgate);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_embedding_bwd(dicow_handle_t h, const float* g, const int64_t* ids, int rows, int S, int d, int past,
                                   int vocab, float* d_embed_tokens, float* d_embed_positions, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, g && ids && rows >= 1 && S >= 1 && d >= 1 && past >= 0, "dicow_embedding_bwd: bad args");
  embedding_bwd_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      g, reinterpret_cast<const long long*>(ids), rows, S, d, past, vocab, d_embed_tokens, d_embed_positions);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_cast_f32_bf16_2d(dicow_handle_t h, const float* in, int64_t ldi, void* out_bf16, int64_t ldo, int rows,
                                      int cols, int cols_out, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, in && out_bf16 && rows >= 1 && cols >= 1 && cols_out >= cols && ldo >= cols_out && ldi >= cols,
                "dicow_cast_f32_bf16_2d: bad args");
  const dim3 grid((unsigned)ceil_div(ceil_div(cols_out, 2), 256), (unsigned)(rows < 65535 ? rows : 65535));
  cast_2d_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      in, ldi, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, rows, cols, cols_out);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
