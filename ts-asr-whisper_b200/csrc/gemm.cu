// gemm.cu -- persistent warp-specialised bf16 GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[b, m, n] = epilogue( sum_k A[b, m, k] * W[n, k] )      A, W bf16 K-major; fp32 accumulate in TMEM
//
// Roles (384 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tile schedule, N fastest so the
// CTAs running concurrently share A rows through L2):
//   warp 0 lane 0 : TMA producer  -- A tile [128 x 64] + W tile [BN x 64] per stage, 128B swizzle, mbarrier tx
//   warp 1 lane 0 : MMA issuer    -- 4 x tcgen05.mma (M128 x N(BN) x K16) per stage, tcgen05.commit frees the stage
//   warp 2        : TMEM allocator (2 accumulator buffers of BN fp32 columns -> epilogue overlaps next mainloop)
//   warps 4..11   : epilogue      -- tcgen05.ld (thread == accumulator row), fused bias / GELU / residual / FDDT
//
// Replaces nn.Linear / nn.Conv1d calls of the reference (see include/dicow_b200.h for the call-site list).
#include <math.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int kNonEpiWarps = 4;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = (kNonEpiWarps + kEpiWarps) * 32;

struct GemmParams {
  int nb, Mb, N, K, K1;
  int tiles_m, tiles_n, total_tiles, kblocks;
  int splits, kb_per_split;  // split-K (DICOW_EPI_ACCUM_F32 only): work item = (tile, split)
  const float* bias;
  void* out;
  long long ldo, out_bs;
  int out_vec_ok;  // rows of `out` (and resid) are 16-byte aligned at every 32-column chunk
  const float* resid;
  long long ldr, resid_bs;
  const float* gate;
  const float* stno;
  long long stno_bs;
  const float* fddt_w;
  const float* fddt_b;
  const float* pos;
  __nv_bfloat16* aux;  // [.., N] bf16, leading dimension ldo: pre-activation saved by GELU_SAVE / read by DGELU
  int serial_drain;    // 1 (default): one chunk at a time, accumulator handed back after the last store; 0: pipelined
  int tma_out;         // CTA-pair kernel, bf16 outputs: drain through swizzled shared memory + cp.async.bulk.tensor stores
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // 512 or 256: both powers of two
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---- epilogue on one 32-column chunk held in registers -------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&v)[32], int b, int m, int n, int ncols,
                                               const float (&mask)[4], const float4* bias_pre = nullptr) {
  // bias (bias_pre: the chunk's 32 values, loaded by the caller while the accumulator was still in flight)
  if (p.bias != nullptr) {
    if (ncols == 32 && bias_pre != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[4 * j] += bias_pre[j].x, v[4 * j + 1] += bias_pre[j].y, v[4 * j + 2] += bias_pre[j].z, v[4 * j + 3] += bias_pre[j].w;
    } else if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += __ldg(p.bias + n + j);
    }
  }
  if constexpr (EPI == DICOW_EPI_GELU_SAVE_BF16) {  // training forward: keep the pre-activation for the backward pass
    __nv_bfloat16* a = p.aux + (long long)b * p.out_bs + (long long)m * p.ldo + n;
    if (ncols == 32 && p.out_vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        q.x = pack_bf16(v[j], v[j + 1]), q.y = pack_bf16(v[j + 2], v[j + 3]);
        q.z = pack_bf16(v[j + 4], v[j + 5]), q.w = pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(a + j) = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) a[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if constexpr (EPI == DICOW_EPI_DGELU_BF16) {  // backward: out = acc * gelu'(pre)
    const __nv_bfloat16* a = p.aux + (long long)b * p.out_bs + (long long)m * p.ldo + n;
    if (ncols == 32 && p.out_vec_ok) {  // 4 x 16-byte loads per thread-row instead of 32 two-byte ones
      uint4 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = __ldg(reinterpret_cast<const uint4*>(a) + j);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
          v[8 * j + 2 * k] *= dgelu_erf_fast(f.x);
          v[8 * j + 2 * k + 1] *= dgelu_erf_fast(f.y);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] *= dgelu_erf_fast(__bfloat162float(a[j]));
    }
  }
  if constexpr (EPI == DICOW_EPI_BIAS_GELU_BF16 || EPI == DICOW_EPI_GELU_FDDT_POS_F32 || EPI == DICOW_EPI_GELU_SAVE_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
  }

  if constexpr (EPI == DICOW_EPI_BIAS_BF16 || EPI == DICOW_EPI_BIAS_GELU_BF16 || EPI == DICOW_EPI_GELU_SAVE_BF16 ||
                EPI == DICOW_EPI_DGELU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)b * p.out_bs + (long long)m * p.ldo + n;
    if (ncols == 32 && p.out_vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        q.x = pack_bf16(v[j], v[j + 1]);
        q.y = pack_bf16(v[j + 2], v[j + 3]);
        q.z = pack_bf16(v[j + 4], v[j + 5]);
        q.w = pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* o = reinterpret_cast<float*>(p.out) + (long long)b * p.out_bs + (long long)m * p.ldo + n;
    if constexpr (EPI == DICOW_EPI_RESIDUAL_F32) {
      const float alpha = (p.gate != nullptr) ? tanhf(__ldg(p.gate)) : 1.0f;
      const float* r = p.resid + (long long)b * p.resid_bs + (long long)m * p.ldr + n;
      if (ncols == 32 && p.out_vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 rv = *reinterpret_cast<const float4*>(r + j);
          v[j] = fmaf(alpha, v[j], rv.x), v[j + 1] = fmaf(alpha, v[j + 1], rv.y);
          v[j + 2] = fmaf(alpha, v[j + 2], rv.z), v[j + 3] = fmaf(alpha, v[j + 3], rv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) v[j] = fmaf(alpha, v[j], r[j]);
      }
    }
    if constexpr (EPI == DICOW_EPI_ACCUM_F32) {
      // out += alpha * acc (gradient accumulation / split-K partial sums): fp32 reductions at L2, no read-back
      const float alpha = (p.gate != nullptr) ? __ldg(p.gate) : 1.0f;
      if (ncols == 32 && p.out_vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(alpha * v[j]),
                       "f"(alpha * v[j + 1]), "f"(alpha * v[j + 2]), "f"(alpha * v[j + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) atomicAdd(o + j, alpha * v[j]);
      }
      return;
    }
    if constexpr (EPI == DICOW_EPI_GELU_FDDT_POS_F32) {
      // FDDT.forward (src/models/dicow/FDDT.py:52-62): sum_c (w_c * x + b_c) * m_c, classes S,T,N,O
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < ncols) {
          float w = 0.f, bb = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (p.fddt_w != nullptr) w = fmaf(mask[c], __ldg(p.fddt_w + (long long)c * p.N + n + j), w);
            bb = fmaf(mask[c], __ldg(p.fddt_b + (long long)c * p.N + n + j), bb);
          }
          if (p.fddt_w == nullptr) w = 1.f;  // bias-only FDDT (FDDT.py:43-51)
          float x = fmaf(v[j], w, bb);
          if (p.pos != nullptr) x += __ldg(p.pos + (long long)m * p.N + n + j);
          v[j] = x;
        }
      }
    }
    if (ncols == 32 && p.out_vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = v[j];
    }
  }
}

// ---- drain one accumulator: NCH chunks of 32 columns for this warp's 32 rows ------------------------------------------
// Software-pipelined: the tcgen05.ld of chunk c + 1 is in flight while chunk c is processed, the chunk's bias is requested
// before the wait on its TMEM load, and the accumulator is handed back to the MMA warp (release()) as soon as the LAST
// chunk is in registers -- before its math and stores.  (r01 profile of the GELU GEMM: 34 % of the stall samples on the
// first bias add / TMEM wait of each chunk, 11 % on the release fence of the hand-back, tensor pipe 69 % active.)
template <int EPI, int NCH, typename Release>
__device__ __forceinline__ void drain_accumulator(const GemmParams& p, uint32_t taddr, int b, int m, int n_base, bool row_ok,
                                                  const float (&mask)[4], Release release) {
  int nvalid = (p.N - n_base + 31) / 32;  // warp-uniform
  nvalid = nvalid > NCH ? NCH : nvalid;
  if (nvalid <= 0) {
    release();
    return;
  }
  if (p.serial_drain) {  // default (see dicow_gemm_bf16): one chunk at a time
#pragma unroll 1
    for (int c = 0; c < nvalid; ++c) {
      uint32_t r1[32];
      tmem_ld_x32(taddr + c * 32, r1);
      tmem_ld_wait_regs(r1);
      if (row_ok) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
        epilogue_chunk<EPI>(p, v, b, m, n_base + c * 32, min(32, p.N - n_base - c * 32), mask);
      }
    }
    __threadfence_block();
    release();
    return;
  }
  const bool vec_bias = p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && (n_base & 3) == 0;
  uint32_t r[2][32];
  tmem_ld_x32(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c < nvalid) {
      const int n = n_base + c * 32;
      const int ncols = min(32, p.N - n);
      float4 bv[8];
      const bool pre = vec_bias && ncols == 32;
      if (pre) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
      }
      tmem_ld_wait_regs(r[c & 1]);
      if (c + 1 < nvalid) tmem_ld_x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
      else release();
      if (row_ok) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c & 1][j]);
        epilogue_chunk<EPI>(p, v, b, m, n, ncols, mask, pre ? bv : nullptr);
      }
    }
  }
}

// ---- drain one accumulator through shared memory + TMA stores (CTA-pair kernel, bf16 outputs) ----------------------------
// Each epilogue warp owns 32 accumulator rows x (BN / 2) columns.  Two 32-column chunks at a time go
// TMEM -> registers -> bias (/ GELU) -> bf16 -> a [32 rows x 64 columns] tile in shared memory laid out as the output tensor
// map's 128B-swizzled box (16-byte chunk index XOR row % 8: the 8 threads of a store phase hit 8 different bank groups), then
// ONE cp.async.bulk.tensor store writes the 4 KB tile as full 128-byte lines.  The per-thread st.global.v4 path it replaces
// touched 32 different lines per store instruction, 16 bytes each (8 partial writes per line).  The warp's two staging
// tiles alternate; a tile is reused once the bulk group that reads it has finished reading (wait_group.read 1).
// Rows / columns beyond the tensor are clipped by the TMA unit.
template <int EPI, int NCH, typename Release>
__device__ __forceinline__ void drain_accumulator_tma(const GemmParams& p, const CUtensorMap* tmO, const CUtensorMap* tmX,
                                                      uint8_t* stage_base, uint32_t taddr, int b, int m_warp, int n_base,
                                                      int lane, Release release) {
  static_assert(NCH % 2 == 0, "pairs of 32-column chunks");
  // GELU_SAVE writes two tensors (pre-activation -> tmX, activation -> tmO): one staging tile each, both reused every pair
  constexpr bool kTwoOut = EPI == DICOW_EPI_GELU_SAVE_BF16;
  int nvalid = (p.N - n_base + 31) / 32;  // warp-uniform
  nvalid = nvalid > NCH ? NCH : nvalid;
  if (nvalid <= 0) {
    release();
    return;
  }
#pragma unroll 1
  for (int pair = 0; pair * 2 < nvalid; ++pair) {
    uint8_t* stage = kTwoOut ? stage_base + 4096 : stage_base + (pair & 1) * 4096;
    if (lane == 0) {  // the store(s) that read this tile have finished reading it
      if constexpr (kTwoOut) bulk_wait_group_read<0>();
      else bulk_wait_group_read<1>();
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = pair * 2 + h;
      if (c < nvalid) {
        uint32_t r[32];
        tmem_ld_x32(taddr + c * 32, r);
        tmem_ld_wait_regs(r);
        if (c == nvalid - 1) release();  // last chunk in registers: the MMA warp may overwrite the accumulator
        const int n = n_base + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          if (n + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              v[j] += bv.x, v[j + 1] += bv.y, v[j + 2] += bv.z, v[j + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) v[j] += __ldg(p.bias + n + j);
          }
        }
        if constexpr (kTwoOut) {  // the pre-activation, as the backward's gelu' wants it
          uint8_t* row = stage_base + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16(v[8 * j], v[8 * j + 1]), q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]), q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(row + (((h * 4 + j) ^ (lane & 7)) << 4)) = q;
          }
        }
        if constexpr (EPI == DICOW_EPI_DGELU_BF16) {  // out = acc * gelu'(pre)
          const int m = m_warp + lane;
          if (m < p.Mb) {
            const __nv_bfloat16* a = p.aux + (long long)b * p.out_bs + (long long)m * p.ldo + n;
            if (n + 32 <= p.N) {
              uint4 q[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) q[j] = __ldg(reinterpret_cast<const uint4*>(a) + j);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                  v[8 * j + 2 * k] *= dgelu_erf_fast(f.x);
                  v[8 * j + 2 * k + 1] *= dgelu_erf_fast(f.y);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n + j < p.N) v[j] *= dgelu_erf_fast(__bfloat162float(a[j]));
            }
          }
        }
        if constexpr (EPI == DICOW_EPI_BIAS_GELU_BF16 || EPI == DICOW_EPI_GELU_SAVE_BF16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        }
        uint8_t* row = stage + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16(v[8 * j], v[8 * j + 1]), q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
          q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]), q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(row + (((h * 4 + j) ^ (lane & 7)) << 4)) = q;
        }
      }
    }
    fence_proxy_async_smem();  // the generic-proxy writes above -> visible to the TMA (async proxy) read
    __syncwarp();
    if (lane == 0) {
      if constexpr (kTwoOut) tma_store_3d(tmX, stage_base, n_base + pair * 64, m_warp, b);
      tma_store_3d(tmO, stage, n_base + pair * 64, m_warp, b);
      bulk_commit_group();
    }
  }
}

// A_MN / B_MN: the operand is given transposed in memory (At[k][m] / Wt[k][n], contraction index on the rows) and is
// consumed MN-major: the stage holds 64 x 64 atoms ([64 contraction rows] x [64 output elements = 128 B], 8 KB each, one
// TMA box per atom), descriptors use SBO = 1024 (8 contraction rows), LBO = 8192 (next 64-element atom).  This is what
// the backward GEMMs need without any transposed copies: dgrad dX = dY W (W row-major [N, K] is the MN-major B operand),
// wgrad dW = dY^T X (both operands MN-major, contraction over the rows).
template <int BN, int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (p.K1 > 0) tma_prefetch_desc(&tmA2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < p.total_tiles * p.splits; work += gridDim.x) {
      const int tile = work % p.total_tiles, split = work / p.total_tiles;
      const int nt = tile % p.tiles_n;
      const int mt = tile / p.tiles_n;
      const int b = mt / p.tiles_m;
      const int m0 = (mt % p.tiles_m) * BM;
      const int n0 = nt * BN;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a) tma_load_3d(&tmA, &full_bar[stage], sA + a * 8192, m0 + 64 * a, k0, b);
          } else {
            if (p.K1 > 0 && k0 >= p.K1)
              tma_load_3d(&tmA2, &full_bar[stage], sA, k0 - p.K1, m0, b);
            else
              tma_load_3d(&tmA, &full_bar[stage], sA, k0, m0, b);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int a = 0; a < BN / 64; ++a)
              tma_load_2d(&tmW, &full_bar[stage], sB + a * 8192, n0 + 64 * a, k0, kEvictLast);
          } else {
            tma_load_2d(&tmW, &full_bar[stage], sB, k0, n0, kEvictLast);
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop; one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint64_t a_step = A_MN ? 128 : 2;  // per 16-deep MMA: 16 contraction rows (2048 B) or 32 B along K
    constexpr uint64_t b_step = B_MN ? 128 : 2;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int work = blockIdx.x; work < p.total_tiles * p.splits; work += gridDim.x, ++it) {
      const int split = work / p.total_tiles;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da = A_MN ? make_sdesc_sw128(a_addr, 1024, 8192) : make_sdesc_sw128(a_addr, 1024, 0);
        const uint64_t db = B_MN ? make_sdesc_sw128(a_addr + Cfg::A_BYTES, 1024, 8192)
                                 : make_sdesc_sw128(a_addr + Cfg::A_BYTES, 1024, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(d_tmem, da + (uint64_t)k * a_step, db + (uint64_t)k * b_step, idesc,
                         (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp >= kNonEpiWarps) {
    // ===================== epilogue =====================
    const int e = warp - kNonEpiWarps;
    const int quad = e & 3;  // == warp % 4: the TMEM lane quadrant this warp may access
    const int half = e >> 2; // which half of the BN columns
    int it = 0;
    for (int work = blockIdx.x; work < p.total_tiles * p.splits; work += gridDim.x, ++it) {
      const int tile = work % p.total_tiles;
      const int nt = tile % p.tiles_n;
      const int mt = tile / p.tiles_n;
      const int b = mt / p.tiles_m;
      const int m = (mt % p.tiles_m) * BM + quad * 32 + lane;
      const int n_base = nt * BN + half * (BN / 2);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool row_ok = m < p.Mb;

      float mask[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (EPI == DICOW_EPI_GELU_FDDT_POS_F32) {
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) mask[c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.Mb + m);
        }
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + half * (BN / 2);
      drain_accumulator<EPI, BN / 64>(p, taddr, b, m, n_base, row_ok, mask, [&] {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&tempty_bar[acc]);
      });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant: a cluster of two CTAs (two SMs of one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2
// (M = 256).  CTA r stages rows [128 r, 128 r + 128) of the A tile and rows [BN/2 r, BN/2 r + BN/2) of the W tile, so
// each SM pulls 32 KB instead of 48 KB per 64-deep k block from L2 (the single-CTA kernel sits at the L2 -> SM
// throughput limit: ncu r01 shows 72-74 % tensor-pipe activity with every stage wait on the TMA side).  The leader CTA
// (rank 0) issues all MMAs; TMA transaction bytes of both CTAs are signalled on the leader's full barrier; commits are
// multicast to the empty / accumulator-full barriers of both CTAs; both CTAs' epilogue warps drain their own 128
// accumulator rows and release the accumulator on the leader's barrier.
// ------------------------------------------------------------------------------------------------------------------
template <int BN>
struct GemmCfg2 {
  static constexpr int STAGES = 5;  // (6 before the epilogue staging tiles took 64 KB)
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static constexpr uint32_t BAR_BYTES = 1024;                      // barriers + TMEM slot, keeps the staging tiles 1024-aligned
  static constexpr uint32_t EPI_BYTES = kEpiWarps * 2 * 4096;      // per epilogue warp: two [32 x 64] bf16 staging tiles
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024 /*align slack*/;
};

// A_MN / B_MN as in the single-CTA kernel: CTA r stages its 128 output rows of the transposed A as two 64 x 64 atoms and its
// BN / 2 columns of the transposed W as BN / 128 atoms; the descriptors are the single-CTA ones (every CTA of the pair
// describes its own half).  Split-K (ACCUM_F32): work item = (tile, split), partial sums reduced with fp32 red.global.
template <int BN, int EPI, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  using Cfg = GemmCfg2<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready (both CTAs, multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained (leader's copy counts both CTAs' epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster = (int)cluster_id_x();
  const int nclusters = (int)cluster_nctaid_x();
  const int total_work = p.total_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (p.K1 > 0) tma_prefetch_desc(&tmA2);
    if (p.tma_out) tma_prefetch_desc(&tmO);
    if (p.tma_out && EPI == DICOW_EPI_GELU_SAVE_BF16) tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its half of the pair's stage) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int work = cluster; work < total_work; work += nclusters) {
      const int tile = work % p.total_tiles, split = work / p.total_tiles;
      const int nt = tile % p.tiles_n;
      const int mt = tile / p.tiles_n;
      const int b = mt / p.tiles_m;
      const int m0 = (mt % p.tiles_m) * (2 * BM) + (int)rank * BM;
      const int n0 = nt * BN + (int)rank * (BN / 2);
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a) tma_load_3d_2cta(&tmA, bar, sA + a * 8192, m0 + 64 * a, k0, b);
          } else {
            if (p.K1 > 0 && k0 >= p.K1)
              tma_load_3d_2cta(&tmA2, bar, sA, k0 - p.K1, m0, b);
            else
              tma_load_3d_2cta(&tmA, bar, sA, k0, m0, b);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int a = 0; a < BN / 128; ++a) tma_load_2d_2cta(&tmW, bar, sB + a * 8192, n0 + 64 * a, k0, kEvictLast);
          } else {
            tma_load_2d_2cta(&tmW, bar, sB, k0, n0, kEvictLast);
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint64_t a_step = A_MN ? 128 : 2;  // per 16-deep MMA: 16 contraction rows (2048 B) or 32 B along K
      constexpr uint64_t b_step = B_MN ? 128 : 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int work = cluster; work < total_work; work += nclusters, ++it) {
        const int split = work / p.total_tiles;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = A_MN ? make_sdesc_sw128(a_addr, 1024, 8192) : make_sdesc_sw128(a_addr, 1024, 0);
          const uint64_t db = B_MN ? make_sdesc_sw128(a_addr + Cfg::A_BYTES, 1024, 8192)
                                   : make_sdesc_sw128(a_addr + Cfg::A_BYTES, 1024, 0);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_2cta(d_tmem, da + (uint64_t)k * a_step, db + (uint64_t)k * b_step, idesc,
                                (kb != kb0 || k != 0) ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit_2cta(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp >= kNonEpiWarps) {
    // ===================== epilogue (each CTA drains its own 128 accumulator rows) =====================
    const int e = warp - kNonEpiWarps;
    const int quad = e & 3;
    const int half = e >> 2;
    int it = 0;
    for (int work = cluster; work < total_work; work += nclusters, ++it) {
      const int tile = work % p.total_tiles;
      const int nt = tile % p.tiles_n;
      const int mt = tile / p.tiles_n;
      const int b = mt / p.tiles_m;
      const int m = (mt % p.tiles_m) * (2 * BM) + (int)rank * BM + quad * 32 + lane;
      const int n_base = nt * BN + half * (BN / 2);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool row_ok = m < p.Mb;

      float mask[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (EPI == DICOW_EPI_GELU_FDDT_POS_F32) {
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) mask[c] = __ldg(p.stno + (long long)b * p.stno_bs + (long long)c * p.Mb + m);
        }
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + half * (BN / 2);
      const uint32_t tempty_remote = map_to_cta(smem_u32(&tempty_bar[acc]), 0);
      auto release = [&] {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(tempty_remote);
      };
      if constexpr (EPI == DICOW_EPI_BIAS_BF16 || EPI == DICOW_EPI_BIAS_GELU_BF16 || EPI == DICOW_EPI_GELU_SAVE_BF16 ||
                    EPI == DICOW_EPI_DGELU_BF16) {
        if (p.tma_out) {
          uint8_t* stage = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + e * 2 * 4096;
          drain_accumulator_tma<EPI, BN / 64>(p, &tmO, &tmX, stage, taddr, b, m - lane, n_base, lane, release);
          continue;
        }
      }
      drain_accumulator<EPI, BN / 64>(p, taddr, b, m, n_base, row_ok, mask, release);
    }
    if (p.tma_out && lane == 0) bulk_wait_group<0>();  // shared memory must outlive the last bulk stores
  }

  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still address this CTA's barriers or TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, bool A_MN = false, bool B_MN = false>
int launch_gemm(dicow_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmW,
                const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kfn = gemm_bf16_kernel<BN, EPI, A_MN, B_MN>;
  static DeviceOnce attr_once;  // per instantiation
  if (attr_once.first(ctx)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int work = p.total_tiles * p.splits;
  const int grid = work < ctx->num_sms ? work : ctx->num_sms;
  kfn<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmA2, tmW, p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

template <int BN, int EPI, bool A_MN = false, bool B_MN = false>
int launch_gemm_2cta(dicow_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmW,
                     const CUtensorMap& tmO, const CUtensorMap& tmX, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg2<BN>;
  auto kfn = gemm_bf16_2cta_kernel<BN, EPI, A_MN, B_MN>;
  static DeviceOnce attr_once;  // per instantiation
  if (attr_once.first(ctx)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int work = p.total_tiles * p.splits;
  int grid = 2 * (work < ctx->num_sms / 2 ? work : ctx->num_sms / 2);
  kfn<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmA2, tmW, tmO, tmX, p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

// which (epilogue, operand layout) combinations exist as CTA-pair instantiations
bool has_2cta(int epi, bool a_t, bool w_t) {
  if (!a_t && !w_t)
    return epi == DICOW_EPI_BIAS_BF16 || epi == DICOW_EPI_BIAS_GELU_BF16 || epi == DICOW_EPI_RESIDUAL_F32 ||
           epi == DICOW_EPI_BIAS_F32 || epi == DICOW_EPI_GELU_FDDT_POS_F32 || epi == DICOW_EPI_GELU_SAVE_BF16 ||
           epi == DICOW_EPI_ACCUM_F32;
  if (!a_t && w_t) return epi == DICOW_EPI_BIAS_BF16 || epi == DICOW_EPI_DGELU_BF16 || epi == DICOW_EPI_ACCUM_F32;
  if (a_t && w_t) return epi == DICOW_EPI_ACCUM_F32;
  return false;
}

int dispatch_epi_2cta(dicow_ctx* ctx, int epi, bool a_t, bool w_t, const CUtensorMap& tmA, const CUtensorMap& tmA2,
                      const CUtensorMap& tmW, const CUtensorMap& tmO, const CUtensorMap& tmX, const GemmParams& p,
                      cudaStream_t stream) {
#define DICOW_2CTA(E, AT, WT) return launch_gemm_2cta<256, E, AT, WT>(ctx, tmA, tmA2, tmW, tmO, tmX, p, stream)
  if (!a_t && !w_t) {
    switch (epi) {
      case DICOW_EPI_BIAS_BF16: DICOW_2CTA(DICOW_EPI_BIAS_BF16, false, false);
      case DICOW_EPI_BIAS_GELU_BF16: DICOW_2CTA(DICOW_EPI_BIAS_GELU_BF16, false, false);
      case DICOW_EPI_RESIDUAL_F32: DICOW_2CTA(DICOW_EPI_RESIDUAL_F32, false, false);
      case DICOW_EPI_BIAS_F32: DICOW_2CTA(DICOW_EPI_BIAS_F32, false, false);
      case DICOW_EPI_GELU_FDDT_POS_F32: DICOW_2CTA(DICOW_EPI_GELU_FDDT_POS_F32, false, false);
      case DICOW_EPI_GELU_SAVE_BF16: DICOW_2CTA(DICOW_EPI_GELU_SAVE_BF16, false, false);
      case DICOW_EPI_ACCUM_F32: DICOW_2CTA(DICOW_EPI_ACCUM_F32, false, false);
      default: break;
    }
  } else if (!a_t && w_t) {
    switch (epi) {
      case DICOW_EPI_BIAS_BF16: DICOW_2CTA(DICOW_EPI_BIAS_BF16, false, true);
      case DICOW_EPI_DGELU_BF16: DICOW_2CTA(DICOW_EPI_DGELU_BF16, false, true);
      case DICOW_EPI_ACCUM_F32: DICOW_2CTA(DICOW_EPI_ACCUM_F32, false, true);
      default: break;
    }
  } else if (a_t && w_t) {
    if (epi == DICOW_EPI_ACCUM_F32) DICOW_2CTA(DICOW_EPI_ACCUM_F32, true, true);
  }
#undef DICOW_2CTA
  return set_error(ctx, DICOW_ERR_INVALID_ARG, "dicow_gemm_bf16: no CTA-pair kernel for epilogue %d (transposed %d, %d)", epi,
                   (int)a_t, (int)w_t);
}

// Split-K factor of an accumulating GEMM: the one that minimises rounds x (k blocks per item + a per-item epilogue cost) over
// `workers` persistent CTAs (CTA pairs).  "About two waves" (round 1) put e.g. 300 items on 148 CTAs: three rounds for 2.03
// waves of work.
int choose_splits(int tiles, int kblocks, int workers) {
  const double epi_cost = 4.0;  // in k blocks: the exposed part of an item's drain + pipeline refill
  int best = 1;
  double best_cost = 1e30;
  const int smax = kblocks < 48 ? kblocks : 48;
  for (int s = 1; s <= smax; ++s) {
    const int per = ceil_div(kblocks, s);
    if (per < 6 && s > 1) break;  // keep the mainloop of an item longer than its drain
    const int s_eff = ceil_div(kblocks, per);
    const long long items = (long long)tiles * s_eff;
    const long long rounds = (items + workers - 1) / workers;
    const double cost = (double)rounds * (per + epi_cost);
    if (cost < best_cost * 0.995) best_cost = cost, best = s_eff;
  }
  return best;
}

template <int BN>
int dispatch_epi(dicow_ctx* ctx, int epi, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmW,
                 const GemmParams& p, cudaStream_t stream) {
  switch (epi) {
    case DICOW_EPI_BIAS_BF16: return launch_gemm<BN, DICOW_EPI_BIAS_BF16>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_BIAS_GELU_BF16: return launch_gemm<BN, DICOW_EPI_BIAS_GELU_BF16>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_RESIDUAL_F32: return launch_gemm<BN, DICOW_EPI_RESIDUAL_F32>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_BIAS_F32: return launch_gemm<BN, DICOW_EPI_BIAS_F32>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_GELU_FDDT_POS_F32:
      return launch_gemm<BN, DICOW_EPI_GELU_FDDT_POS_F32>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_ACCUM_F32: return launch_gemm<BN, DICOW_EPI_ACCUM_F32>(ctx, tmA, tmA2, tmW, p, stream);
    case DICOW_EPI_GELU_SAVE_BF16: return launch_gemm<BN, DICOW_EPI_GELU_SAVE_BF16>(ctx, tmA, tmA2, tmW, p, stream);
    default: return set_error(ctx, DICOW_ERR_INVALID_ARG, "dicow_gemm_bf16: unknown epilogue %d", epi);
  }
}

// backward-pass variants (MN-major operands): dgrad (W transposed) writes bf16 / fp32, wgrad (both transposed) accumulates
template <int BN>
int dispatch_transposed(dicow_ctx* ctx, int epi, bool a_t, bool w_t, const CUtensorMap& tmA, const CUtensorMap& tmW,
                        const GemmParams& p, cudaStream_t stream) {
  if (!a_t && w_t) {
    switch (epi) {
      case DICOW_EPI_BIAS_BF16: return launch_gemm<BN, DICOW_EPI_BIAS_BF16, false, true>(ctx, tmA, tmA, tmW, p, stream);
      case DICOW_EPI_BIAS_F32: return launch_gemm<BN, DICOW_EPI_BIAS_F32, false, true>(ctx, tmA, tmA, tmW, p, stream);
      case DICOW_EPI_ACCUM_F32: return launch_gemm<BN, DICOW_EPI_ACCUM_F32, false, true>(ctx, tmA, tmA, tmW, p, stream);
      case DICOW_EPI_DGELU_BF16: return launch_gemm<BN, DICOW_EPI_DGELU_BF16, false, true>(ctx, tmA, tmA, tmW, p, stream);
      default: break;
    }
  } else if (a_t && w_t) {
    switch (epi) {
      case DICOW_EPI_ACCUM_F32: return launch_gemm<BN, DICOW_EPI_ACCUM_F32, true, true>(ctx, tmA, tmA, tmW, p, stream);
      case DICOW_EPI_BIAS_F32: return launch_gemm<BN, DICOW_EPI_BIAS_F32, true, true>(ctx, tmA, tmA, tmW, p, stream);
      default: break;
    }
  } else if (a_t && !w_t) {
    switch (epi) {
      case DICOW_EPI_ACCUM_F32: return launch_gemm<BN, DICOW_EPI_ACCUM_F32, true, false>(ctx, tmA, tmA, tmW, p, stream);
      case DICOW_EPI_BIAS_BF16: return launch_gemm<BN, DICOW_EPI_BIAS_BF16, true, false>(ctx, tmA, tmA, tmW, p, stream);
      default: break;
    }
  }
  return set_error(ctx, DICOW_ERR_UNSUPPORTED, "dicow_gemm_bf16: epilogue %d is not built for transposed operands (%d, %d)", epi,
                   (int)a_t, (int)w_t);
}

}  // namespace

}  // namespace dicow

using namespace dicow;

extern "C" int dicow_gemm_bf16(dicow_handle_t h, const dicow_gemm_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_gemm_args_t),
                "dicow_gemm_bf16: bad args struct (size %zu, expected %zu)", a ? a->struct_size : 0,
                sizeof(dicow_gemm_args_t));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DICOW_REQUIRE(ctx, a->nb >= 1 && a->Mb >= 1 && a->N >= 1 && a->K >= 8, "dicow_gemm_bf16: bad shape nb=%d Mb=%d N=%d K=%d",
                a->nb, a->Mb, a->N, a->K);
  DICOW_REQUIRE(ctx, a->A && a->W && a->out, "dicow_gemm_bf16: null operand");
  DICOW_REQUIRE(ctx, (a->lda % 8) == 0 && (a->ldw % 8) == 0 && (a->a_batch_stride % 8) == 0,
                "dicow_gemm_bf16: lda/ldw/a_batch_stride must be multiples of 8 elements (16-byte TMA strides)");
  DICOW_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->W) % 16) == 0,
                "dicow_gemm_bf16: A/W must be 16-byte aligned");
  const bool split = a->A2 != nullptr;
  if (split) {
    DICOW_REQUIRE(ctx, a->K1 > 0 && a->K1 < a->K && (a->K1 % BK) == 0 && (a->lda2 % 8) == 0 &&
                           (a->a2_batch_stride % 8) == 0 && (reinterpret_cast<uintptr_t>(a->A2) % 16) == 0,
                  "dicow_gemm_bf16: bad split-K source (K1=%d)", a->K1);
  }
  if (a->epilogue == DICOW_EPI_RESIDUAL_F32) DICOW_REQUIRE(ctx, a->resid != nullptr, "dicow_gemm_bf16: resid is NULL");
  if (a->epilogue == DICOW_EPI_GELU_FDDT_POS_F32)
    DICOW_REQUIRE(ctx, a->stno && a->fddt_b, "dicow_gemm_bf16: FDDT epilogue needs stno / fddt_b (fddt_w NULL = bias-only)");

  const bool a_t = (a->flags & 4) != 0;  // A given as At[k][m] (row stride lda)
  const bool w_t = (a->flags & 8) != 0;  // W given as Wt[k][n] (row stride ldw)
  DICOW_REQUIRE(ctx, !(a_t && split), "dicow_gemm_bf16: a transposed A cannot be split over two sources");
  DICOW_REQUIRE(ctx, (!a_t || (a->Mb % 8) == 0) && (!w_t || (a->N % 8) == 0),
                "dicow_gemm_bf16: transposed operands need Mb / N multiples of 8");
  const int BN = a->N >= 256 ? 256 : 128;
  // CTA pairs (256-row tiles) when every pair gets at least ~2 tiles (split-K GEMMs: whenever a tile has 256 rows to fill);
  // `flags & 1` forces the single-CTA kernel, `flags & 2` forces pairs (tests / comparison).  DICOW_GEMM_2CTA_BWD=0 keeps
  // the round-1 behaviour (backward / training epilogues on the single-CTA kernel) for A/B measurements.
  static const int bwd_pairs = [] {
    const char* e = getenv("DICOW_GEMM_2CTA_BWD");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  const long long tiles128 = (long long)a->nb * ceil_div(a->Mb, BM) * ceil_div(a->N, BN);
  const bool auto_split = a->epilogue == DICOW_EPI_ACCUM_F32 && a->splits != 1;
  bool two_cta = BN == 256 && has_2cta(a->epilogue, a_t, w_t) &&
                 (auto_split ? a->Mb >= 2 * BM : tiles128 >= 2 * (long long)ctx->num_sms);
  if (!bwd_pairs && (a_t || w_t || a->epilogue >= DICOW_EPI_ACCUM_F32)) two_cta = false;
  if (a->flags & 1) two_cta = false;
  if ((a->flags & 2) && BN == 256 && has_2cta(a->epilogue, a_t, w_t)) two_cta = true;
  GemmParams p{};
  p.nb = a->nb, p.Mb = a->Mb, p.N = a->N, p.K = a->K, p.K1 = split ? a->K1 : 0;
  p.tiles_m = ceil_div(a->Mb, two_cta ? 2 * BM : BM);
  p.tiles_n = ceil_div(a->N, BN);
  p.total_tiles = a->nb * p.tiles_m * p.tiles_n;
  p.kblocks = ceil_div(a->K, BK);
  p.splits = 1;
  if (auto_split) {  // explicit a->splits > 1 overrides
    int want = a->splits > 1 ? a->splits : choose_splits(p.total_tiles, p.kblocks, two_cta ? ctx->num_sms / 2 : ctx->num_sms);
    want = want < 1 ? 1 : (want > p.kblocks ? p.kblocks : want);
    p.splits = want;
  }
  p.kb_per_split = ceil_div(p.kblocks, p.splits);
  p.splits = ceil_div(p.kblocks, p.kb_per_split);
  p.bias = a->bias;
  p.out = a->out, p.ldo = a->ldo, p.out_bs = a->out_batch_stride;
  p.resid = a->resid, p.ldr = a->ldr, p.resid_bs = a->resid_batch_stride, p.gate = a->gate;
  p.stno = a->stno, p.stno_bs = a->stno_batch_stride, p.fddt_w = a->fddt_w, p.fddt_b = a->fddt_b, p.pos = a->pos;
  p.aux = reinterpret_cast<__nv_bfloat16*>(a->aux_bf16);
  // default: the serial drain.  A/B on one box, interleaved, 3 runs each (tools/gpu_ab.sh): serial 398.9 / 402.6 / 405.2
  // utt/s (GEMM family 47.8 / 47.3 / 47.0 ms per step), pipelined 395.8 / 399.0 / 398.3 (48.7 / 48.4 / 48.4 ms) -- the
  // step is power-capped, the pipelined schedule removes stalls but not energy and runs 2.5 % slower.
  static const int serial_drain = [] {
    const char* e = getenv("DICOW_GEMM_PIPELINED_DRAIN");
    return (e != nullptr && e[0] == '1') ? 0 : 1;
  }();
  p.serial_drain = serial_drain;
  const bool out_is_bf16 = a->epilogue == DICOW_EPI_BIAS_BF16 || a->epilogue == DICOW_EPI_BIAS_GELU_BF16 ||
                           a->epilogue == DICOW_EPI_GELU_SAVE_BF16 || a->epilogue == DICOW_EPI_DGELU_BF16;
  if (a->epilogue == DICOW_EPI_GELU_SAVE_BF16 || a->epilogue == DICOW_EPI_DGELU_BF16)
    DICOW_REQUIRE(ctx, a->aux_bf16 != nullptr && (reinterpret_cast<uintptr_t>(a->aux_bf16) % 16) == 0,
                  "dicow_gemm_bf16: this epilogue needs aux_bf16 (same leading dimension as out)");
  const int vec = out_is_bf16 ? 8 : 4;
  bool vec_ok = (a->ldo % vec) == 0 && (a->out_batch_stride % vec) == 0 &&
                (reinterpret_cast<uintptr_t>(a->out) % 16) == 0;
  if (a->bias) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(a->bias) % 16) == 0;
  if (a->epilogue == DICOW_EPI_RESIDUAL_F32)
    vec_ok = vec_ok && (a->ldr % 4) == 0 && (a->resid_batch_stride % 4) == 0 &&
             (reinterpret_cast<uintptr_t>(a->resid) % 16) == 0;
  p.out_vec_ok = vec_ok ? 1 : 0;

  CUtensorMap tmA, tmA2, tmW;
  const int Ka = split ? a->K1 : a->K;
  if (a_t) {  // At[k][m]: inner dimension = output rows, boxes of 64 x 64
    uint64_t dims[3] = {(uint64_t)a->Mb, (uint64_t)a->K, (uint64_t)a->nb};
    uint64_t bs = a->nb > 1 ? (uint64_t)a->a_batch_stride : (uint64_t)a->lda * (uint64_t)a->K;
    uint64_t strides[2] = {(uint64_t)a->lda * 2, bs * 2};
    uint32_t box[3] = {64, BK, 1};
    int rc = make_tmap_bf16(ctx, &tmA, a->A, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    uint64_t dims[3] = {(uint64_t)Ka, (uint64_t)a->Mb, (uint64_t)a->nb};
    uint64_t bs = a->nb > 1 ? (uint64_t)a->a_batch_stride : (uint64_t)a->lda * (uint64_t)a->Mb;
    uint64_t strides[2] = {(uint64_t)a->lda * 2, bs * 2};
    uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(ctx, &tmA, a->A, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (split) {
    uint64_t dims[3] = {(uint64_t)(a->K - a->K1), (uint64_t)a->Mb, (uint64_t)a->nb};
    uint64_t bs = a->nb > 1 ? (uint64_t)a->a2_batch_stride : (uint64_t)a->lda2 * (uint64_t)a->Mb;
    uint64_t strides[2] = {(uint64_t)a->lda2 * 2, bs * 2};
    uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(ctx, &tmA2, a->A2, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
  }
  if (w_t) {  // Wt[k][n]
    uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->K};
    uint64_t strides[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {64, BK};
    int rc = make_tmap_bf16(ctx, &tmW, a->W, 2, dims, strides, box);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)(two_cta ? BN / 2 : BN)};
    int rc = make_tmap_bf16(ctx, &tmW, a->W, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (two_cta) {
    // output tensor map(s) for the TMA-store epilogue: [N, Mb, nb] bf16, box 64 columns x 32 rows (one epilogue warp's tile)
    CUtensorMap tmO = tmA, tmX = tmA;
    static const int tma_store = [] {
      const char* e = getenv("DICOW_GEMM_TMA_STORE");
      return (e != nullptr && e[0] == '0') ? 0 : 1;
    }();
    if (tma_store && out_is_bf16 && (a->ldo % 8) == 0 && (a->out_batch_stride % 8) == 0 &&
        (reinterpret_cast<uintptr_t>(a->out) % 16) == 0 && (a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) % 16) == 0)) {
      uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->Mb, (uint64_t)a->nb};
      uint64_t bs = a->nb > 1 ? (uint64_t)a->out_batch_stride : (uint64_t)a->ldo * (uint64_t)a->Mb;
      uint64_t strides[2] = {(uint64_t)a->ldo * 2, bs * 2};
      uint32_t box[3] = {64, 32, 1};
      int rc = make_tmap_bf16(ctx, &tmO, a->out, 3, dims, strides, box);
      if (rc) return rc;
      if (a->epilogue == DICOW_EPI_GELU_SAVE_BF16) {  // the saved pre-activation: same shape and strides as out
        rc = make_tmap_bf16(ctx, &tmX, a->aux_bf16, 3, dims, strides, box);
        if (rc) return rc;
      }
      p.tma_out = 1;
    }
    return dispatch_epi_2cta(ctx, a->epilogue, a_t, w_t, tmA, tmA2, tmW, tmO, tmX, p, stream);
  }
  if (a_t || w_t) {
    if (BN == 256) return dispatch_transposed<256>(ctx, a->epilogue, a_t, w_t, tmA, tmW, p, stream);
    return dispatch_transposed<128>(ctx, a->epilogue, a_t, w_t, tmA, tmW, p, stream);
  }
  if (BN == 256) return dispatch_epi<256>(ctx, a->epilogue, tmA, tmA2, tmW, p, stream);
  return dispatch_epi<128>(ctx, a->epilogue, tmA, tmA2, tmW, p, stream);
}
