// attention_bwd_fused.cu -- single-pass flash-attention backward (head_dim 64, non-causal) on tcgen05: 5 GEMMs per
// (key tile, query block) pair instead of the 7 of the two-pass kernels (attention_bwd.cu), one exponential per score.
//
// One persistent CTA per SM walks work items (batch, head, tile of 128 keys).  The keys are the TMEM lanes; query blocks
// of 128 stream through a 3-stage TMA ring:
//     S^T  = K Q_i^T          dP^T = V dO_i^T                 (SS MMAs, M128 N128 K64, fp32 in TMEM)
//     P^T  = 2^(S^T log2e - lse_q)        dS^T = P^T (dP^T - D_q)          (8 warps, thread = key row)
//     dV  += P^T dO_i         (A = P^T bf16 from TMEM, dO_i consumed MN-major)
//     dK  += dS^T Q_i         (A = dS^T bf16 from shared memory, K-major;  Q_i consumed MN-major)
//     dQ_i = dS_i K           (A = the SAME shared-memory tile consumed MN-major, K consumed MN-major)
// dV / dK stay in TMEM for the whole item; dQ_i leaves through TMA reductions (cp.reduce.async.bulk.tensor, fp32 add at
// L2) into a [B, Tq, H*64] workspace where the partial sums of the key tiles of one head meet, converted to bf16 by a small kernel afterwards.
//
// TMEM (512 columns): S^T 128 | dP^T 128 | P^T bf16 64 | dV 64 | dK 64 | dQ 64.
// Warps: 0-7 probabilities (warps 0-3: query columns 0-63, 4-7: 64-127), 8-11 dQ reduction + dV / dK epilogue,
// 12 TMA producer, 13 MMA issuer.  The issuer queues S^T / dP^T of block i + 1 BEFORE the three accumulating GEMMs of block
// i, so the probabilities of block i + 1 are computed while the tensor pipe works on block i.
//
// Replaces the autograd backward of the SDPA call inside HF WhisperAttention (HF:modeling_whisper.py:342-352) for the
// fine-tuning step, at the encoder's shapes; causal / short-query shapes (decoder) stay on the two-pass kernels.
#include <math.h>
#include <stdlib.h>

#include "attention_common.h"
#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int HD = 64;
constexpr int KV = 128;  // owned keys per item (TMEM lanes)
constexpr int QB = 128;  // streamed queries per block
constexpr int QSTAGES = 3;  // block i + 2 is requested while block i is still being accumulated
constexpr int kSoftmaxWarps = 8;
constexpr int kReduceWarps = 4;
constexpr int kThreads = (kSoftmaxWarps + kReduceWarps + 2) * 32;  // 448
constexpr int kProducerWarp = kSoftmaxWarps + kReduceWarps;
constexpr int kMmaWarp = kProducerWarp + 1;

constexpr uint32_t TILE_BYTES = 128 * HD * 2;  // 16 KB: one [128 rows x 64] bf16 tile
constexpr uint32_t DS_BYTES = 2 * TILE_BYTES;  // dS^T [128 keys x 128 queries] bf16: two 64-query atoms

// TMEM columns
constexpr uint32_t ST_COL = 0;
constexpr uint32_t DP_COL = 128;
constexpr uint32_t P_COL = 256;
constexpr uint32_t DV_COL = 320;
constexpr uint32_t DK_COL = 384;
constexpr uint32_t DQ_COL = 448;
constexpr uint32_t TMEM_COLS = 512;

// shared memory
constexpr uint32_t K_OFF = 0;                                    // [2 items] K (needed until the item's last dQ GEMM)
constexpr uint32_t V_OFF = K_OFF + 2 * TILE_BYTES;               // V (free once the item's last dP^T GEMM has been issued)
constexpr uint32_t RED_OFF = V_OFF + TILE_BYTES;                 // [4 reducer warps][32 rows x 32 fp32]: dQ staging tiles
constexpr uint32_t Q_OFF = RED_OFF + TILE_BYTES;                 // [QSTAGES][Q | dO]
constexpr uint32_t DS_OFF = Q_OFF + QSTAGES * 2 * TILE_BYTES;    // [2][dS^T]
constexpr uint32_t STAT_OFF = DS_OFF + 2 * DS_BYTES;             // [2][-lse | -D][QB] floats
constexpr uint32_t BAR_OFF = STAT_OFF + 2 * 2 * QB * 4;
constexpr uint32_t BAR_BYTES = 192;  // 19 mbarriers + the TMEM slot
constexpr uint32_t SMEM_BYTES = BAR_OFF + BAR_BYTES;  // 231 616 B of the 232 448 available: no slack for re-alignment (see below)

constexpr float kLog2e = 1.4426950408889634f;

struct FusedParams {
  int B, H, Tq, Tk, nkv, nq, total;
  const float* lse;  // [B, H, Tq] log2 units
  const float* D;    // [B, H, Tq]
  float* dq_acc;     // [B, Tq, H * 64] fp32, zeroed
  __nv_bfloat16* dK;
  __nv_bfloat16* dV;
  long long dkv_rs, dkv_bs;
  int dbg;  // DICOW_BWD_FUSED_DBG (measurement only): 1 = no dQ adds, 2 = no probability arithmetic
};

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                           const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                           const __grid_constant__ CUtensorMap tmAcc, const FusedParams p) {
  // the 128B-swizzled tiles need a 1024-byte aligned base; there is no room for an alignment pad, so the alignment is requested
  // from the compiler / driver and checked
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [2]
  uint64_t* k_empty = k_full + 2;                                  // [2]
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 1;
  uint64_t* q_full = v_empty + 1;                                  // [QSTAGES]
  uint64_t* q_empty = q_full + QSTAGES;                            // [QSTAGES]
  uint64_t* sdp_full = q_empty + QSTAGES;   // S^T / dP^T of a block in TMEM
  uint64_t* pds_full = sdp_full + 1;        // P^T (TMEM) / dS^T (shared memory) of a block written
  uint64_t* p_free = pds_full + 1;          // dV GEMM of a block retired: the P^T columns may be overwritten
  uint64_t* dq_full = p_free + 1;           // dQ_i in TMEM
  uint64_t* dq_empty = dq_full + 1;         // dQ_i in the reducers' registers
  uint64_t* acc_full = dq_empty + 1;        // dV / dK of the item complete
  uint64_t* acc_empty = acc_full + 1;       // dV / dK read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  static_assert((2 + 2 + 2 + 2 * QSTAGES + 7) * 8 + 4 <= BAR_BYTES, "barrier block");
  float* s_stat = reinterpret_cast<float*>(smem + STAT_OFF);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmAcc);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int s = 0; s < QSTAGES; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, kSoftmaxWarps);
    mbar_init(p_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, kReduceWarps);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kReduceWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const int nq = p.nq;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    int g = 0;
    int it = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const int kvt = w % p.nkv, bh = w / p.nkv;
      const int h = bh % p.H, b = bh / p.H;
      const int kvb = it & 1;
      mbar_wait(&k_empty[kvb], ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[kvb], TILE_BYTES);
        tma_load_4d(&tmK, &k_full[kvb], smem + K_OFF + kvb * TILE_BYTES, 0, kvt * KV, h, b, kEvictFirst);
      }
      __syncwarp();
      mbar_wait(v_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full, TILE_BYTES);
        tma_load_4d(&tmV, v_full, smem + V_OFF, 0, kvt * KV, h, b, kEvictFirst);
      }
      __syncwarp();
      for (int i = 0; i < nq; ++i, ++g) {
        const int st = g % QSTAGES;
        mbar_wait(&q_empty[st], ((g / QSTAGES) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + Q_OFF + st * 2 * TILE_BYTES;
          mbar_arrive_expect_tx(&q_full[st], 2 * TILE_BYTES);
          const int qb = (i + kvt) % nq;  // staggered start: the key tiles of a head add to different dQ rows at a time
          tma_load_4d(&tmQ, &q_full[st], dst, 0, qb * QB, h, b, kEvictLast);
          tma_load_4d(&tmdO, &q_full[st], dst + TILE_BYTES, 0, qb * QB, h, b, kEvictLast);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(KV, QB, 0, 0);   // S^T / dP^T: both operands K-major
    constexpr uint32_t idesc_kn = make_idesc_bf16(KV, HD, 0, 1);  // dV / dK: A K-major (TMEM / shared), B MN-major
    constexpr uint32_t idesc_nn = make_idesc_bf16(QB, HD, 1, 1);  // dQ: A MN-major (dS^T read by columns), B MN-major
    int g = 0;
    int it = 0;
    const uint32_t v_addr = smem_u32(smem + V_OFF);
    auto issue_sdp = [&](int gj, uint32_t k_addr, bool last_of_item) {
      const int st = gj % QSTAGES;
      mbar_wait(&q_full[st], (gj / QSTAGES) & 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem + Q_OFF + st * 2 * TILE_BYTES);
      const uint64_t kd = make_sdesc_sw128(k_addr, 1024, 0), vd = make_sdesc_sw128(v_addr, 1024, 0);
      const uint64_t qd = make_sdesc_sw128(q_addr, 1024, 0), od = make_sdesc_sw128(q_addr + TILE_BYTES, 1024, 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + ST_COL, kd + (uint64_t)(k * 2), qd + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + DP_COL, vd + (uint64_t)(k * 2), od + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(sdp_full);
        if (last_of_item) umma_commit(v_empty);  // V is not read again in this item: the next item's V may land
      }
      __syncwarp();
    };
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const int kvb = it & 1;
      mbar_wait(&k_full[kvb], (it >> 1) & 1);
      mbar_wait(v_full, it & 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smem + K_OFF + kvb * TILE_BYTES);
      issue_sdp(g, k_addr, nq == 1);
      for (int i = 0; i < nq; ++i) {
        const int gi = g + i;
        mbar_wait(pds_full, gi & 1);
        tc_fence_after();
        if (i + 1 < nq) issue_sdp(gi + 1, k_addr, i + 2 == nq);
        if (i == 0) {  // the previous item's dV / dK have been read out
          mbar_wait(acc_empty, (it & 1) ^ 1);
          tc_fence_after();
        }
        const int st = gi % QSTAGES;
        const uint32_t q_addr = smem_u32(smem + Q_OFF + st * 2 * TILE_BYTES);
        const uint32_t ds_addr = smem_u32(smem + DS_OFF + (gi & 1) * DS_BYTES);
        const uint64_t q_mn = make_sdesc_sw128(q_addr, 1024, 1024), o_mn = make_sdesc_sw128(q_addr + TILE_BYTES, 1024, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < QB / 16; ++k)  // dV += P^T dO_i
            umma_bf16_ts(tmem_base + DV_COL, tmem_base + P_COL + k * 8, o_mn + (uint64_t)(k * 128), idesc_kn,
                         (i | k) != 0 ? 1u : 0u);
          umma_commit(p_free);
#pragma unroll
          for (int k = 0; k < QB / 16; ++k) {  // dK += dS^T Q_i
            const uint64_t a = make_sdesc_sw128(ds_addr + (k >> 2) * TILE_BYTES, 1024, 0) + (uint64_t)((k & 3) * 2);
            umma_bf16_ss(tmem_base + DK_COL, a, q_mn + (uint64_t)(k * 128), idesc_kn, (i | k) != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        mbar_wait(dq_empty, (gi & 1) ^ 1);  // dQ of the previous block is in the reducers' registers
        tc_fence_after();
        const uint64_t ds_mn = make_sdesc_sw128(ds_addr, 1024, TILE_BYTES), k_mn = make_sdesc_sw128(k_addr, 1024, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KV / 16; ++k)  // dQ_i = dS_i K
            umma_bf16_ss(tmem_base + DQ_COL, ds_mn + (uint64_t)(k * 128), k_mn + (uint64_t)(k * 128), idesc_nn, k != 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(&q_empty[st]);
          if (i == nq - 1) {
            umma_commit(acc_full);
            umma_commit(&k_empty[kvb]);
          }
        }
        __syncwarp();
      }
      g += nq;
    }
  } else if (warp < kSoftmaxWarps) {
    // ===================== P^T / dS^T: thread == key row == TMEM lane; warps 0-3 / 4-7 take 64 query columns each ==========
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int sm_tid = threadIdx.x;  // 0..255
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    int g = 0;
    float pre_nlse = 0.f, pre_nD = 0.f;
    auto load_stats = [&](int w, int i) {  // the first 128 threads fetch the block's per-query statistics (negated)
      if (sm_tid < QB) {
        const int bh = w / p.nkv;
        const int q = ((i + w % p.nkv) % nq) * QB + sm_tid;
        if (q < p.Tq) {
          const long long idx = (long long)bh * p.Tq + q;
          pre_nlse = -__ldg(p.lse + idx), pre_nD = -__ldg(p.D + idx);
        } else {  // a query beyond the sequence: P = 2^(-inf) = 0, dS = 0
          pre_nlse = -INFINITY, pre_nD = 0.f;
        }
      }
    };
    if ((int)blockIdx.x < p.total) load_stats(blockIdx.x, 0);
    for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
      for (int i = 0; i < nq; ++i) {
        const int gi = g + i;
        float* st_nlse = s_stat + (gi & 1) * 2 * QB;
        float* st_nD = st_nlse + QB;
        if (sm_tid < QB) st_nlse[sm_tid] = pre_nlse, st_nD[sm_tid] = pre_nD;
        if (i + 1 < nq) load_stats(w, i + 1);
        else if (w + (int)gridDim.x < p.total) load_stats(w + gridDim.x, 0);
        named_bar_sync(1, kSoftmaxWarps * 32);
        mbar_wait(sdp_full, gi & 1);
        tc_fence_after();
        uint8_t* ds_row = smem + DS_OFF + (gi & 1) * DS_BYTES + half * TILE_BYTES + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 32 query columns at a time
          const int col0 = half * 64 + c * 32;
          uint32_t sr[32], dr[32];
          tmem_ld_x32(lane_addr + ST_COL + col0, sr);
          tmem_ld_x32(lane_addr + DP_COL + col0, dr);
          tmem_ld_wait();
          uint32_t pp[16], ds[16];
          if (p.dbg & 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pp[j] = sr[j], ds[j] = dr[j];
          } else
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 nl = *reinterpret_cast<const float4*>(st_nlse + col0 + j);
            const float4 nd = *reinterpret_cast<const float4*>(st_nD + col0 + j);
            const float2 l2 = make_float2(kLog2e, kLog2e);
            const float2 e0 = fma_f32x2(make_float2(__uint_as_float(sr[j]), __uint_as_float(sr[j + 1])), l2, make_float2(nl.x, nl.y));
            const float2 e1 = fma_f32x2(make_float2(__uint_as_float(sr[j + 2]), __uint_as_float(sr[j + 3])), l2, make_float2(nl.z, nl.w));
            const float2 p0 = make_float2(fast_exp2(e0.x), fast_exp2(e0.y)), p1 = make_float2(fast_exp2(e1.x), fast_exp2(e1.y));
            const float2 t0 = add_f32x2(make_float2(__uint_as_float(dr[j]), __uint_as_float(dr[j + 1])), make_float2(nd.x, nd.y));
            const float2 t1 = add_f32x2(make_float2(__uint_as_float(dr[j + 2]), __uint_as_float(dr[j + 3])), make_float2(nd.z, nd.w));
            const float2 z2 = make_float2(0.f, 0.f);
            const float2 d0 = fma_f32x2(p0, t0, z2), d1 = fma_f32x2(p1, t1, z2);
            pp[j >> 1] = pack_bf16(p0.x, p0.y), pp[(j >> 1) + 1] = pack_bf16(p1.x, p1.y);
            ds[j >> 1] = pack_bf16(d0.x, d0.y), ds[(j >> 1) + 1] = pack_bf16(d1.x, d1.y);
          }
          if (c == 0) {  // the dV GEMM of the previous block has consumed the P^T columns
            mbar_wait(p_free, (gi & 1) ^ 1);
            tc_fence_after();
          }
          tmem_st_x16(lane_addr + P_COL + (col0 >> 1), pp);
          // dS^T row -> shared memory, 128B-swizzled (16-byte chunk index XOR row % 8), 64 queries per atom row
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = ds[4 * j], q.y = ds[4 * j + 1], q.z = ds[4 * j + 2], q.w = ds[4 * j + 3];
            *reinterpret_cast<uint4*>(ds_row + (((c * 4 + j) ^ (row & 7)) << 4)) = q;
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();  // generic-proxy writes of dS^T -> visible to the MMA's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
      }
      g += nq;
    }
  } else {
    // ===================== dQ reduction + dV / dK epilogue (warps 8..11): thread == TMEM lane =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    int g = 0;
    int it = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const int kvt = w % p.nkv, bh = w / p.nkv;
      const int h = bh % p.H, b = bh / p.H;
      for (int i = 0; i < nq; ++i) {
        const int gi = g + i;
        mbar_wait(dq_full, gi & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_x32(lane_addr + DQ_COL, r0);
        tmem_ld_x32(lane_addr + DQ_COL + 32, r1);
        tmem_ld_wait_regs(r0);
        tmem_ld_wait_regs(r1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
        if (!(p.dbg & 1)) {
          // dQ_i -> this warp's staging tile [32 queries x 32 fp32] (128B-swizzled) -> one TMA reduction per 32-column half:
          // the adds happen at L2 on whole lines (per-thread red.global.add.v4 from the accumulator layout -- thread = row --
          // touched 32 different lines per instruction and cost 170 us of this kernel's 510 at B = 8)
          const int q0 = ((i + kvt) % nq) * QB + quad * 32;
          uint8_t* stage = smem + RED_OFF + quad * 4096;
          uint8_t* srow = stage + lane * 128;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (lane == 0) bulk_wait_group_read<0>();  // the previous reduction has read the tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t* r = hh == 0 ? r0 : r1;
              *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(&tmAcc, stage, h * HD + hh * 32, q0, b);
              bulk_commit_group();
            }
          }
        }
      }
      g += nq;
      // ---- dV / dK of the item: accumulators -> bf16 -> global ----
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const int key = kvt * KV + row;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* orow = (which == 0 ? p.dV : p.dK) + (long long)b * p.dkv_bs + (long long)key * p.dkv_rs + h * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          uint32_t o[32];
          tmem_ld_x32(lane_addr + (which == 0 ? DV_COL : DK_COL) + c, o);
          tmem_ld_wait_regs(o);
          if (key < p.Tk) {
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(o[k]), __uint_as_float(o[k + 1]));
              v.y = pack_bf16(__uint_as_float(o[k + 2]), __uint_as_float(o[k + 3]));
              v.z = pack_bf16(__uint_as_float(o[k + 4]), __uint_as_float(o[k + 5]));
              v.w = pack_bf16(__uint_as_float(o[k + 6]), __uint_as_float(o[k + 7]));
              *reinterpret_cast<uint4*>(orow + c + k) = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (lane == 0) bulk_wait_group<0>();  // the staging tiles must outlive the last reductions
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dq[b, t, c] (bf16, strided) = acc[b, t, c] (fp32 [B * Tq, C] contiguous); 8 columns per thread
__global__ void __launch_bounds__(256) dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int Tq,
                                                         int C, long long total8, long long rs, long long bs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total8) return;
  const int c8 = C / 8;
  const long long r = gid / c8;
  const int c = (int)(gid % c8) * 8;
  const int t = (int)(r % Tq);
  const long long b = r / Tq;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(acc + r * C + c));
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(acc + r * C + c) + 1);
  uint4 v;
  v.x = pack_bf16(a0.x, a0.y), v.y = pack_bf16(a0.z, a0.w), v.z = pack_bf16(a1.x, a1.y), v.w = pack_bf16(a1.z, a1.w);
  *reinterpret_cast<uint4*>(dq + b * bs + (long long)t * rs + c) = v;
}

}  // namespace

int launch_attention_bwd_fused(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmdO, const CUtensorMap& tmK,
                               const CUtensorMap& tmV, const FusedBwdArgs& a, cudaStream_t stream) {
  static DeviceOnce attr_once;
  if (attr_once.first(ctx))
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(attention_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  FusedParams p{};
  p.B = a.B, p.H = a.H, p.Tq = a.Tq, p.Tk = a.Tk;
  p.nkv = ceil_div(a.Tk, KV), p.nq = ceil_div(a.Tq, QB);
  p.total = a.B * a.H * p.nkv;
  p.lse = a.lse, p.D = a.D, p.dq_acc = a.dq_acc;
  p.dK = a.dK, p.dV = a.dV, p.dkv_rs = a.dkv_rs, p.dkv_bs = a.dkv_bs;
  static const int dbg = [] {
    const char* e = getenv("DICOW_BWD_FUSED_DBG");
    return e != nullptr ? atoi(e) : 0;
  }();
  p.dbg = dbg;
  const long long C = (long long)a.H * HD;
  DICOW_CUDA_OK(ctx, cudaMemsetAsync(a.dq_acc, 0, sizeof(float) * (size_t)a.B * a.Tq * C, stream));
  const int grid = p.total < ctx->num_sms ? p.total : ctx->num_sms;
  CUtensorMap tmAcc;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)a.Tq, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)a.Tq * C * 4};
    uint32_t box[3] = {32, 32, 1};
    int rc = make_tmap_f32(ctx, &tmAcc, a.dq_acc, 3, dims, strides, box);
    if (rc) return rc;
  }
  attention_bwd_fused_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(tmQ, tmdO, tmK, tmV, tmAcc, p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  const long long total8 = (long long)a.B * a.Tq * C / 8;
  dq_convert_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, stream>>>(a.dq_acc, a.dQ, a.Tq, (int)C, total8, a.dq_rs, a.dq_bs);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

}  // namespace dicow
