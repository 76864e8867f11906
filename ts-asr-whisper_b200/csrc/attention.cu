// attention.cu -- flash-style multi-head attention (head_dim 64) on tcgen05 tensor cores (sm_100a).
//
//   O[b, t, h, :] = softmax_k( Q[b, t, h, :] . K[b, k, h, :] ) V[b, k, h, :]        (q already carries hd^-0.5)
//
// One CTA = one 128-row query tile of one (batch, head); 192 threads; 2 CTAs co-reside per SM so that one CTA's
// softmax overlaps the other's MMAs (the kernel is MUFU(exp2)-bound at head_dim 64: 128x128 exps per KV block take
// ~1024 clk on the 16/clk SFU while the two MMAs of the block take 512 clk of tensor pipe).
//   warps 0..3 : softmax  -- thread == query row (TMEM lane); S row via tcgen05.ld, running max / sum in registers,
//                P = exp2(s*log2e - m*log2e) packed to bf16 and handed back to the tensor core either through TMEM
//                (tcgen05.st, A-operand-from-TMEM MMA) or through 128B-swizzled shared memory; lazy O rescale
//                (only when a row max grows by more than 2^8) done in TMEM by the same threads
//   warp 4     : TMA producer -- Q once; K_j, V_j tiles [128 x 64] bf16 through a 2-stage mbarrier ring
//   warp 5     : MMA issuer + TMEM allocator -- S = Q K_j^T (M128 N128 K64), O += P V_j (M128 N64 K128,
//                V consumed MN-major straight from its row-major [key, hd] tile)
// In-order completion of tcgen05.mma + tcgen05.commit gives all cross-role ordering: s_full(j+1) implies PV(j) retired.
//
// Replaces the SDPA call inside HF WhisperAttention (HF:modeling_whisper.py:342-352) for encoder self-attention
// (src/models/dicow/encoder.py:216-221), the SE-DiCoW enrollment cross-attention (src/models/dicow/layers.py:156-160)
// and the decoder's teacher-forced attention.
#include <math.h>

#include "attention_common.h"
#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per block
constexpr int KV_STAGES = 2;
constexpr int kAttnThreads = 192;
constexpr uint32_t TILE_BYTES = BQ * HD * 2;  // 16 KB: Q, K_j and V_j tiles are all [128 x 64] bf16

// TMEM columns
constexpr uint32_t S_COL = 0;    // 128 fp32 columns
constexpr uint32_t O_COL = 128;  // 64 fp32 columns
constexpr uint32_t P_COL = 192;  // 64 columns = 128 bf16 (P through TMEM)
constexpr uint32_t TMEM_COLS = 256;

template <bool P_SMEM>
struct AttnSmem {
  static constexpr uint32_t Q_OFF = 0;
  static constexpr uint32_t K_OFF = TILE_BYTES;
  static constexpr uint32_t V_OFF = K_OFF + KV_STAGES * TILE_BYTES;
  static constexpr uint32_t P_OFF = V_OFF + KV_STAGES * TILE_BYTES;
  static constexpr uint32_t BAR_OFF = P_OFF + (P_SMEM ? 2 * TILE_BYTES : 0);
  static constexpr uint32_t BYTES = BAR_OFF + 128 + 1024 /*alignment slack*/;
};

constexpr float kLog2e = 1.4426950408889634f;


template <bool P_SMEM, int EMU, bool TWO_PASS>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = AttnSmem<P_SMEM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* k_full = q_full + 1;               // [KV_STAGES]
  uint64_t* v_full = k_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* kv_empty = v_full + KV_STAGES;     // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;     // S_j ready in TMEM (and PV_{j-1} retired)
  uint64_t* p_full = s_full + 1;               // P_j written, S_j drained
  uint64_t* o_full = p_full + 1;               // last PV retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int off = p.Tk - p.Tq;  // causal: key k visible to query t iff k <= t + off
  int kv_end = p.Tk;
  if (p.causal) kv_end = min(p.Tk, q0 + BQ + off);
  if (kv_end < 1) kv_end = 1;
  const int nkv = (kv_end + BKV - 1) / BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);  // one arrive per softmax warp
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(&tmQ, q_full, smem + L::Q_OFF, 0, q0, h, b, kEvictFirst);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(&tmK, &k_full[st], smem + L::K_OFF + st * TILE_BYTES, 0, j * BKV, h, b, kEvictLast);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_4d(&tmV, &v_full[st], smem + L::V_OFF + st * TILE_BYTES, 0, j * BKV, h, b, kEvictLast);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);   // P K-major, V MN-major ([key, hd] rows)
      const uint32_t q_addr = smem_u32(smem + L::Q_OFF);
      const uint32_t d_s = tmem_base + S_COL;
      const uint32_t d_o = tmem_base + O_COL;
      auto issue_s = [&](int j) {
        const int st = j % KV_STAGES;
        mbar_wait(&k_full[st], (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + L::K_OFF + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(d_s, make_sdesc_sw128(q_addr + k * 32, 1024, 0), make_sdesc_sw128(k_addr + k * 32, 1024, 0),
                       idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      const bool prof_mma = p.prof != nullptr && blockIdx.x == 3 && blockIdx.y == 1 && blockIdx.z == 1;
      for (int j = 0; j < nkv; ++j) {
        const int st = j % KV_STAGES;
        mbar_wait(p_full, j & 1);
        if (prof_mma) p.prof[j * 8 + 6] = clock64();
        mbar_wait(&v_full[st], (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + L::V_OFF + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t db = make_sdesc_sw128(v_addr + k * 2048, 1024, 1024);  // 16 key rows per step
          const uint32_t accum = (j | k) != 0 ? 1u : 0u;
          if constexpr (P_SMEM) {
            const uint32_t p_addr = smem_u32(smem + L::P_OFF) + (k >> 2) * TILE_BYTES + (k & 3) * 32;
            umma_bf16_ss(d_o, make_sdesc_sw128(p_addr, 1024, 0), db, idesc_o, accum);
          } else {
            umma_bf16_ts(d_o, tmem_base + P_COL + k * 8, db, idesc_o, accum);
          }
        }
        umma_commit(&kv_empty[st]);
        if (prof_mma) p.prof[j * 8 + 7] = clock64();
        if (j + 1 < nkv)
          issue_s(j + 1);
        else
          umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 0..3) =====================
    const int row = warp * 32 + lane;  // TMEM lane == query row within the tile
    const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
    const int q_idx = q0 + row;
    float m_used = -INFINITY;  // max used as the exponent offset (natural units)
    float l = 0.f;
    const bool prof_on = p.prof != nullptr && blockIdx.x == 3 && blockIdx.y == 1 && blockIdx.z == 1 && row == 0;
#define ATTN_STAMP(slot) do { if (prof_on) p.prof[j * 8 + (slot)] = clock64(); } while (0)
    for (int j = 0; j < nkv; ++j) {
      ATTN_STAMP(0);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      ATTN_STAMP(1);
      // key-tail / causal masking (block-uniform test, per-row limit)
      const int k0 = j * BKV;
      const bool need_mask = (k0 + BKV > p.Tk) || (p.causal && (k0 + BKV - 1 > q0 + off));
      int limit = BKV;
      if (need_mask) {
        limit = p.Tk;
        if (p.causal) limit = min(limit, q_idx + off + 1);
        limit = max(limit, 1) - k0;  // columns c < limit are visible (key 0 always is: no all -inf row in block 0)
      }
      // ---- pass 1: row max.  TWO_PASS re-reads S from TMEM in pass 2 (32 live values at a time) instead of
      //      holding the whole 128-value row in registers (which spills at 168 regs/thread) ----
      float s[TWO_PASS ? 1 : BKV];
      float m8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m8[i] = -INFINITY;
      if constexpr (TWO_PASS) {
        uint32_t r[2][32];
        tmem_ld_x32(lane_addr + S_COL, r[0]);
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          tmem_ld_wait();
          if (c + 1 < BKV / 32) tmem_ld_x32(lane_addr + S_COL + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(r[c & 1][i]);
            if (need_mask && c * 32 + i >= limit) v = -INFINITY;
            m8[i & 7] = fmaxf(m8[i & 7], v);
          }
        }
      } else {
        uint32_t r[BKV / 32][32];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) tmem_ld_x32(lane_addr + S_COL + c * 32, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < BKV; ++c) {
          float v = __uint_as_float(r[c / 32][c % 32]);
          if (need_mask && c >= limit) v = -INFINITY;
          s[c] = v;
          m8[c & 7] = fmaxf(m8[c & 7], v);
        }
      }
      const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                             fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      ATTN_STAMP(2);
      // lazy rescale: keep the old offset unless the row max grew by more than 2^8
      const bool grow = (mx - m_used) * kLog2e > 8.0f;  // true on the first block (m_used = -inf), false for mx = -inf
      float scale = 1.0f;
      if (grow) {
        scale = fast_exp2((m_used - mx) * kLog2e);  // 0 on the first block
        m_used = mx;
        l *= scale;
      }
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // O rows live in TMEM; PV_{j-1} has retired (s_full(j) was committed after it) and PV_j is not issued yet
        uint32_t o[32];
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          tmem_ld_x32(lane_addr + O_COL + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
          tmem_st_x32(lane_addr + O_COL + c, o);
        }
        tmem_st_wait();
      }
      ATTN_STAMP(3);
      // ---- pass 2: p = 2^(s*log2e - m*log2e), bf16 P to the tensor core, row sum.  EMU of every 8 exponentials
      //      run as a polynomial on the FMA pipe instead of the SFU ----
      const float moff = m_used * kLog2e;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t r2[2][32];
      if constexpr (TWO_PASS) tmem_ld_x32(lane_addr + S_COL, r2[0]);
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) {
        if constexpr (TWO_PASS) {
          tmem_ld_wait();
          if (c + 1 < BKV / 32) tmem_ld_x32(lane_addr + S_COL + (c + 1) * 32, r2[(c + 1) & 1]);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float v;
            if constexpr (TWO_PASS) {
              v = __uint_as_float(r2[c & 1][i + u]);
              if (need_mask && c * 32 + i + u >= limit) v = -INFINITY;
            } else {
              v = s[c * 32 + i + u];
            }
            const float x = fmaf(v, kLog2e, -moff);
            e[u] = (((i + u) & 7) < EMU) ? poly_exp2(x) : fast_exp2(x);
          }
          ls[(i >> 1) & 1] += e[0];
          ls[2 + ((i >> 1) & 1)] += e[1];
          pk[i >> 1] = pack_bf16(e[0], e[1]);
        }
        if constexpr (P_SMEM) {
          uint8_t* prow = smem + L::P_OFF + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {  // 16-byte chunks of 8 keys, 128B-swizzled K-major rows
            const int q = c * 4 + q4;
            *reinterpret_cast<uint4*>(prow + (q >> 3) * TILE_BYTES + (((q & 7) ^ (row & 7)) << 4)) =
                make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
          }
        } else {
          tmem_st_x16(lane_addr + P_COL + c * 16, pk);
        }
      }
      if constexpr (P_SMEM)
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the MMA (async proxy)
      else
        tmem_st_wait();
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      ATTN_STAMP(4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ATTN_STAMP(5);
    }
#undef ATTN_STAMP
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint32_t o[32];
    __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)q_idx * p.o_rs + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      tmem_ld_x32(lane_addr + O_COL + c, o);
      tmem_ld_wait();
      if (q_idx < p.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + i) = v;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int make_qkv_map(dicow_ctx* ctx, CUtensorMap* m, const void* base, int T, int H, int B, long long row_stride,
                 long long batch_stride) {
  uint64_t dims[4] = {(uint64_t)HD, (uint64_t)T, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)row_stride * 2, (uint64_t)HD * 2, (uint64_t)batch_stride * 2};
  uint32_t box[4] = {HD, BQ, 1, 1};
  return make_tmap_bf16(ctx, m, base, 4, dims, strides, box);
}

template <bool P_SMEM, int EMU, bool TWO_PASS>
int launch_attention(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                     const AttnParams& p, cudaStream_t stream) {
  using L = AttnSmem<P_SMEM>;
  auto kfn = attention_kernel<P_SMEM, EMU, TWO_PASS>;
  static DeviceOnce attr_once;
  if (attr_once.first(ctx)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
  }
  dim3 grid(ceil_div(p.Tq, BQ), p.H, p.B);
  kfn<<<grid, kAttnThreads, L::BYTES, stream>>>(tmQ, tmK, tmV, p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_attention_bf16(dicow_handle_t h, const dicow_attention_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_attention_args_t),
                "dicow_attention_bf16: bad args struct");
  DICOW_REQUIRE(ctx, a->Q && a->K && a->V && a->out, "dicow_attention_bf16: null operand");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->H >= 1 && a->Tq >= 1 && a->Tk >= 1 && a->B <= 65535 && a->H <= 65535,
                "dicow_attention_bf16: bad shape B=%d H=%d Tq=%d Tk=%d", a->B, a->H, a->Tq, a->Tk);
  DICOW_REQUIRE(ctx, !a->causal || a->Tk >= a->Tq, "dicow_attention_bf16: causal needs Tk >= Tq");
  const long long strides[6] = {a->q_row_stride,  a->q_batch_stride, a->kv_row_stride,
                                a->kv_batch_stride, a->o_row_stride,   a->o_batch_stride};
  for (int i = 0; i < 6; ++i)
    DICOW_REQUIRE(ctx, strides[i] >= 0 && (strides[i] % 8) == 0, "dicow_attention_bf16: stride %d not a multiple of 8", i);
  DICOW_REQUIRE(ctx, a->q_row_stride >= (long long)a->H * HD && a->kv_row_stride >= (long long)a->H * HD &&
                         a->o_row_stride >= (long long)a->H * HD,
                "dicow_attention_bf16: row strides must cover H*64 elements");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->Q) | reinterpret_cast<uintptr_t>(a->K) |
                       reinterpret_cast<uintptr_t>(a->V) | reinterpret_cast<uintptr_t>(a->out);
  DICOW_REQUIRE(ctx, (al % 16) == 0, "dicow_attention_bf16: operands must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  // a batch stride of 0 is legal for B == 1 only; TMA wants a non-zero stride
  const long long qbs = a->B > 1 ? a->q_batch_stride : (long long)a->Tq * a->q_row_stride;
  const long long kbs = a->B > 1 ? a->kv_batch_stride : (long long)a->Tk * a->kv_row_stride;
  int rc = make_qkv_map(ctx, &tmQ, a->Q, a->Tq, a->H, a->B, a->q_row_stride, qbs);
  if (rc) return rc;
  rc = make_qkv_map(ctx, &tmK, a->K, a->Tk, a->H, a->B, a->kv_row_stride, kbs);
  if (rc) return rc;
  rc = make_qkv_map(ctx, &tmV, a->V, a->Tk, a->H, a->B, a->kv_row_stride, kbs);
  if (rc) return rc;
  AttnParams p{};
  p.B = a->B, p.H = a->H, p.Tq = a->Tq, p.Tk = a->Tk, p.causal = a->causal ? 1 : 0;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.o_rs = a->o_row_stride, p.o_bs = a->o_batch_stride;
  p.prof = reinterpret_cast<long long*>(ctx->attn_prof);
  p.lse = a->lse;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // variant 0..3: ping-pong kernel (attention_fa.cu) with {0, 2, 4, 6} of every 8 exponentials on the FMA pipe; 4: packed bf16.
  // variant 16 + v: the single-tile kernel of this file, kept for comparison -- v bits: [0] P through shared memory;
  // [1..2] poly-exp2 share: 0 -> 2/8, 1 -> 0/8, 2 -> 3/8, 3 -> 4/8; [3] single pass (hold the S row in registers)
  DICOW_REQUIRE(ctx, a->lse == nullptr || (a->variant >= 0 && a->variant <= 4),
                "dicow_attention_bf16: the lse output needs the default kernel");
  if (a->variant >= 0 && a->variant <= 4) return launch_attention_fa(ctx, tmQ, tmK, tmV, p, a->variant, stream);
  switch (a->variant - 16) {
    case 0: return launch_attention<false, 2, true>(ctx, tmQ, tmK, tmV, p, stream);
    case 1: return launch_attention<true, 2, true>(ctx, tmQ, tmK, tmV, p, stream);
    case 8: return launch_attention<false, 2, false>(ctx, tmQ, tmK, tmV, p, stream);
    case 10: return launch_attention<false, 0, false>(ctx, tmQ, tmK, tmV, p, stream);
    default: return set_error(ctx, DICOW_ERR_INVALID_ARG, "dicow_attention_bf16: unknown variant %d", a->variant);
  }
}

// Debug aid (not part of the data path): per-KV-step clock64 stamps of CTA (3,1,1) are written to `buf`
// ([steps][8] int64: 0 loop top, 1 S ready, 2 max done, 3 O rescaled, 4 P written, 5 arrived, 6 MMA saw P, 7 MMA issued PV).
extern "C" int dicow_debug_set_attention_profile(dicow_handle_t h, void* buf) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  h->attn_prof = buf;
  return DICOW_OK;
}
