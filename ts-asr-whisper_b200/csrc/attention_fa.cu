// attention_fa.cu -- flash attention (head_dim 64) on tcgen05, two query tiles per CTA in ping-pong (sm_100a).
//
//   O[b, t, h, :] = softmax_k( Q[b, t, h, :] . K[b, k, h, :] ) V[b, k, h, :]        (q already carries hd^-0.5)
//
// At head_dim 64 the kernel is bound by the exponentials (128 x 128 per KV block per tile on the 16-lane/clk SFU =
// 1024 clk) rather than by the two MMAs of the block (512 clk of tensor pipe), so the design goal is to keep the
// SFU busy: one CTA per SM owns TWO 128-row query tiles of one (batch, head); each tile has its own softmax warp
// group, and the single MMA-issuing thread alternates between the tiles so that the tensor-core work of one tile
// (P V and the next Q K^T) runs under the other tile's softmax.  K/V tiles are loaded once for both query tiles.
//
//   warps 0..3 : softmax of tile 0 -- thread == query row (TMEM lane); the S row is pulled into registers with
//                tcgen05.ld (single pass), running max / sum kept per thread, P = exp2(s*log2e - m*log2e) packed to
//                bf16 and written over the S columns in TMEM (tcgen05.st) as the A operand of the P V MMA; lazy O
//                rescale (only when a row max grows by more than 2^8) done in TMEM by the same threads
//   warps 4..7 : softmax of tile 1 (same TMEM lanes, different columns)
//   warp 8     : TMA producer -- Q tiles once; K_j, V_j tiles [128 x 64] bf16 through a KV_STAGES-deep mbarrier ring
//   warp 9     : TMEM allocator + issuer of S_t = Q_t K_j^T (M128 N128 K64)
//   warp 10    : issuer of O_t += P_t V_j (M128 N64 K128, A from TMEM, V consumed MN-major straight from its
//                row-major [key, hd] tile); separate from warp 9 because preparing a batch costs more issue latency
//                than the batch's tensor-pipe time -- the two streams are ordered through mbarriers only
//   warp 11    : idle; completes the third warp group so that it can hand its registers (setmaxnreg.dec 56) to
//                the softmax groups (setmaxnreg.inc 224: the 128-value S row is held in registers)
// TMEM columns: three S/P buffers [0,384) shared round-robin by the two tiles, O0 [384,448), O1 [448,512).  S of
// production i+3 is issued as soon as the P V of production i has retired (p_consumed[buffer]), 1.5 blocks ahead of use, so
// every tile's next S is ready long before its softmax group asks for it -- the softmax groups never wait on the
// tensor core in steady state.  P (bf16) aliases the first 64 columns of its S buffer.
//
// Replaces the SDPA call inside HF WhisperAttention (HF:modeling_whisper.py:342-352) for encoder self-attention
// (src/models/dicow/encoder.py:216-221), the SE-DiCoW enrollment cross-attention (src/models/dicow/layers.py:156-160)
// and the decoder's teacher-forced self/cross attention (HF:modeling_whisper.py:449-506).
#include <math.h>

#include <type_traits>

#include "attention_common.h"
#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr int HD = 64;
constexpr int BQ = 128;   // query rows per tile
constexpr int BKV = 128;  // keys per block
constexpr int KV_STAGES = 4;
constexpr int kThreads = 384;  // 3 warp groups: softmax tile 0, softmax tile 1, {TMA, MMA, 2 idle}
constexpr uint32_t TILE_BYTES = BQ * HD * 2;  // 16 KB: Q, K_j and V_j tiles are all [128 x 64] bf16

constexpr int S_BUFS = 3;         // S/P buffers shared round-robin by the two tiles
constexpr uint32_t S_COL = 0;     // + buf * 128
constexpr uint32_t O_COL = 384;   // + t * 64
constexpr uint32_t TMEM_COLS = 512;

constexpr uint32_t Q_OFF = 0;               // 2 slots x 2 tiles (the next item's Q loads while the current one runs)
constexpr uint32_t K_OFF = 4 * TILE_BYTES;
constexpr uint32_t V_OFF = K_OFF + KV_STAGES * TILE_BYTES;
constexpr uint32_t BAR_OFF = V_OFF + KV_STAGES * TILE_BYTES;
constexpr uint32_t SMEM_BYTES = BAR_OFF + 256 + 1024 /*alignment slack*/;

constexpr float kLog2e = 1.4426950408889634f;

// The (tile, block) pairs are produced in the order (0,0) (1,0) (0,1) (1,1) ... while both tiles have blocks left
// (n1 >= n0 whenever tile 1 exists: it holds the later queries), then the rest of tile 1.  Production #i uses S/P
// buffer i % 3.
__device__ __forceinline__ int prod_index(int t, int j, int n0, int n1) {
  if (n1 == 0) return j;
  return j < n0 ? 2 * j + t : n0 + j;
}
__device__ __forceinline__ void prod_decode(int i, int n0, int n1, int& t, int& j) {
  if (n1 == 0) {
    t = 0, j = i;
  } else if (i < 2 * n0) {
    t = i & 1, j = i >> 1;
  } else {
    t = 1, j = i - n0;
  }
}

// Work item w of a launch = (query tile pair qt, head h, batch b), qt fastest so that CTAs running at the same time share
// K/V through L2.  The kernel is PERSISTENT: CTA c runs items c, c + gridDim.x, ... with every mbarrier ring, S/P buffer
// rotation and phase carried across items, so the producer / S-issuer run ahead into the next item (Q double-buffered)
// while the softmax groups finish the current one.  (One CTA per item paid ~4100 clk of exposed prologue -- TMEM
// allocation, first TMA round trip, first S -- out of ~29 000 clk per item on the in-kernel timeline.)
struct Item {
  int q0, h, b, n0, n1;
};
__device__ __forceinline__ Item decode_item(int w, int nqt, const AttnParams& p) {
  Item it;
  const int qt = w % nqt;
  const int bh = w / nqt;
  it.q0 = qt * (2 * BQ);
  it.h = bh % p.H;
  it.b = bh / p.H;
  const int off = p.Tk - p.Tq;  // causal: key k visible to query t iff k <= t + off
  int kv_end = p.causal ? min(p.Tk, it.q0 + BQ + off) : p.Tk;
  it.n0 = (max(kv_end, 1) + BKV - 1) / BKV;
  kv_end = p.causal ? min(p.Tk, it.q0 + 2 * BQ + off) : p.Tk;
  it.n1 = (it.q0 + BQ < p.Tq) ? (max(kv_end, 1) + BKV - 1) / BKV : 0;  // tile 1 may be absent
  return it;
}

template <int EMU>
__global__ void __launch_bounds__(kThreads, 1)
attention_fa_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [2] Q slot loaded
  uint64_t* q_empty = q_full + 2;           // [2] every S MMA reading the slot has retired
  uint64_t* k_full = q_empty + 2;           // [KV_STAGES]
  uint64_t* v_full = k_full + KV_STAGES;    // [KV_STAGES]
  uint64_t* kv_empty = v_full + KV_STAGES;  // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;  // [S_BUFS] S of production i ready in buffer i % 3
  uint64_t* p_full = s_full + S_BUFS;       // [S_BUFS] P written over it (S drained)
  uint64_t* pv_done = p_full + S_BUFS;      // [2] P_t(j) V_j retired (O_t up to date)
  uint64_t* p_consumed = pv_done + 2;       // [S_BUFS] the P V MMAs that read P out of buffer b have retired
  uint64_t* o_free = p_consumed + S_BUFS;   // [2] the softmax group has read O_t of the finished item out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int nqt = (p.Tq + 2 * BQ - 1) / (2 * BQ);
  const int off = p.Tk - p.Tq;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&pv_done[s], 1);
      mbar_init(&o_free[s], 4);  // one arrive per softmax warp of the tile
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < S_BUFS; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);  // one arrive per softmax warp of the tile
      mbar_init(&p_consumed[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp >= 8) {
    reg_dealloc<56>();
    if (warp == 8) {
      // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
      int g = 0;  // K/V blocks loaded so far (all items): stage g % KV_STAGES
      int it_idx = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it_idx) {
        const Item it = decode_item(w, nqt, p);
        const int nmax = max(it.n0, it.n1);
        const int qs = it_idx & 1;
        mbar_wait(&q_empty[qs], ((it_idx >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qs], TILE_BYTES * (it.n1 > 0 ? 2 : 1));
          tma_load_4d(&tmQ, &q_full[qs], smem + Q_OFF + (2 * qs) * TILE_BYTES, 0, it.q0, it.h, it.b, kEvictFirst);
          if (it.n1 > 0)
            tma_load_4d(&tmQ, &q_full[qs], smem + Q_OFF + (2 * qs + 1) * TILE_BYTES, 0, it.q0 + BQ, it.h, it.b, kEvictFirst);
        }
        __syncwarp();
        for (int j = 0; j < nmax; ++j, ++g) {
          const int st = g & (KV_STAGES - 1);
          mbar_wait(&kv_empty[st], ((g / KV_STAGES) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
            tma_load_4d(&tmK, &k_full[st], smem + K_OFF + st * TILE_BYTES, 0, j * BKV, it.h, it.b, kEvictLast);
            mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
            tma_load_4d(&tmV, &v_full[st], smem + V_OFF + st * TILE_BYTES, 0, j * BKV, it.h, it.b, kEvictLast);
          }
          __syncwarp();
        }
      }
    } else if (warp == 9) {
      // ===================== S = Q K^T issuer (whole warp runs the loop; one elected lane issues) =====================
      // Two issuing warps: on the in-kernel timeline of round 1 a single warp needed ~500 clk of (dependent, mostly
      // uniform-datapath) instructions to prepare EACH batch -- S (256 clk of tensor-pipe work) and P V (384 clk) alike
      // -- so the pipe idled half the time and the softmax groups waited ~650 clk per block for their next S.  The S and
      // P V streams are independent instruction streams ordered only through mbarriers, so each gets its own warp.
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      int buf = 0;
      uint32_t free_par = 1;  // parity of the p_consumed[buf] completion this production must see (none in round 0)
      int g0 = 0;             // K/V block counter at the item's block 0
      int it_idx = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it_idx) {
        const Item it = decode_item(w, nqt, p);
        const int total = it.n0 + it.n1;
        const int qs = it_idx & 1;
        mbar_wait(&q_full[qs], (it_idx >> 1) & 1);
        for (int i = 0; i < total; ++i) {
          int t, j;
          prod_decode(i, it.n0, it.n1, t, j);
          const int g = g0 + j;
          const int st = g & (KV_STAGES - 1);
          mbar_wait(&p_consumed[buf], free_par);  // the P V that read P out of this buffer 3 productions ago has retired
          mbar_wait(&k_full[st], (g / KV_STAGES) & 1);
          tc_fence_after();
          const uint64_t qdesc = make_sdesc_sw128(smem_u32(smem + Q_OFF + (2 * qs + t) * TILE_BYTES), 1024, 0);
          const uint64_t kdesc = make_sdesc_sw128(smem_u32(smem + K_OFF + st * TILE_BYTES), 1024, 0);
          const uint32_t d_s = tmem_base + S_COL + buf * BKV;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)  // +32 B along K per step: +2 in the descriptor's (addr >> 4) field
              umma_bf16_ss(d_s, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_full[buf]);
            if (i == total - 1) umma_commit(&q_empty[qs]);  // every S of the item has read its Q tiles
          }
          __syncwarp();
          if (buf == S_BUFS - 1) {
            buf = 0;
            free_par ^= 1;
          } else {
            ++buf;
          }
        }
        g0 += max(it.n0, it.n1);
      }
    } else if (warp == 10) {
      // ===================== O += P V issuer =====================
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);   // P K-major (TMEM), V MN-major ([key, hd] rows)
      int buf = 0;
      uint32_t p_par = 0;
      int g0 = 0;
      uint32_t o_par[2] = {1, 1};  // o_free[t] parity to see before the first P V of an item (nothing to wait for at first)
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const Item it = decode_item(w, nqt, p);
        const int total = it.n0 + it.n1;
        for (int i = 0; i < total; ++i) {
          int t, j;
          prod_decode(i, it.n0, it.n1, t, j);
          const int g = g0 + j;
          const int st = g & (KV_STAGES - 1);
          mbar_wait(&p_full[buf], p_par);
          mbar_wait(&v_full[st], (g / KV_STAGES) & 1);
          if (j == 0) {  // the accumulator of the previous item must have been read out by its softmax group
            mbar_wait(&o_free[t], o_par[t]);
            o_par[t] ^= 1;
          }
          tc_fence_after();
          const uint64_t vdesc = make_sdesc_sw128(smem_u32(smem + V_OFF + st * TILE_BYTES), 1024, 1024);
          const uint32_t d_o = tmem_base + O_COL + t * HD;
          const uint32_t a_p = tmem_base + S_COL + buf * BKV;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)  // 16 keys per step: 8 TMEM columns of P, 16 rows (2048 B) of V
              umma_bf16_ts(d_o, a_p + k * 8, vdesc + (uint64_t)(k * 128), idesc_o, (j | k) != 0 ? 1u : 0u);
            umma_commit(&pv_done[t]);
            umma_commit(&p_consumed[buf]);
            // K_j / V_j are free once the last tile that uses them has issued its PV (tile 1 whenever it exists)
            if (it.n1 == 0 || t == 1) umma_commit(&kv_empty[st]);
          }
          __syncwarp();
          if (buf == S_BUFS - 1) {
            buf = 0;
            p_par ^= 1;
          } else {
            ++buf;
          }
        }
        g0 += max(it.n0, it.n1);
      }
    }
  } else {
    reg_alloc<224>();
    // ===================== softmax / correction / epilogue (warps 0..7) =====================
    const int t = warp >> 2;  // query tile of this warp group
    const int row = (warp & 3) * 32 + lane;  // TMEM lane == query row within the tile
    const uint32_t lane_addr = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t o_addr = lane_addr + O_COL + t * HD;
    int prod_base = 0;  // productions of all earlier items, mod 3 (buffer rotation) ...
    uint32_t round_base = 0;  // ... and the parity of (earlier productions / 3)
    uint32_t pv_base = 0;     // P V batches of this tile in earlier items
    bool first_item = true;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item it = decode_item(w, nqt, p);
      const int n0 = it.n0, n1 = it.n1, q0 = it.q0, h = it.h, b = it.b;
      const int nkv = t == 0 ? n0 : n1;
      if (nkv > 0) {
      const int q_idx = q0 + t * BQ + row;
      float m_used = -INFINITY;  // max used as the exponent offset (natural units)
      float l = 0.f;
      const bool prof_on = p.prof != nullptr && first_item && blockIdx.x == 0 && row == 0;
#define FA_STAMP(slot) do { if (prof_on) p.prof[(t * 16 + j) * 8 + (slot)] = clock64(); } while (0)
      // one KV block; MASK is compiled in only for the key-tail / causal-diagonal blocks
      auto block = [&](auto mask_tag, const int j) {
        constexpr bool MASK = decltype(mask_tag)::value;
        const int i = prod_base + prod_index(t, j, n0, n1);
        const int buf = i % S_BUFS;
        const uint32_t s_par = (round_base + (uint32_t)(i / S_BUFS)) & 1;
        const uint32_t s_addr = lane_addr + S_COL + buf * BKV;
        FA_STAMP(0);
        mbar_wait(&s_full[buf], s_par);
        tc_fence_after();
        FA_STAMP(1);
        float s[BKV];
        float mx;
        {
          // the S row arrives in two halves: the running max of the first half is computed under the second half's
          // tcgen05.ld round trip
          uint32_t r[BKV / 32][32];
          tmem_ld_x32(s_addr, r[0]);
          tmem_ld_x32(s_addr + 32, r[1]);
          tmem_ld_wait_regs(r[0]);
          tmem_ld_wait_regs(r[1]);
          tmem_ld_x32(s_addr + 64, r[2]);
          tmem_ld_x32(s_addr + 96, r[3]);
          FA_STAMP(2);
          int limit = BKV;
          if constexpr (MASK) {
            limit = p.Tk;
            if (p.causal) limit = min(limit, q_idx + off + 1);
            limit = max(limit, 1) - j * BKV;  // columns c < limit are visible (key 0 always is: no all -inf row in block 0)
          }
          float m8[8];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 1) {
              tmem_ld_wait_regs(r[2]);
              tmem_ld_wait_regs(r[3]);
            }
#pragma unroll
            for (int c = half * 64; c < half * 64 + 64; ++c) {
              float v = __uint_as_float(r[c / 32][c % 32]);
              if constexpr (MASK) {
                if (c >= limit) v = -INFINITY;
              }
              s[c] = v;
              m8[c & 7] = c < 8 ? v : fmaxf(m8[c & 7], v);
            }
          }
          mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
        }
        // lazy rescale: keep the old offset unless the row max grew by more than 2^kGrow.  With the packed bf16 exponential
        // the exponent itself is rounded to bf16 (relative 2^-8), so the offset is kept within one octave of the row max:
        // the dominant probabilities then have |x| <= 1 and carry no more error than their own bf16 rounding.
        constexpr float kGrow = EMU < 0 ? 1.0f : 8.0f;
        const bool grow = (mx - m_used) * kLog2e > kGrow;  // true on the first block (m_used = -inf); false for mx = -inf
        float scale = 1.0f;
        if (grow) {
          scale = fast_exp2((m_used - mx) * kLog2e);  // 0 on the first block
          m_used = mx;
          l *= scale;
        }
        FA_STAMP(3);
        const float2 noff2 = make_float2(-m_used * kLog2e, -m_used * kLog2e);
        const float2 l2e2 = make_float2(kLog2e, kLog2e);
        float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if constexpr (EMU < 0) {
          // p = 2^(s*log2e - m*log2e) with TWO exponentials per SFU operation (ex2.approx.ftz.bf16x2): the exponent pair is
          // rounded to bf16, the packed result is the bf16 P operand itself.  Row sum: three levels of packed bf16 adds
          // (unbiased 2^-9 rounding per level), then fp32 -- the sum sees the same rounded probabilities as the P V MMA.
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 x = fma_f32x2(make_float2(s[c * 32 + k], s[c * 32 + k + 1]), l2e2, noff2);
              pk[k >> 1] = ex2_bf16x2(pack_bf16(x.x, x.y));
            }
            tmem_st_x16(s_addr + c * 16, pk);
            uint32_t t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = add_bf16x2(pk[2 * k], pk[2 * k + 1]);
#pragma unroll
            for (int k = 0; k < 4; ++k) t8[k] = add_bf16x2(t8[2 * k], t8[2 * k + 1]);
            ls[0] = add_f32x2(ls[0], unpack_bf16x2(add_bf16x2(t8[0], t8[1])));
            ls[1] = add_f32x2(ls[1], unpack_bf16x2(add_bf16x2(t8[2], t8[3])));
          }
        } else {
          // fp32 exponentials; EMU of every 8 run as a polynomial on the FMA pipe instead of the SFU
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 x = fma_f32x2(make_float2(s[c * 32 + k], s[c * 32 + k + 1]), l2e2, noff2);
              float2 e;
              if ((k & 7) < EMU) {  // EMU is even: whole pairs go to the FMA pipe
                e = poly_exp2_x2(x);
              } else {
                e.x = fast_exp2(x.x);
                e.y = fast_exp2(x.y);
              }
              ls[(k >> 1) & 1] = add_f32x2(ls[(k >> 1) & 1], e);
              pk[k >> 1] = pack_bf16(e.x, e.y);
            }
            tmem_st_x16(s_addr + c * 16, pk);
          }
        }
        FA_STAMP(4);
        if (j > 0) {
          // O_t may only be touched once P_t(j-1) V has retired, and must be rescaled before P_t(j) V is issued --
          // which cannot happen before this thread arrives below.  Waited every block (it has long completed by
          // now) so that the phase parity never runs ahead of this thread.
          mbar_wait(&pv_done[t], (pv_base + j - 1) & 1);
          if (__any_sync(0xffffffffu, grow)) {
            tc_fence_after();
            uint32_t o[32];
#pragma unroll
            for (int c = 0; c < HD; c += 32) {
              tmem_ld_x32(o_addr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * scale);
              tmem_st_x32(o_addr + c, o);
            }
          }
        }
        tmem_st_wait();
        FA_STAMP(5);
        l += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
        FA_STAMP(6);
      };
      // blocks [0, n_plain) need no mask: entirely below Tk and (causal) entirely left of the tile's diagonal
      int n_plain = min(nkv, p.Tk / BKV);
      if (p.causal) n_plain = min(n_plain, max(0, (q0 + t * BQ + off + 1) / BKV));
      for (int j = 0; j < n_plain; ++j) block(std::false_type{}, j);
      for (int j = n_plain; j < nkv; ++j) block(std::true_type{}, j);
#undef FA_STAMP
      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(&pv_done[t], (pv_base + nkv - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      if (p.lse != nullptr && q_idx < p.Tq)
        p.lse[((long long)b * p.H + h) * p.Tq + q_idx] = fmaf(m_used, kLog2e, log2f(l));
      uint32_t o0[32], o1[32];
      tmem_ld_x32(o_addr, o0);
      tmem_ld_x32(o_addr + 32, o1);
      tmem_ld_wait_regs(o0);
      tmem_ld_wait_regs(o1);
      // O_t is in registers: the P V issuer may start the next item's accumulation into it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)q_idx * p.o_rs + h * HD;
      if (q_idx < p.Tq) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o0[k]) * inv, __uint_as_float(o0[k + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o0[k + 2]) * inv, __uint_as_float(o0[k + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o0[k + 4]) * inv, __uint_as_float(o0[k + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o0[k + 6]) * inv, __uint_as_float(o0[k + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + k) = v;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o1[k]) * inv, __uint_as_float(o1[k + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o1[k + 2]) * inv, __uint_as_float(o1[k + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o1[k + 4]) * inv, __uint_as_float(o1[k + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o1[k + 6]) * inv, __uint_as_float(o1[k + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 32 + k) = v;
        }
      }
      pv_base += (uint32_t)nkv;
      }
      // buffer rotation / parity state after this item's n0 + n1 productions
      {
        const int adv = prod_base + n0 + n1;
        round_base = (round_base + (uint32_t)(adv / S_BUFS)) & 1;
        prod_base = adv % S_BUFS;
      }
      first_item = false;
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EMU>
int launch(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
           const AttnParams& p, cudaStream_t stream) {
  auto kfn = attention_fa_kernel<EMU>;
  static DeviceOnce attr_once;
  if (attr_once.first(ctx)) {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  const int n_items = ceil_div(p.Tq, 2 * BQ) * p.H * p.B;
  const int grid = n_items < ctx->num_sms ? n_items : ctx->num_sms;  // persistent: one CTA per SM
  kfn<<<grid, kThreads, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p, n_items);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

}  // namespace

int launch_attention_fa(dicow_ctx* ctx, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                        const AttnParams& p, int emu, cudaStream_t stream) {
  switch (emu) {
    // default: every exponential on the SFU.  Measured by looping each variant alone for 2 s at the bench shape (NVML energy
    // counter, tools/energy_probe.py; every variant draws the ~1 kW board cap, so inside the power-capped encoder step a
    // launch costs its JOULES, not its isolated time): EMU 0 = 500 mJ / launch, EMU 2 = 509 mJ (0.4 % faster alone at
    // burst clocks, 2 % more energy), EMU 4 = 544 mJ, packed-bf16 exponentials = 579 mJ.
    case 0: return launch<0>(ctx, tmQ, tmK, tmV, p, stream);
    case 1: return launch<2>(ctx, tmQ, tmK, tmV, p, stream);   // 2 of 8 exponentials as a polynomial on the FMA pipe
    // packed bf16 exponentials: ptxas 12.9 lowers ex2.approx.ftz.bf16x2 to TWO MUFU.EX2.BF16 (+ PRMT) on sm_100a, so there
    // is no 2-per-SFU-op gain; kept as a measured variant
    case 4: return launch<-1>(ctx, tmQ, tmK, tmV, p, stream);
    case 2: return launch<4>(ctx, tmQ, tmK, tmV, p, stream);
    case 3: return launch<6>(ctx, tmQ, tmK, tmV, p, stream);
    default: return set_error(ctx, DICOW_ERR_INVALID_ARG, "dicow_attention_bf16: bad poly-exp share %d", emu);
  }
}

}  // namespace dicow
