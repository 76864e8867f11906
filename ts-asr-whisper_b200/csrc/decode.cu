// decode.cu -- HBM-bound kernels of the token-by-token decoder step (greedy generate()):
//   * skinny GEMM  out[M <= 64, N] = epilogue(A[M, K] W[N, K]^T): every weight byte is read exactly once from HBM and
//     multiplied against all M rows (the batch) with warp-level mma.sync m16n8k16 -- at M = 16 the CUDA-core FMA rate
//     would cap the kernel below the HBM roofline, and a 128-row tcgen05 tile would leave most SMs without work
//     (N / 256 tiles) while re-reading W once per batch row group;
//   * decode attention: one query row per (batch, head) against a K/V cache (self: Tk = pos + 1; cross: Tk = 1500);
//   * token + position embedding gather;
//   * fused logits rules + argmax: SuppressTokensLogitsProcessor + WhisperTimeStampLogitsProcessor + the DiCoW
//     EOS-at-begin exception + argmax + finished-row handling in one pass over the [B, V] logits, no host sync.
// Every kernel can take the current position from a device scalar, so a whole decode step is a fixed sequence of
// launches that is captured once in a CUDA graph and replayed per token.
//
// Replaces, per generated token, HF WhisperDecoder.forward with a KV cache (HF:models/whisper/modeling_whisper.py:
// 449-506, 691-796), proj_out (src/models/dicow/modeling_dicow.py:302) and the greedy branch of
// DiCoWGenerationMixin._sample (src/models/dicow/generation.py:707-782) with its logits processors
// (HF:generation/logits_process.py:1905-2043; src/models/dicow/utils.py:5-14).
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// skinny GEMM
// ------------------------------------------------------------------------------------------------------------------
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;

struct SkinnyParams {
  const __nv_bfloat16* A;
  long long lda;
  const __nv_bfloat16* W;
  long long ldw;
  int M, N, K;
  const float* bias;
  void* out;
  long long ldo;
  int epilogue;
  const float* resid;
  long long ldr;
  const int* pos;
  long long pos_stride;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {  // weights: read once, do not pollute L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void cp_async_cg_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, uint2 v) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v),
               "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ void cluster_arrive_all() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_all() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_s32(uint32_t cluster_addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// One CTA = 8 output columns; its 8 warps split K in 32-element blocks (warp w takes blocks w, w + 8, ...).  Lane
// (g = lane / 4, t = lane % 4) loads 16 contiguous bytes of W row n0 + g at k = 32 kb + 8 t and the same bytes of A rows
// g and g + 8 of every 16-row tile; the 8 k values feed two m16n8k16 MMAs with a k permutation that is identical on the
// A and B side (so the contraction is unchanged).  Partial sums meet in shared memory.
template <int MT>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const SkinnyParams p) {
  __shared__ float red[SK_WARPS][MT * 16][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8;
  const int nrow = n0 + g;
  const bool n_ok = nrow < p.N;
  const __nv_bfloat16* wrow = p.W + (long long)(n_ok ? nrow : 0) * p.ldw + t * 8;
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
  const int kblocks = p.K >> 5;
#pragma unroll 4
  for (int kb = warp; kb < kblocks; kb += SK_WARPS) {
    uint4 wv = make_uint4(0u, 0u, 0u, 0u);
    if (n_ok) wv = ldg_stream_128(wrow + kb * 32);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int r0 = m * 16 + g, r1 = r0 + 8;
      uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
      if (r0 < p.M) lo = __ldg(reinterpret_cast<const uint4*>(p.A + (long long)r0 * p.lda + kb * 32 + t * 8));
      if (r1 < p.M) hi = __ldg(reinterpret_cast<const uint4*>(p.A + (long long)r1 * p.lda + kb * 32 + t * 8));
      mma_bf16_16816(acc[m], lo.x, hi.x, lo.y, hi.y, wv.x, wv.y);
      mma_bf16_16816(acc[m], lo.z, hi.z, lo.w, hi.w, wv.z, wv.w);
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    red[warp][m * 16 + g][2 * t] = acc[m][0];
    red[warp][m * 16 + g][2 * t + 1] = acc[m][1];
    red[warp][m * 16 + g + 8][2 * t] = acc[m][2];
    red[warp][m * 16 + g + 8][2 * t + 1] = acc[m][3];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MT * 16 * 8; i += SK_THREADS) {
    const int row = i >> 3, col = i & 7;
    const int n = n0 + col;
    if (row >= p.M || n >= p.N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += red[w][row][col];
    if (p.bias != nullptr) v += __ldg(p.bias + n);
    long long o = (long long)row * p.ldo + n;
    if (p.pos != nullptr) o += (long long)(*p.pos) * p.pos_stride;
    switch (p.epilogue) {
      case DICOW_EPI_BIAS_BF16: reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v); break;
      case DICOW_EPI_BIAS_GELU_BF16:
        reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(gelu_erf_fast(v));
        break;
      case DICOW_EPI_RESIDUAL_F32:
        reinterpret_cast<float*>(p.out)[o] = p.resid[(long long)row * p.ldr + n] + v;
        break;
      default: reinterpret_cast<float*>(p.out)[o] = v; break;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// decode_linear: [LayerNorm ->] Linear for the M <= 64 rows of one decode step, with everything around the weight
// stream folded in (one kernel instead of LayerNorm + 1..2 skinny GEMMs):
//   * the CTA's first weight slab is requested BEFORE griddep_wait(): weights do not depend on the previous kernel, so
//     their HBM latency overlaps its tail;
//   * the A operand is staged once per CTA in shared memory as bf16 -- either LayerNorm(x) computed in the prologue from
//     the fp32 residual stream (two-pass statistics, one warp per row: the arithmetic of fddt_ln_kernel), or a copy of
//     a bf16 activation -- with a row pitch = 64 (mod 128) bytes so that the fragment loads are bank-conflict free;
//   * a CTA owns `tiles_per_cta` consecutive 8-column tiles (large N: proj_out has 6 484) and prefetches the next
//     tile's weights into registers while the current one is multiplied, so A is staged ~600 times per step, not 6 484;
//   * small-N / large-K layers (out_proj, fc2) split K over a thread-block cluster of `ksplit` CTAs; partial sums
//     meet in rank 0's shared memory (DSMEM) and are added in rank order (deterministic, no atomics);
//   * columns >= n_split go to a second output (fused q | k,v projection: q to its buffer, k,v appended to the cache).
// ------------------------------------------------------------------------------------------------------------------
constexpr int DL_MAX_KS = 1280;   // K / ksplit: the CTA's k-slice, held in registers as 32-wide k-blocks dealt to the warps
constexpr int DL_MAX_KSPLIT = 8;  // portable cluster size

struct DecLinParams {
  const float* x;  // LayerNorm source [M, K] fp32 (LN variant)
  long long ldx;
  const float* gamma;
  const float* beta;
  float eps;
  const __nv_bfloat16* A;  // bf16 source [M, K] (plain variant)
  long long lda;
  const __nv_bfloat16* W;
  long long ldw;
  int M, N, K;
  const float* bias;
  void* out;
  long long ldo;
  int epilogue;
  const float* resid;
  long long ldr;
  int n_split;  // columns >= n_split go to out2 (n_split = N: single output)
  void* out2;
  long long ldo2;
  const int* pos;  // out2 base advanced by *pos * pos_stride elements (KV-cache append); applies to out when out2 == NULL
  long long pos_stride;
  int ksplit, tiles_per_cta, Ks, a_pitch;  // Ks = K / ksplit; a_pitch = shared-memory row pitch of A in elements
};

// bias / resid / pos_off: the epilogue's operands, requested by the caller before the multiply (they do not depend on it)
__device__ __forceinline__ void dl_store(const DecLinParams& p, int row, int n, float v, float bias, float resid, long long pos_off) {
  if (row >= p.M || n >= p.N) return;
  v += bias;
  void* dst = p.out;
  long long o = (long long)row * p.ldo + n;
  if (p.out2 != nullptr && n >= p.n_split) {
    dst = p.out2;
    o = (long long)row * p.ldo2 + (n - p.n_split) + pos_off;
  } else if (p.out2 == nullptr) {
    o += pos_off;
  }
  switch (p.epilogue) {
    case DICOW_EPI_BIAS_BF16: reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v); break;
    case DICOW_EPI_BIAS_GELU_BF16: reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(gelu_erf_fast(v)); break;
    case DICOW_EPI_RESIDUAL_F32: reinterpret_cast<float*>(dst)[o] = resid + v; break;
    default: reinterpret_cast<float*>(dst)[o] = v; break;
  }
}

// NT = 8-column tiles a CTA multiplies AT ONCE (single pass, tiles_per_cta == NT): the layers' linears at M <= 32 take
// NT = 2..4 so that ~one CTA per SM has its whole weight share in flight from its first instruction and the A slab is
// staged 2-4 x less often (it was 1-2 x the weight bytes in L2 -> SM traffic); NT = 1 loops over tiles_per_cta tiles with
// the next tile's weights prefetched (proj_out).
// NW = warps per CTA (8 or 16): warp w multiplies k-blocks w, w + NW, ...  The single-pass variants take 16: a warp of
// this kernel executes a few hundred dependent integer / address instructions around its loads, and with two warps per
// scheduler nothing hides their latency (ncu, profiles/r02c_ncu_decode_linear.md: issue slots 10 % busy, "wait" stalls 20 %).
template <int MT, bool LN, int NT, int NW>
__global__ void __launch_bounds__(NW * 32) decode_linear_kernel(const DecLinParams p) {
  constexpr int DL_WARPS = NW, DL_THREADS = NW * 32;
  constexpr int DL_KBW = (DL_MAX_KS / 32 + NW - 1) / NW;  // k-blocks per warp: 5 (8 warps) or 3 (16 warps)
  constexpr int TE = MT * 16 * 8;  // output elements of one tile
  constexpr int GE = NT * TE;      // ... of the tiles of one pass
  extern __shared__ __align__(16) unsigned char dl_smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(dl_smem);
  float* red = reinterpret_cast<float*>(dl_smem + (size_t)MT * 16 * p.a_pitch * sizeof(__nv_bfloat16));
  float* cpart = red + DL_WARPS * GE;  // [ksplit - 1][GE], used on cluster rank 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int krank = p.ksplit > 1 ? (int)cluster_ctarank() : 0;
  const int m0 = (int)blockIdx.y * (MT * 16);  // M > 32 rows are split over blockIdx.y (two 32-row CTAs per column tile)
  const int ntiles = (p.N + 7) >> 3;
  const int tile_first = ((int)blockIdx.x / p.ksplit) * p.tiles_per_cta;
  const int tile_end = min(tile_first + p.tiles_per_cta, ntiles);
  const int kblocks = p.Ks >> 5;
  const int kbase = krank * p.Ks;

  // cluster launches (LayerNorm sharing, split K) write into peer CTAs' shared memory: arrive now, wait right before the
  // first remote store, so that "the peer has started" costs no latency
  __shared__ __align__(8) uint64_t split_bar;  // rank 0: counts the bytes of the peers' partial sums (split K)
  if (p.ksplit > 1) {
    if (krank == 0 && threadIdx.x == 0) {
      mbar_init(&split_bar, 1);
      fence_barrier_init();
    }
    cluster_arrive_relaxed();  // (a release here is a MEMBAR.ALL.GPU: 10 % of the kernel's stall samples)
  }
  uint4 wv[NT][DL_KBW], wn[DL_KBW];
  auto load_w = [&](int tile, uint4(&w)[DL_KBW]) {
    const int nrow = tile * 8 + g;
    const bool ok = tile < tile_end && nrow < p.N;
    const __nv_bfloat16* wrow = p.W + (long long)(ok ? nrow : 0) * p.ldw + kbase + t * 8;
#pragma unroll
    for (int i = 0; i < DL_KBW; ++i) {
      const int kb = warp + i * DL_WARPS;
      w[i] = make_uint4(0u, 0u, 0u, 0u);
      if (ok && kb < kblocks) w[i] = ldg_stream_128(wrow + kb * 32);
    }
  };
#pragma unroll
  for (int j = 0; j < NT; ++j) load_w(tile_first + j, wv[j]);  // in flight across the dependency wait
  if constexpr (LN) {  // gamma | beta -> shared memory, once per CTA (fetched per use they were twenty load round trips in a row)
    float* sGB = cpart + (size_t)(p.ksplit - 1) * GE;
    const int nvec = p.K >> 2;
    for (int c = threadIdx.x; c < 2 * nvec; c += DL_THREADS)
      cp_async_cg_16(smem_u32(sGB + 4 * c), (c < nvec ? p.gamma : p.beta - p.K) + 4 * c);
  }
  griddep_launch();
  griddep_wait();

  // ---- stage A (bf16) in shared memory: rows >= M are zero ----
  if constexpr (LN) {
    // Opt-in form (fused_decode_step = "ln_prologue"); the default is the few-rows LayerNorm kernel in front of this one.
    // Every CTA normalises the rows itself (one warp per row, two rows of a warp in flight, two-pass statistics in
    // registers: the arithmetic of fddt_ln_kernel) and keeps the bf16 columns of its k-slice: 80 KB of fp32 rows from L2
    // per CTA instead of 40 KB of bf16.  Measured at B = 16: 10.9-14.4 us per layer linear against 2.1 (LayerNorm kernel)
    // + 4.0-6.3 us -- the ~160 CTAs re-reading the same rows cost more than the launch they replace.  (Round 1 shared the
    // LayerNorm over a cluster of 8 through DSMEM instead: 15-17 us, half of it in the cluster barriers' fences.)
    const int nvec = p.K >> 2;  // K <= 1280: at most 10 float4 per lane
    const int c_lo = kbase >> 2, c_hi = (kbase + p.Ks) >> 2;
    const float4* sG = reinterpret_cast<const float4*>(cpart + (size_t)(p.ksplit - 1) * GE);  // gamma | beta, staged at kernel entry
    const float4* sB = sG + nvec;
    float4 v[2][10];
    auto load_rows = [&](int r0) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = r0 + u * DL_WARPS;
        const bool ok = r < MT * 16 && m0 + r < p.M;
        const float4* xr = reinterpret_cast<const float4*>(p.x + (long long)(ok ? m0 + r : 0) * p.ldx);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int c = lane + 32 * i;
          v[u][i] = (ok && c < nvec) ? __ldcg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    load_rows(warp);
    cp_async_wait_all();  // gamma / beta
    __syncthreads();
    for (int r0 = warp;;) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = r0 + u * DL_WARPS;
        if (r >= MT * 16) break;
        __nv_bfloat16* srow = sA + (size_t)r * p.a_pitch;
        if (m0 + r >= p.M) {  // rows past M multiply as zeros
          for (int c = lane * 4; c < p.Ks; c += 128) *reinterpret_cast<uint2*>(srow + c) = make_uint2(0u, 0u);
          continue;
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) sum += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
        const float mean = warp_sum(sum) / (float)p.K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          if (lane + 32 * i < nvec) {
            const float a = v[u][i].x - mean, b2 = v[u][i].y - mean, c = v[u][i].z - mean, e = v[u][i].w - mean;
            q += (a * a + b2 * b2) + (c * c + e * e);
          }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)p.K + p.eps);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int c4 = lane + 32 * i;
          if (c4 >= c_lo && c4 < c_hi) {
            const float4 ga = sG[c4], be = sB[c4];
            const float y0 = fmaf((v[u][i].x - mean) * rstd, ga.x, be.x), y1 = fmaf((v[u][i].y - mean) * rstd, ga.y, be.y);
            const float y2 = fmaf((v[u][i].z - mean) * rstd, ga.z, be.z), y3 = fmaf((v[u][i].w - mean) * rstd, ga.w, be.w);
            *reinterpret_cast<uint2*>(srow + (size_t)(c4 - c_lo) * 4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
          }
        }
      }
      r0 += 2 * DL_WARPS;
      if (r0 >= MT * 16) break;
      load_rows(r0);
    }
  } else {
    // cp.async (L2 only, like ld.global.cg): every 16-byte piece of the slab is in flight at once.  The load + st.shared
    // loop this replaces compiled to one L2 round trip per piece and thread (ten in a row for 16 rows of K = 1280), which
    // was most of the duration of a layer's linear kernel.
    const int vec_per_row = p.Ks >> 3;
    for (int r = warp; r < MT * 16; r += DL_WARPS) {  // a warp copies whole rows: no index division, 512 contiguous bytes per request
      __nv_bfloat16* drow = sA + (size_t)r * p.a_pitch;
      if (m0 + r < p.M) {
        const __nv_bfloat16* srow = p.A + (long long)(m0 + r) * p.lda + kbase;
        for (int c = lane; c < vec_per_row; c += 32) cp_async_cg_16(smem_u32(drow + c * 8), srow + c * 8);
      } else {
        for (int c = lane; c < vec_per_row; c += 32) *reinterpret_cast<uint4*>(drow + c * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    cp_async_wait_all();
  }
  // the KV-cache append position: one L2 read here instead of one per stored element after the multiply
  const long long pos_off = p.pos != nullptr ? (long long)__ldcg(p.pos) * p.pos_stride : 0ll;
  __syncthreads();
  if (p.ksplit > 1) {
    // the one arrival of the phase, with the byte count of the peers' partial sums (bytes that land before this are
    // counted against it: the transaction count may run negative)
    if (krank == 0 && threadIdx.x == 0) mbar_arrive_expect_tx(&split_bar, (uint32_t)((p.ksplit - 1) * GE * sizeof(float)));
    cluster_wait_all();
  }

  constexpr int NV = (GE + DL_THREADS - 1) / DL_THREADS;  // output elements per thread and pass
  for (int tile = tile_first; tile < tile_end; tile += NT) {
    if constexpr (NT == 1) load_w(tile + 1, wn);  // next tile's weights stream in while this one is multiplied (zeros past the last tile)
    // bias and residual of the elements this thread stores: requested now, needed after the cross-warp / cluster sums
    float ebias[NV], eres[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = threadIdx.x + j * DL_THREADS;
      const int tj = e / TE, et = e - tj * TE;
      const int row = m0 + (et >> 3), n = (tile + tj) * 8 + (et & 7);
      ebias[j] = eres[j] = 0.f;
      if (krank == 0 && e < GE && tile + tj < tile_end && row < p.M && n < p.N) {
        if (p.bias != nullptr) ebias[j] = __ldg(p.bias + n);
        if (p.epilogue == DICOW_EPI_RESIDUAL_F32) eres[j] = __ldcg(p.resid + (long long)row * p.ldr + n);
      }
    }
    float acc[NT][MT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[j][m][0] = acc[j][m][1] = acc[j][m][2] = acc[j][m][3] = 0.f;
#pragma unroll
    for (int i = 0; i < DL_KBW; ++i) {
      const int kb = warp + i * DL_WARPS;
      if (kb < kblocks) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const __nv_bfloat16* a0 = sA + (size_t)(m * 16 + g) * p.a_pitch + kb * 32 + t * 8;
          const uint4 lo = *reinterpret_cast<const uint4*>(a0);
          const uint4 hi = *reinterpret_cast<const uint4*>(a0 + 8 * (size_t)p.a_pitch);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            mma_bf16_16816(acc[j][m], lo.x, hi.x, lo.y, hi.y, wv[j][i].x, wv[j][i].y);
            mma_bf16_16816(acc[j][m], lo.z, hi.z, lo.w, hi.w, wv[j][i].z, wv[j][i].w);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float* r0 = red + (size_t)warp * GE + j * TE + (m * 16 + g) * 8 + 2 * t;
        r0[0] = acc[j][m][0], r0[1] = acc[j][m][1];
        r0[64] = acc[j][m][2], r0[65] = acc[j][m][3];  // row g + 8
      }
    __syncthreads();
    // cross-warp sums; with split K the partial sums of ranks 1.. meet in rank 0's shared memory and are added in rank order
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = threadIdx.x + j * DL_THREADS;
      v[j] = 0.f;
      if (e < GE) {
#pragma unroll
        for (int w = 0; w < DL_WARPS; ++w) v[j] += red[(size_t)w * GE + e];
        // split K: the partial sum travels to rank 0's shared memory as an asynchronous store that reports its bytes to
        // rank 0's mbarrier -- no cluster barrier (and none of its memory fences) after the multiply; ranks > 0 are done
        if (krank != 0)
          st_async_b32(map_to_cta(smem_u32(cpart + (size_t)(krank - 1) * GE + e), 0), __float_as_uint(v[j]),
                       map_to_cta(smem_u32(&split_bar), 0));
      }
    }
    if (p.ksplit > 1 && krank == 0) mbar_wait(&split_bar, 0);  // (split K is single-pass: one phase)
    if (krank == 0) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int e = threadIdx.x + j * DL_THREADS;
        const int tj = e / TE, et = e - tj * TE;
        if (e < GE && tile + tj < tile_end) {
          for (int r = 1; r < p.ksplit; ++r) v[j] += cpart[(size_t)(r - 1) * GE + e];
          dl_store(p, m0 + (et >> 3), (tile + tj) * 8 + (et & 7), v[j], ebias[j], eres[j], pos_off);
        }
      }
    }
    if constexpr (NT == 1) {
      __syncthreads();  // red is rewritten by the next tile
#pragma unroll
      for (int k = 0; k < DL_KBW; ++k) wv[0][k] = wn[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// decode attention: softmax(q . K^T) V for one query row per (batch, head), head_dim 64
// ------------------------------------------------------------------------------------------------------------------
// DA_WARPS warps per (batch, head): 4 for the short self-attention cache, 8 for the 1500-key cross attention.  The kernel
// is a pure K/V stream bound by the loads in flight per SM: B x H = 320 CTAs must be ONE wave (16 warps at 72 registers
// fit one CTA per SM -> 2.2 waves, measured 42 us against 36 us with 4 warps; 8 warps fit three per SM)

struct DecAttnParams {
  const __nv_bfloat16* Q;
  long long q_bs;
  const __nv_bfloat16* K;
  const __nv_bfloat16* V;
  long long kv_rs, kv_bs, kv_hs;
  __nv_bfloat16* out;
  long long o_bs;
  int Tk;
  const int* pos;
  int kv_batch_div;      // K/V batch index = query row / kv_batch_div (beam search: the beams of an utterance share its cross K/V)
  const int* ancestry;   // [B, anc_stride] or NULL: key t of query row b lives in cache row ancestry[b][t] (beam search
  long long anc_stride;  //  self-attention: hypotheses share prefixes without the cache ever being re-ordered)
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x, f[2 * i + 1] = v.y;
  }
}

// grid (H, B), 4 warps.  A group of 8 lanes owns one key at a time (16 B = 8 dims per lane, so every K / V row is
// read as one full 128-byte line); each group keeps an online-softmax state (m, l, o[8 dims per lane]); groups and
// warps are merged at the end.
template <int DA_WARPS>
__global__ void __launch_bounds__(DA_WARPS * 32) decode_attention_kernel(const DecAttnParams p) {
  __shared__ float sm_m[DA_WARPS], sm_l[DA_WARPS], sm_o[DA_WARPS][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  griddep_launch();
  griddep_wait();
  const int Tk = p.pos != nullptr ? (__ldcg(p.pos) + 1) : p.Tk;
  float q[8];
  unpack8(__ldcg(reinterpret_cast<const uint4*>(p.Q + (long long)b * p.q_bs + h * 64 + sub * 8)), q);
  const int bkv = p.kv_batch_div > 1 ? b / p.kv_batch_div : b;
  const long long head_off = (long long)h * p.kv_hs + sub * 8;
  const __nv_bfloat16* kb = p.K + (long long)bkv * p.kv_bs + head_off;
  const __nv_bfloat16* vb = p.V + (long long)bkv * p.kv_bs + head_off;
  const int* anc = p.ancestry != nullptr ? p.ancestry + (long long)b * p.anc_stride : nullptr;
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  constexpr int UNROLL = 4;
  const int stride = DA_WARPS * 4;
  for (int kw = warp * 4; kw < Tk; kw += stride * UNROLL) {  // warp-uniform trip count (shuffles below)
    const int k0 = kw + grp;
    uint4 kv[UNROLL], vv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int k = k0 + u * stride;
      if (k < Tk) {
        long long off = (long long)k * p.kv_rs;
        if (anc != nullptr) off += (long long)(__ldcg(anc + k) - bkv) * p.kv_bs;  // the cache row that holds key k
        kv[u] = __ldcg(reinterpret_cast<const uint4*>(kb + off));  // L2 only: the newest row was
        vv[u] = __ldcg(reinterpret_cast<const uint4*>(vb + off));  // written by the previous kernel
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int k = k0 + u * stride;
      const bool ok = k < Tk;  // uniform within the 8-lane group
      float kf[8];
      float s = 0.f;
      if (ok) {
        unpack8(kv[u], kf);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok) {
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn);  // 0 for the first key (m = -inf)
        const float pw = __expf(s - mn);
        float vf[8];
        unpack8(vv[u], vf);
        l = fmaf(l, corr, pw);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(o[i], corr, pw * vf[i]);
        m = mn;
      }
    }
  }
  // merge the 4 groups of the warp (lanes with the same `sub`)
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
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
    for (int i = 0; i < 8; ++i) sm_o[warp][sub * 8 + i] = o[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float mt = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) mt = fmaxf(mt, sm_m[w]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float c = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - mt);
      lt = fmaf(sm_l[w], c, lt);
      ot = fmaf(sm_o[w][threadIdx.x], c, ot);
    }
    p.out[(long long)b * p.o_bs + h * 64 + threadIdx.x] = __float2bfloat16_rn(ot / lt);
  }
}

// Cross attention over the head-major cache ([B, H, T, k(64) | v(64)]: ONE contiguous 256 B x T stream per (batch, head)): the
// stream goes through a shared-memory ring filled by cp.async.bulk instead of through registers.  With the register version a
// CTA keeps 32 KB of loads in flight and streams at a fixed ~15 GB/s whatever runs next to it (tools/probe_decode_attn.py:
// 160, 320 and 440 CTAs take 26, 31 and 33 us) -- B x H = 320 CTAs on 148 SMs (2.2 per SM, three fit) leave the device at
// 3.95 TB/s.  Here a CTA keeps CX_STAGES x 16 KB in flight without holding a register for it.
// Same mapping as decode_attention_kernel: an 8-lane group owns a key (lane = 8 dims), online softmax per group.
constexpr int CX_WARPS = 8;       // consumer warps (+ 1 producer warp)
constexpr int CX_KEYS = 64;       // keys per stage: 16 KB
constexpr int CX_STAGES = 4;

__device__ __forceinline__ void cx_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__((CX_WARPS + 1) * 32) decode_cross_attention_ring_kernel(const DecAttnParams p) {
  extern __shared__ __align__(128) uint8_t cx_smem[];
  __shared__ uint64_t full[CX_STAGES], empty[CX_STAGES];
  __shared__ float sm_m[CX_WARPS], sm_l[CX_WARPS], sm_o[CX_WARPS][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CX_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  griddep_launch();
  const int Tk = p.Tk;
  const int nchunks = (Tk + CX_KEYS - 1) / CX_KEYS;
  const int bkv = p.kv_batch_div > 1 ? b / p.kv_batch_div : b;
  const __nv_bfloat16* stream = p.K + (long long)bkv * p.kv_bs + (long long)h * p.kv_hs;  // 128 elements per key: k | v
  if (warp == CX_WARPS) {
    // ===================== producer: the cache is immutable during the step (written once per window) =====================
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % CX_STAGES;
      mbar_wait(&empty[st], ((c / CX_STAGES) & 1) ^ 1);
      if (elect_one()) {
        const int nk = min(CX_KEYS, Tk - c * CX_KEYS);
        const uint32_t bytes = (uint32_t)nk * 256u;
        mbar_arrive_expect_tx(&full[st], bytes);
        cx_bulk_g2s(cx_smem + st * CX_KEYS * 256, stream + (long long)c * CX_KEYS * 128, bytes, &full[st]);
      }
      __syncwarp();
    }
    return;
  }
  const int grp = lane >> 3, sub = lane & 7;
  griddep_wait();  // q comes from the previous kernel of the step
  float q[8];
  unpack8(__ldcg(reinterpret_cast<const uint4*>(p.Q + (long long)b * p.q_bs + h * 64 + sub * 8)), q);
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int st = c % CX_STAGES;
    mbar_wait(&full[st], (c / CX_STAGES) & 1);
    const uint8_t* base = cx_smem + st * CX_KEYS * 256 + sub * 16;
    constexpr int PER = CX_KEYS / (CX_WARPS * 4);  // keys of this chunk per 8-lane group
    uint4 kv[PER], vv[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int kl = (u * CX_WARPS + warp) * 4 + grp;  // key within the chunk
      kv[u] = *reinterpret_cast<const uint4*>(base + kl * 256);
      vv[u] = *reinterpret_cast<const uint4*>(base + kl * 256 + 128);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // the stage is in registers
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int k = c * CX_KEYS + (u * CX_WARPS + warp) * 4 + grp;
      const bool ok = k < Tk;  // uniform within the 8-lane group
      float kf[8];
      float s = 0.f;
      if (ok) {
        unpack8(kv[u], kf);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok) {
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn);  // 0 for the first key (m = -inf)
        const float pw = __expf(s - mn);
        float vf[8];
        unpack8(vv[u], vf);
        l = fmaf(l, corr, pw);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(o[i], corr, pw * vf[i]);
        m = mn;
      }
    }
  }
  // merge the 4 groups of the warp (lanes with the same `sub`), then the warps
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
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
    for (int i = 0; i < 8; ++i) sm_o[warp][sub * 8 + i] = o[i];
  }
  named_bar_sync(1, CX_WARPS * 32);  // the producer warp has left
  if (threadIdx.x < 64) {
    float mt = -INFINITY;
#pragma unroll
    for (int w = 0; w < CX_WARPS; ++w) mt = fmaxf(mt, sm_m[w]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int w = 0; w < CX_WARPS; ++w) {
      const float c = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - mt);
      lt = fmaf(sm_l[w], c, lt);
      ot = fmaf(sm_o[w][threadIdx.x], c, ot);
    }
    p.out[(long long)b * p.o_bs + h * 64 + threadIdx.x] = __float2bfloat16_rn(ot / lt);
  }
}

// Several queries against ONE K/V stream: beam search, where the NQ hypotheses of an utterance attend to the same encoder
// states (cross attention).  One CTA per (head, utterance): every key / value row is loaded once and used by all NQ
// queries (the single-query kernel re-reads the 384 KB stream once per hypothesis: 5 x at 5 beams, 74 us per layer for
// 60 hypotheses).  Same thread mapping as decode_attention_kernel: an 8-lane group owns a key.  NOT the default: see the
// measurement at the launch site.
template <int NQ>
__global__ void __launch_bounds__(256) decode_attention_mq_kernel(const DecAttnParams p) {
  constexpr int NWARPS = 8;
  __shared__ float sm_m[NWARPS][NQ], sm_l[NWARPS][NQ], sm_o[NWARPS][NQ][64];
  const int h = blockIdx.x, u = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  griddep_launch();
  griddep_wait();
  const int Tk = p.Tk;
  float q[NQ][8];
#pragma unroll
  for (int n = 0; n < NQ; ++n)
    unpack8(__ldcg(reinterpret_cast<const uint4*>(p.Q + (long long)(u * NQ + n) * p.q_bs + h * 64 + sub * 8)), q[n]);
  const __nv_bfloat16* kb = p.K + (long long)u * p.kv_bs + (long long)h * p.kv_hs + sub * 8;
  const __nv_bfloat16* vb = p.V + (long long)u * p.kv_bs + (long long)h * p.kv_hs + sub * 8;
  float m[NQ], l[NQ], o[NQ][8];
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
    m[n] = -INFINITY, l[n] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[n][i] = 0.f;
  }
  constexpr int UNROLL = 2;
  const int stride = NWARPS * 4;
  for (int kw = warp * 4; kw < Tk; kw += stride * UNROLL) {  // warp-uniform trip count
    const int k0 = kw + grp;
    uint4 kv[UNROLL], vv[UNROLL];
#pragma unroll
    for (int x = 0; x < UNROLL; ++x) {
      const int k = k0 + x * stride;
      if (k < Tk) {
        kv[x] = __ldcg(reinterpret_cast<const uint4*>(kb + (long long)k * p.kv_rs));
        vv[x] = __ldcg(reinterpret_cast<const uint4*>(vb + (long long)k * p.kv_rs));
      }
    }
#pragma unroll
    for (int x = 0; x < UNROLL; ++x) {
      const int k = k0 + x * stride;
      const bool ok = k < Tk;  // uniform within the 8-lane group
      float kf[8], vf[8];
      if (ok) unpack8(kv[x], kf), unpack8(vv[x], vf);
#pragma unroll
      for (int n = 0; n < NQ; ++n) {
        float s = 0.f;
        if (ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(q[n][i], kf[i], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ok) {
          const float mn = fmaxf(m[n], s);
          const float corr = __expf(m[n] - mn);
          const float pw = __expf(s - mn);
          l[n] = fmaf(l[n], corr, pw);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[n][i] = fmaf(o[n][i], corr, pw * vf[i]);
          m[n] = mn;
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
#pragma unroll
    for (int x = 8; x <= 16; x <<= 1) {  // merge the 4 groups of the warp
      const float m2 = __shfl_xor_sync(0xffffffffu, m[n], x);
      const float l2 = __shfl_xor_sync(0xffffffffu, l[n], x);
      const float mn = fmaxf(m[n], m2);
      const float c1 = (m[n] == -INFINITY) ? 0.f : __expf(m[n] - mn);
      const float c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
      l[n] = l[n] * c1 + l2 * c2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float o2 = __shfl_xor_sync(0xffffffffu, o[n][i], x);
        o[n][i] = o[n][i] * c1 + o2 * c2;
      }
      m[n] = mn;
    }
    if (grp == 0) {
      if (sub == 0) sm_m[warp][n] = m[n], sm_l[warp][n] = l[n];
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_o[warp][n][sub * 8 + i] = o[n][i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NQ * 64; idx += blockDim.x) {
    const int n = idx >> 6, c = idx & 63;
    float mt = -INFINITY;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) mt = fmaxf(mt, sm_m[w][n]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const float cc = (sm_m[w][n] == -INFINITY) ? 0.f : __expf(sm_m[w][n] - mt);
      lt = fmaf(sm_l[w][n], cc, lt);
      ot = fmaf(sm_o[w][n][c], cc, ot);
    }
    p.out[(long long)(u * NQ + n) * p.o_bs + h * 64 + c] = __float2bfloat16_rn(ot / lt);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// cross-attention K/V re-layout, once per 30 s window: [B*T, (k | v) x H x 64] (the projection GEMM's output) ->
// [B, H, T, 128] with k | v of one key adjacent.  A decode step then streams, per (batch, head), ONE contiguous
// T x 256 B region instead of 2 x T 128-byte pieces 5 KB apart (measured: 3.0 -> see DESIGN.md TB/s in decode_attention).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) kv_to_head_major_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T,
                                                               int H) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int per_half = H * 8;  // 16-byte vectors per k (or v) row
  const uint4* src = in + ((long long)b * T + t) * (2 * per_half);
  for (int i = threadIdx.x; i < 2 * per_half; i += blockDim.x) {
    const int half = i / per_half, rem = i - half * per_half;
    const int head = rem >> 3, j = rem & 7;
    out[(((long long)b * H + head) * T + t) * 16 + half * 8 + j] = __ldg(src + i);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// x[b, s, :] = embed_tokens[ids[b, past + s], :] + embed_positions[past + s, :]      (fp32)
// ------------------------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const long long* __restrict__ ids, long long ids_rs, const float* __restrict__ tok,
                             const float* __restrict__ posw, float* __restrict__ x, int S, int d, int past,
                             const int* pos, int vocab) {
  const int s = blockIdx.x, b = blockIdx.y;
  griddep_launch();
  griddep_wait();
  const int pp = (pos != nullptr ? __ldcg(pos) : past) + s;
  long long id = __ldcg(ids + (long long)b * ids_rs + pp);
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* tr = reinterpret_cast<const float4*>(tok + id * d);
  const float4* pr = reinterpret_cast<const float4*>(posw + (long long)pp * d);
  float4* xr = reinterpret_cast<float4*>(x + ((long long)b * S + s) * d);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
    const float4 a = __ldg(tr + i), c = __ldg(pr + i);
    xr[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

__global__ void advance_kernel(int* pos, int by) {
  griddep_launch();
  griddep_wait();
  *pos = __ldcg(pos) + by;
}

// ------------------------------------------------------------------------------------------------------------------
// logits rules + argmax
// ------------------------------------------------------------------------------------------------------------------
struct RulesParams {
  const float* logits;
  long long ld;
  int V;
  long long* ids;
  long long ids_rs;
  const int* pos;  // column of the token that produced these logits; the new token goes to column *pos + 1
  int cur_len;     // used when pos == NULL: number of tokens already in ids
  int begin_index, eos, pad, no_timestamps, ts_begin, max_initial_ts, ts_rules;
  const unsigned* suppress;  // bitmap, ceil(V / 32) words, or NULL
  int* unfinished;           // [B] 1 = still decoding
  float* processed;          // optional [B, V] fp32 copy of the processed scores (tests)
  int no_select;             // only materialise the processed scores
};

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ Best best_of(Best a, Best b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = best_of(x, y);
  }
  return x;
}

// One CTA -- or a cluster of `csplit` CTAs, each scanning a slice of the vocabulary -- per batch row.  Restates, in one
// pass over the row:
//   SuppressTokensLogitsProcessor -> scores[suppress] = -inf
//   WhisperTimeStampLogitsProcessor (HF:generation/logits_process.py:1996-2043): no_timestamps = -inf; timestamps come
//   in pairs; no decreasing timestamps; first token must be a timestamp; "if the probability mass over timestamps
//   exceeds the most likely text token, sample a timestamp"
//   DiCoW: the EOS score survives at the first generated position (src/models/dicow/utils.py:10-12)
//   argmax (lowest index on ties), finished rows emit pad, unfinished &= token != eos (generation.py:756-779)
// With a cluster the per-slice results (best text token, best timestamp token, online logsumexp of the timestamp region)
// are written to rank 0's shared memory (DSMEM) and merged there in rank order.
struct RulesPartial {
  float text_v;
  int text_i;
  float ts_v;
  int ts_i;
  float tmax, tsum;
};

__global__ void __launch_bounds__(512) logits_rules_argmax_kernel(const RulesParams p, const int csplit) {
  __shared__ float s_max[16], s_sum[16];
  __shared__ Best s_text[16], s_ts[16];
  __shared__ RulesPartial s_part[8];  // slices of cluster ranks 1..7 (on rank 0)
  __shared__ int s_text_off;
  __shared__ unsigned long long s_last_ts;
  const int b = (int)blockIdx.x / csplit;
  const int rank = csplit > 1 ? (int)cluster_ctarank() : 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  if (csplit > 1) cluster_arrive_relaxed();  // "this CTA runs": waited for before the first store into rank 0's shared memory
  griddep_launch();
  griddep_wait();
  const int len = p.pos != nullptr ? (__ldcg(p.pos) + 1) : p.cur_len;  // tokens in the sequence so far
  long long* row_ids = p.ids + (long long)b * p.ids_rs;
  const int ngen = len - p.begin_index;
  const bool rules = p.ts_rules != 0;
  const bool at_begin = rules && ngen == 0;
  const long long last = ngen >= 1 ? __ldcg(row_ids + len - 1) : -1;
  const long long penult = ngen >= 2 ? __ldcg(row_ids + len - 2) : -1;
  const bool last_was_ts = rules && ngen >= 1 && last >= p.ts_begin;
  const bool penult_was_ts = ngen < 2 || penult >= p.ts_begin;
  // last timestamp token of the generated part: every thread looks at its own positions and the latest one wins through
  // a shared-memory maximum of (position, token).  (Every thread walking back from the end was a chain of dependent L2
  // reads as long as the text since the last timestamp -- the whole sequence when there is none.)
  if (tid == 0) s_last_ts = 0ull;
  __syncthreads();
  for (int i = p.begin_index + tid; rules && i < len; i += blockDim.x) {
    const long long tk = __ldcg(row_ids + i);
    if (tk >= p.ts_begin) atomicMax(&s_last_ts, ((unsigned long long)(i + 1) << 32) | (unsigned long long)(unsigned)tk);
  }
  __syncthreads();
  const int ts_last = s_last_ts != 0ull ? (int)(unsigned)(s_last_ts & 0xffffffffull) : -1;
  int ts_floor = p.ts_begin;  // timestamp ids below this are forbidden
  if (ts_last >= 0) ts_floor = (last_was_ts && !penult_was_ts) ? ts_last : ts_last + 1;
  const bool ts_all_masked = last_was_ts && penult_was_ts;
  const bool text_lt_eos_masked = last_was_ts && !penult_was_ts;
  const int ts_cap = (at_begin && p.max_initial_ts >= 0) ? p.ts_begin + p.max_initial_ts : p.V;  // ids > cap forbidden

  const float* lg = p.logits + (long long)b * p.ld;
  float* proc = p.processed != nullptr ? p.processed + (long long)b * p.V : nullptr;
  const int chunk = (((p.V + csplit - 1) / csplit) + 31) & ~31;  // whole bitmap words per slice
  const int v_begin = rank * chunk, v_end = min(p.V, v_begin + chunk);
  Best bt{-INFINITY, p.V}, bs{-INFINITY, p.V};
  float tmax = -INFINITY, tsum = 0.f;  // online logsumexp over the timestamp region
  // RU scores (and their bitmap words) per thread are requested before any is looked at; the order in which a thread
  // folds its scores is unchanged
  constexpr int RU = 8;
  for (int v0 = v_begin + tid; v0 < v_end; v0 += RU * (int)blockDim.x) {
    float xs[RU];
    unsigned sw[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int v = v0 + u * (int)blockDim.x;
      xs[u] = v < v_end ? __ldcg(lg + v) : 0.f;
      sw[u] = (p.suppress != nullptr && v < v_end) ? __ldg(p.suppress + (v >> 5)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int v = v0 + u * (int)blockDim.x;
      if (v >= v_end) break;
      float x = xs[u];
      bool masked = ((sw[u] >> (v & 31)) & 1u) || (rules && v == p.no_timestamps);
      if (v < p.ts_begin) {
        masked = masked || at_begin || (text_lt_eos_masked && v < p.eos);
        if (masked) x = -INFINITY;
        bt = best_of(bt, Best{x, v});
      } else {
        masked = masked || ts_all_masked || v < ts_floor || v > ts_cap;
        if (masked) x = -INFINITY;
        bs = best_of(bs, Best{x, v});
        if (x > -INFINITY) {
          const float mn = fmaxf(tmax, x);
          tsum = tsum * __expf(tmax - mn) + __expf(x - mn);
          tmax = mn;
        }
      }
      if (proc != nullptr) proc[v] = x;
    }
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, tmax, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, tsum, o);
    const float mn = fmaxf(tmax, m2);
    const float c1 = tmax == -INFINITY ? 0.f : __expf(tmax - mn);
    const float c2 = m2 == -INFINITY ? 0.f : __expf(m2 - mn);
    tsum = tsum * c1 + s2 * c2;
    tmax = mn;
  }
  if (lane == 0) s_text[warp] = bt, s_ts[warp] = bs, s_max[warp] = tmax, s_sum[warp] = tsum;
  __syncthreads();
  if (csplit > 1) cluster_wait_all();  // every CTA of the cluster has started (compute-sanitizer: a store into the shared
                                       // memory of a CTA that "might not have entered yet" otherwise)
  if (tid == 0) {
    for (int w = 1; w < nw; ++w) {
      bt = best_of(bt, s_text[w]);
      bs = best_of(bs, s_ts[w]);
      const float mn = fmaxf(tmax, s_max[w]);
      const float c1 = tmax == -INFINITY ? 0.f : __expf(tmax - mn);
      const float c2 = s_max[w] == -INFINITY ? 0.f : __expf(s_max[w] - mn);
      tsum = tsum * c1 + s_sum[w] * c2;
      tmax = mn;
    }
    if (rank != 0) {  // hand the slice to rank 0
      const uint32_t dst = map_to_cta(smem_u32(&s_part[rank]), 0);
      st_cluster_f32(dst + 0, bt.v), st_cluster_s32(dst + 4, bt.i);
      st_cluster_f32(dst + 8, bs.v), st_cluster_s32(dst + 12, bs.i);
      st_cluster_f32(dst + 16, tmax), st_cluster_f32(dst + 20, tsum);
    }
  }
  if (csplit > 1) cluster_sync_all();
  if (tid == 0 && rank == 0) {
    for (int r = 1; r < csplit; ++r) {
      const RulesPartial q = s_part[r];
      bt = best_of(bt, Best{q.text_v, q.text_i});
      bs = best_of(bs, Best{q.ts_v, q.ts_i});
      const float mn = fmaxf(tmax, q.tmax);
      const float c1 = tmax == -INFINITY ? 0.f : __expf(tmax - mn);
      const float c2 = q.tmax == -INFINITY ? 0.f : __expf(q.tmax - mn);
      tsum = tsum * c1 + q.tsum * c2;
      tmax = mn;
    }
    // logsumexp(timestamp log-probs) > max(text log-probs)  <=>  lse(ts scores) > max(text scores)
    const float ts_lse = tsum > 0.f ? tmax + __logf(tsum) : -INFINITY;
    const bool text_off = rules && ts_lse > bt.v;
    Best win = text_off ? bs : best_of(bt, bs);
    if (at_begin) {  // DiCoW: EOS keeps the score it had before the timestamp rules
      float e = __ldcg(lg + p.eos);
      if ((p.suppress != nullptr && ((p.suppress[p.eos >> 5] >> (p.eos & 31)) & 1u)) || p.eos == p.no_timestamps)
        e = -INFINITY;
      win = best_of(win, Best{e, p.eos});
    }
    if (win.i >= p.V) win.i = p.eos;  // every score -inf: cannot happen with finite logits; stay defined
    long long tok = win.i;
    const int unf = __ldcg(p.unfinished + b);
    if (!unf) tok = p.pad;
    if (!p.no_select) {
      row_ids[len] = tok;
      p.unfinished[b] = unf && (tok != p.eos);
    }
    s_text_off = text_off ? 1 : 0;
  }
  if (proc != nullptr) {  // materialise the processed scores exactly like the reference processors (tests only; csplit = 1)
    __syncthreads();
    if (s_text_off)
      for (int v = tid; v < p.ts_begin; v += blockDim.x) proc[v] = -INFINITY;
    __syncthreads();
    if (tid == 0 && at_begin) {
      float e = lg[p.eos];
      if ((p.suppress != nullptr && ((p.suppress[p.eos >> 5] >> (p.eos & 31)) & 1u)) || p.eos == p.no_timestamps)
        e = -INFINITY;
      proc[p.eos] = e;
    }
  }
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_gemm_skinny_bf16(dicow_handle_t h, const dicow_gemm_skinny_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_gemm_skinny_args_t), "dicow_gemm_skinny_bf16: bad args struct");
  DICOW_REQUIRE(ctx, a->A && a->W && a->out, "dicow_gemm_skinny_bf16: null operand");
  DICOW_REQUIRE(ctx, a->M >= 1 && a->M <= 64 && a->N >= 1 && a->K >= 32 && (a->K % 32) == 0,
                "dicow_gemm_skinny_bf16: need 1 <= M <= 64, K %% 32 == 0 (got M=%d N=%d K=%d)", a->M, a->N, a->K);
  DICOW_REQUIRE(ctx, (a->lda % 8) == 0 && (a->ldw % 8) == 0 && (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->W) % 16) == 0,
                "dicow_gemm_skinny_bf16: A / W rows must be 16-byte aligned");
  DICOW_REQUIRE(ctx, a->epilogue >= DICOW_EPI_BIAS_BF16 && a->epilogue <= DICOW_EPI_BIAS_F32,
                "dicow_gemm_skinny_bf16: unsupported epilogue %d", a->epilogue);
  if (a->epilogue == DICOW_EPI_RESIDUAL_F32) DICOW_REQUIRE(ctx, a->resid != nullptr, "dicow_gemm_skinny_bf16: resid is NULL");
  SkinnyParams p{};
  p.A = reinterpret_cast<const __nv_bfloat16*>(a->A), p.lda = a->lda;
  p.W = reinterpret_cast<const __nv_bfloat16*>(a->W), p.ldw = a->ldw;
  p.M = a->M, p.N = a->N, p.K = a->K, p.bias = a->bias, p.out = a->out, p.ldo = a->ldo, p.epilogue = a->epilogue;
  p.resid = a->resid, p.ldr = a->ldr, p.pos = a->pos, p.pos_stride = a->pos_stride;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = ceil_div(a->N, 8);
  switch (ceil_div(a->M, 16)) {
    case 1: gemm_skinny_kernel<1><<<grid, SK_THREADS, 0, stream>>>(p); break;
    case 2: gemm_skinny_kernel<2><<<grid, SK_THREADS, 0, stream>>>(p); break;
    case 3: gemm_skinny_kernel<3><<<grid, SK_THREADS, 0, stream>>>(p); break;
    default: gemm_skinny_kernel<4><<<grid, SK_THREADS, 0, stream>>>(p); break;
  }
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_decode_linear(dicow_handle_t h, const dicow_decode_linear_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_decode_linear_args_t), "dicow_decode_linear: bad args struct");
  const bool ln = a->x != nullptr;
  DICOW_REQUIRE(ctx, (ln || a->A) && a->W && a->out, "dicow_decode_linear: null operand");
  DICOW_REQUIRE(ctx, a->M >= 1 && a->M <= 64 && a->N >= 1 && a->K >= 32 && (a->K % 32) == 0,
                "dicow_decode_linear: need 1 <= M <= 64, K %% 32 == 0 (got M=%d N=%d K=%d)", a->M, a->N, a->K);
  DICOW_REQUIRE(ctx, (a->ldw % 8) == 0 && (reinterpret_cast<uintptr_t>(a->W) % 16) == 0, "dicow_decode_linear: W rows must be 16-byte aligned");
  if (ln)
    DICOW_REQUIRE(ctx, a->gamma && a->beta && a->K <= 1280 && (a->K % 4) == 0 && (a->ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(a->x) % 16) == 0 &&
                           (reinterpret_cast<uintptr_t>(a->gamma) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->beta) % 16) == 0,
                  "dicow_decode_linear: the LayerNorm prologue needs gamma/beta, K <= 1280 and 16-byte aligned fp32 rows");
  else
    DICOW_REQUIRE(ctx, (a->lda % 8) == 0 && (reinterpret_cast<uintptr_t>(a->A) % 16) == 0, "dicow_decode_linear: A rows must be 16-byte aligned");
  DICOW_REQUIRE(ctx, a->epilogue >= DICOW_EPI_BIAS_BF16 && a->epilogue <= DICOW_EPI_BIAS_F32, "dicow_decode_linear: unsupported epilogue %d", a->epilogue);
  if (a->epilogue == DICOW_EPI_RESIDUAL_F32) DICOW_REQUIRE(ctx, a->resid != nullptr, "dicow_decode_linear: resid is NULL");
  if (a->out2 != nullptr)
    DICOW_REQUIRE(ctx, a->n_split > 0 && a->n_split < a->N && (a->n_split % 8) == 0, "dicow_decode_linear: bad n_split %d", a->n_split);
  const int tiles = ceil_div(a->N, 8);
  // more than 32 rows (beam search: utterances x beams hypotheses) are split over two CTAs of 32 rows per column tile:
  // a 64-row A slab is 168 KB of shared memory = one CTA per SM and too few weight loads in flight (measured 22.6 us
  // per layer at M = 60 against 8 us at M = 16); the second CTA's weight requests are L2 hits
  const int msplit = a->M > 32 ? ceil_div(a->M, 32) : 1;
  const int MT = msplit > 1 ? 2 : ceil_div(a->M, 16);
  // split K over a cluster until a CTA's slice fits the per-warp register slab (<= 1280), then once more while the grid
  // would leave half of the SMs idle; slices stay multiples of 32.  The two-output form keeps K whole (fused q|k,v: N = 3d).
  int ksplit = 1;
  while ((a->K / ksplit > DL_MAX_KS || (a->K % ksplit) != 0 || ((a->K / ksplit) % 32) != 0) && ksplit < DL_MAX_KSPLIT) ++ksplit;
  DICOW_REQUIRE(ctx, a->K / ksplit <= DL_MAX_KS && (a->K % ksplit) == 0 && ((a->K / ksplit) % 32) == 0,
                "dicow_decode_linear: K = %d cannot be split into <= 8 slices of <= 1280 (multiples of 32)", a->K);
  if (a->out2 != nullptr) DICOW_REQUIRE(ctx, ksplit == 1, "dicow_decode_linear: two outputs need K <= 1280");
  if (a->out2 == nullptr && tiles * ksplit < 2 * ctx->num_sms && 2 * ksplit <= DL_MAX_KSPLIT && a->K / ksplit >= 640 &&
      ((a->K / (2 * ksplit)) % 32) == 0)
    ksplit *= 2;
  DecLinParams p{};
  p.x = a->x, p.ldx = a->ldx, p.gamma = a->gamma, p.beta = a->beta, p.eps = a->eps;
  p.A = reinterpret_cast<const __nv_bfloat16*>(a->A), p.lda = a->lda;
  p.W = reinterpret_cast<const __nv_bfloat16*>(a->W), p.ldw = a->ldw;
  p.M = a->M, p.N = a->N, p.K = a->K, p.bias = a->bias, p.out = a->out, p.ldo = a->ldo, p.epilogue = a->epilogue;
  p.resid = a->resid, p.ldr = a->ldr;
  p.n_split = a->out2 != nullptr ? a->n_split : a->N, p.out2 = a->out2, p.ldo2 = a->ldo2, p.pos = a->pos, p.pos_stride = a->pos_stride;
  p.ksplit = ksplit, p.Ks = a->K / ksplit, p.a_pitch = p.Ks + 32;
  // tiles per pass: as many (<= 3; DICOW_DL_NT=1..4 overrides the cap) as still leave about one CTA per SM -- the layers'
  // linears of a step; proj_out (thousands of tiles) keeps one tile per pass and loops
  static const int nt_cap = [] {
    const char* e = getenv("DICOW_DL_NT");
    return e != nullptr && e[0] >= '1' && e[0] <= '4' ? e[0] - '0' : 3;
  }();
  // (beam search, M > 32: two 32-row CTAs per tile group.  One tile per pass staged an 84 KB slab per 20 KB of weights --
  // 1 280 CTAs for fc1 at 60 hypotheses, 107 MB of L2 -> SM traffic for 13 MB of weights)
  int NT = 1;
  for (int c = nt_cap; c >= 2; --c) {
    const int ctas = ceil_div(tiles, c) * ksplit * msplit;
    if (ctas * 10 >= ctx->num_sms * 9 && ctas <= (msplit > 1 ? 3 : 2) * ctx->num_sms) {
      NT = c;
      break;
    }
  }
  // 8 warps.  16 (3 k-blocks per warp instead of 5) measured slower for every single-pass shape: fc1 7.4 -> 9.3 us, the
  // fused q|k,v projection 6.1 -> 8.1 us, the decode step 0.318 -> 0.344 ms (gpurun_out/s3_decode_nt.json).
  constexpr int NW = 8;
  const size_t smem = (size_t)MT * 16 * p.a_pitch * 2 + (size_t)(NW + ksplit - 1) * NT * MT * 16 * 8 * sizeof(float) +
                      (ln ? 2 * (size_t)a->K * sizeof(float) : 0);
  DICOW_REQUIRE(ctx, smem <= (size_t)ctx->max_smem_optin, "dicow_decode_linear: %zu bytes of shared memory needed", smem);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  auto go = [&](auto kern) -> int {
    DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (the carveout stays at the driver's default: forcing cudaSharedmemCarveoutMaxShared to fit four CTAs per SM made
    // every instance SLOWER -- proj_out 32 -> 48 us, fc2 12 -> 16 us -- the weight stream needs the L1 array for its
    // loads in flight more than it needs residency)
    // one wave: a CTA takes as many consecutive tiles as it needs for the grid to fit the resident CTA slots
    int per_sm = 1;
    DICOW_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
    per_sm = per_sm < 1 ? 1 : per_sm;
    p.tiles_per_cta = NT > 1 ? NT : ksplit > 1 ? 1 : ceil_div(tiles * msplit, per_sm * ctx->num_sms);
    const int grid = ceil_div(tiles, p.tiles_per_cta) * ksplit;
    const unsigned cluster = (unsigned)ksplit;
    DICOW_CUDA_OK(ctx, launch_step_kernel(kern, dim3(grid, msplit), dim3(NW * 32), smem, stream, cluster, p));
    return DICOW_OK;
  };
  switch (NT * 16 + MT * 2 + (ln ? 1 : 0)) {
    case 16 + 2: return go(decode_linear_kernel<1, false, 1, NW>);
    case 16 + 3: return go(decode_linear_kernel<1, true, 1, NW>);
    case 16 + 4: return go(decode_linear_kernel<2, false, 1, NW>);
    case 16 + 5: return go(decode_linear_kernel<2, true, 1, NW>);
    case 32 + 2: return go(decode_linear_kernel<1, false, 2, NW>);
    case 32 + 3: return go(decode_linear_kernel<1, true, 2, NW>);
    case 32 + 4: return go(decode_linear_kernel<2, false, 2, NW>);
    case 32 + 5: return go(decode_linear_kernel<2, true, 2, NW>);
    case 48 + 2: return go(decode_linear_kernel<1, false, 3, NW>);
    case 48 + 3: return go(decode_linear_kernel<1, true, 3, NW>);
    case 48 + 4: return go(decode_linear_kernel<2, false, 3, NW>);
    case 48 + 5: return go(decode_linear_kernel<2, true, 3, NW>);
    case 64 + 2: return go(decode_linear_kernel<1, false, 4, NW>);
    case 64 + 3: return go(decode_linear_kernel<1, true, 4, NW>);
    case 64 + 4: return go(decode_linear_kernel<2, false, 4, NW>);
    case 64 + 5: return go(decode_linear_kernel<2, true, 4, NW>);
    default: DICOW_REQUIRE(ctx, false, "dicow_decode_linear: no kernel for MT=%d NT=%d NW=%d ln=%d", MT, NT, NW, (int)ln);
  }
  return DICOW_OK;
}

extern "C" int dicow_decode_attention_bf16(dicow_handle_t h, const dicow_decode_attention_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_decode_attention_args_t),
                "dicow_decode_attention_bf16: bad args struct");
  DICOW_REQUIRE(ctx, a->Q && a->K && a->V && a->out, "dicow_decode_attention_bf16: null operand");
  DICOW_REQUIRE(ctx, a->B >= 1 && a->B <= 65535 && a->H >= 1 && (a->Tk >= 1 || a->pos != nullptr),
                "dicow_decode_attention_bf16: bad shape B=%d H=%d Tk=%d", a->B, a->H, a->Tk);
  DICOW_REQUIRE(ctx, (a->q_batch_stride % 8) == 0 && (a->kv_row_stride % 8) == 0 && (a->kv_batch_stride % 8) == 0 &&
                         ((reinterpret_cast<uintptr_t>(a->Q) | reinterpret_cast<uintptr_t>(a->K) |
                           reinterpret_cast<uintptr_t>(a->V)) % 16) == 0,
                "dicow_decode_attention_bf16: strides must be multiples of 8 elements, bases 16-byte aligned");
  DecAttnParams p{};
  p.Q = reinterpret_cast<const __nv_bfloat16*>(a->Q), p.q_bs = a->q_batch_stride;
  p.K = reinterpret_cast<const __nv_bfloat16*>(a->K), p.V = reinterpret_cast<const __nv_bfloat16*>(a->V);
  p.kv_rs = a->kv_row_stride, p.kv_bs = a->kv_batch_stride, p.kv_hs = a->kv_head_stride > 0 ? a->kv_head_stride : 64;
  DICOW_REQUIRE(ctx, (p.kv_hs % 8) == 0, "dicow_decode_attention_bf16: kv_head_stride must be a multiple of 8 elements");
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out), p.o_bs = a->o_batch_stride, p.Tk = a->Tk, p.pos = a->pos;
  p.kv_batch_div = a->kv_batch_div > 1 ? a->kv_batch_div : 1;
  p.ancestry = a->ancestry, p.anc_stride = a->ancestry_stride;
  dim3 grid(a->H, a->B);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // long fixed-length caches (cross attention, Tk = 1500) get 8 warps per (batch, head), the growing self-attention
  // cache (Tk <= 448, read from the device scalar) 4
  // Measured (60 hypotheses = 12 utterances x 5 beams, tools/profile_decode.py): 92.8 us per layer against 73.9 us for the
  // single-query kernel -- NQ = 5 needs 150 registers (one CTA per SM) and leaves 240 CTAs for 148 SMs; the re-reads it
  // saves are L2 hits anyway.  Opt-in (DICOW_MQ_ATTENTION=1) until it splits the keys over more CTAs.
  static const bool mq_enabled = [] {
    const char* e = getenv("DICOW_MQ_ATTENTION");
    return e != nullptr && e[0] == '1';
  }();
  if (mq_enabled && a->pos == nullptr && a->ancestry == nullptr && p.kv_batch_div > 1 && p.kv_batch_div <= 8 &&
      (a->B % p.kv_batch_div) == 0) {
    // beam search cross attention: the beams of an utterance share one K/V stream -> one CTA per (head, utterance)
    const dim3 g2(a->H, a->B / p.kv_batch_div);
    switch (p.kv_batch_div) {
#define DICOW_MQ(N) \
  case N: DICOW_CUDA_OK(ctx, launch_step_kernel(decode_attention_mq_kernel<N>, g2, dim3(256), 0, stream, 1, p)); break;
      DICOW_MQ(2) DICOW_MQ(3) DICOW_MQ(4) DICOW_MQ(5) DICOW_MQ(6) DICOW_MQ(7) DICOW_MQ(8)
#undef DICOW_MQ
      default: break;
    }
    return DICOW_OK;
  }
  // the head-major cross-attention cache: k | v of a key are 256 contiguous bytes, keys contiguous -> shared-memory ring kernel
  static const bool ring_enabled = [] {
    const char* e = getenv("DICOW_DECODE_ATTN_RING");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool ring_ok = ring_enabled && a->pos == nullptr && a->ancestry == nullptr && a->Tk >= 512 && a->kv_row_stride == 128 &&
                       reinterpret_cast<const __nv_bfloat16*>(a->V) == reinterpret_cast<const __nv_bfloat16*>(a->K) + 64 &&
                       (p.kv_hs % 8) == 0 && (p.kv_bs % 8) == 0;
  if (ring_ok) {
    constexpr size_t smem = (size_t)CX_STAGES * CX_KEYS * 256;
    static DeviceOnce attr_once;
    if (attr_once.first(ctx))
      DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(decode_cross_attention_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DICOW_CUDA_OK(ctx, launch_step_kernel(decode_cross_attention_ring_kernel, grid, dim3((CX_WARPS + 1) * 32), smem, stream, 1, p));
  } else if (a->pos == nullptr && a->Tk >= 512)
    DICOW_CUDA_OK(ctx, launch_step_kernel(decode_attention_kernel<8>, grid, dim3(8 * 32), 0, stream, 1, p));
  else
    DICOW_CUDA_OK(ctx, launch_step_kernel(decode_attention_kernel<4>, grid, dim3(4 * 32), 0, stream, 1, p));
  return DICOW_OK;
}

extern "C" int dicow_kv_to_head_major(dicow_handle_t h, const void* kv_bf16, void* out_bf16, int B, int T, int H, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, kv_bf16 && out_bf16 && B >= 1 && B <= 65535 && T >= 1 && H >= 1 &&
                         ((reinterpret_cast<uintptr_t>(kv_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) % 16) == 0,
                "dicow_kv_to_head_major: bad args");
  kv_to_head_major_kernel<<<dim3(T, B), 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const uint4*>(kv_bf16), reinterpret_cast<uint4*>(out_bf16), T, H);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}

extern "C" int dicow_embed_tokens(dicow_handle_t h, const int64_t* ids, int64_t ids_row_stride, const float* embed_tokens,
                                  const float* embed_positions, float* x, int B, int S, int d, int vocab, int past,
                                  const int32_t* pos, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, ids && embed_tokens && embed_positions && x && B >= 1 && B <= 65535 && S >= 1 && d >= 4 && (d % 4) == 0,
                "dicow_embed_tokens: bad args");
  dim3 grid(S, B);
  DICOW_CUDA_OK(ctx, launch_step_kernel(embed_kernel, grid, dim3(128), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
                                        reinterpret_cast<const long long*>(ids), (long long)ids_row_stride, embed_tokens,
                                        embed_positions, x, S, d, past, pos, vocab));
  return DICOW_OK;
}

extern "C" int dicow_advance(dicow_handle_t h, int32_t* pos, int by, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, pos != nullptr, "dicow_advance: null pos");
  DICOW_CUDA_OK(ctx, launch_step_kernel(advance_kernel, dim3(1), dim3(1), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
                                        reinterpret_cast<int*>(pos), by));
  return DICOW_OK;
}

extern "C" int dicow_logits_rules_argmax(dicow_handle_t h, const dicow_logits_rules_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_logits_rules_args_t),
                "dicow_logits_rules_argmax: bad args struct");
  DICOW_REQUIRE(ctx, a->logits && a->ids && a->unfinished && a->B >= 1 && a->V >= 2, "dicow_logits_rules_argmax: bad args");
  DICOW_REQUIRE(ctx, a->eos >= 0 && a->eos < a->V && a->ts_begin > a->eos && a->ts_begin <= a->V && a->begin_index >= 0,
                "dicow_logits_rules_argmax: need 0 <= eos < ts_begin <= V");
  RulesParams p{};
  p.logits = a->logits, p.ld = a->ld, p.V = a->V;
  p.ids = reinterpret_cast<long long*>(a->ids), p.ids_rs = a->ids_row_stride, p.pos = a->pos, p.cur_len = a->cur_len;
  p.begin_index = a->begin_index, p.eos = a->eos, p.pad = a->pad, p.no_timestamps = a->no_timestamps;
  p.ts_begin = a->ts_begin, p.max_initial_ts = a->max_initial_timestamp_index, p.ts_rules = a->timestamp_rules;
  p.suppress = a->suppress_bitmap, p.unfinished = a->unfinished, p.processed = a->processed_scores;
  p.no_select = a->no_select;
  DICOW_REQUIRE(ctx, !a->no_select || a->processed_scores != nullptr, "dicow_logits_rules_argmax: no_select needs processed_scores");
  // the vocabulary of one row is scanned by a cluster of 8 CTAs (B rows alone would occupy B of the 148 SMs); the
  // test-only materialisation of the processed scores needs the row's verdict in every slice and runs unsplit
  const int csplit = (a->processed_scores == nullptr && a->V >= 8192) ? 8 : 1;
  DICOW_CUDA_OK(ctx, launch_step_kernel(logits_rules_argmax_kernel, dim3(a->B * csplit), dim3(512), 0,
                                        reinterpret_cast<cudaStream_t>(stream_), (unsigned)csplit, p, csplit));
  return DICOW_OK;
}
