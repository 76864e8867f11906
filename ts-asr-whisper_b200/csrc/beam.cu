// beam.cu -- one step of beam search bookkeeping on the device.
//
// Replaces, per generated token, the tensor bookkeeping of DiCoWGenerationMixin._beam_search
// (src/models/dicow/generation.py:992-1107; = HF GenerationMixin._beam_search: _get_top_k_continuations,
// _get_running_beams_for_next_iteration, _update_finished_beams, _check_early_stop_heuristic) and the two state moves
// that follow it: past_key_values.reorder_cache(beam_idx) and ctc_rescorer.update_state(tokens, beam_idx)
// (generation.py:1080-1088).  ~40 eager tensor ops on [batch, 2 x beams] arrays plus a top-k over beams x 51 866 scores
// per utterance in the reference; here:
//
//   beam_select_kernel   one CTA per utterance.  The candidates of a beam are the K text ids the joint CTC step scored
//                        (every other text id has K better-scored ids of the same beam in front of it, so it can never
//                        be among the top 2 x beams continuations) plus all timestamp ids.  The 2 x beams best
//                        continuations are extracted in order by 2 x beams rounds of a block-wide arg-max over
//                        (score, flat index) -- exact, deterministic tie-breaking (lower beam * V + token first), no
//                        per-thread lists.  One thread then restates the bookkeeping (hits, next running beams,
//                        finished set with length penalty, early-stop heuristic); all threads write the re-linked token
//                        sequences into scratch.
//   beam_gather_kernel   per new hypothesis: ancestry row of its parent + its own slot, CTC forward variables of the
//                        parent or of the chosen candidate -> scratch
//   beam_commit_kernel   scratch -> state (the hypotheses of an utterance read each other's rows, so the move is
//                        two-phase)
//
// The K/V cache itself is never re-ordered: decode_attention follows the ancestry table (decode.cu).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace dicow {
namespace {

constexpr float BEAM_NEG = -1.0e9f;
constexpr float BEAM_LOGZERO = -1e10f;
constexpr int BS_THREADS = 256;
constexpr int MAX_K2 = 16;  // 2 x beams

struct BeamParams {
  int U, NB, V, K, T;
  const float* proc;
  const int* cs;
  const float* lse;
  const float* att;
  const float* psi;
  float w;
  const float* states;
  float* r_prev;
  float* score_prev;
  float* r_tmp;
  float* run_score;
  float* fin_score;
  int* fin_flag;
  int* unsat;
  long long* ids;
  long long* fin_ids;
  long long* ids_tmp;  // [2 R, ids_rs]: new running rows, then new finished rows
  long long ids_rs;
  int* anc;
  int* anc_tmp;
  long long anc_rs;
  const int* pos;
  int eos, pad, first_ts, max_length, prompt_len;
  float lp;
  int early;
  int* parent;  // scratch_i32: [R] parent row | [R] candidate slot (-1: timestamp) | [R] token
  int* slot;
  int* tok;
  float* new_run;    // scratch_f32: [R] | [R] new ctc score
  float* new_score;
  int* flags;
};

struct Item {
  float s;
  int flat;  // beam * V + token
  int slot;  // candidate slot or -1
};
__device__ __forceinline__ bool item_before(const Item& a, const Item& b) {  // a ranks before b
  return a.s > b.s || (a.s == b.s && a.flat < b.flat);
}

__global__ void __launch_bounds__(BS_THREADS) beam_select_kernel(const BeamParams p) {
  __shared__ float s_maxpsi[8], s_run[8], s_prev[8], s_lse[8];
  __shared__ Item s_red[BS_THREADS / 32];
  __shared__ Item s_top[MAX_K2];
  __shared__ int s_run_src[8];                 // index into s_top of the continuation each new running beam takes
  __shared__ int s_fin_src[8];                 // new finished slot i <- old finished slot (0..NB-1) or NB + continuation j
  const int u = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = p.NB, K2 = 2 * NB, R0 = u * NB;
  const int cur_len = *p.pos + 1;  // tokens in every running sequence
  // ---- per-beam constants ----
  if (warp < NB) {
    const int r = R0 + warp;
    float mx = BEAM_LOGZERO;
    if (p.w > 0.f)
      for (int j = lane; j < p.K; j += 32) mx = fmaxf(mx, p.psi[(long long)r * p.K + j]);
    mx = warp_max(mx);
    if (lane == 0) {
      s_maxpsi[warp] = mx, s_run[warp] = p.run_score[r], s_lse[warp] = p.lse[r];
      s_prev[warp] = p.w > 0.f ? p.score_prev[r] : 0.f;
    }
  }
  __syncthreads();
  // ---- the 2 x NB best continuations, in order ----
  // Scores of all candidates of the utterance are evaluated once into shared memory; every round is then a scan of that
  // array for the best item strictly after the previous winner in (score desc, flat index asc) order.
  extern __shared__ float s_sc[];
  const int n_ts = p.V - p.first_ts;
  const int per_beam = p.K + n_ts;
  const int n_items = NB * per_beam;
  auto flat_of = [&](int it) {
    const int i = it / per_beam, j = it - i * per_beam;
    return i * p.V + (j < p.K ? p.cs[(long long)(R0 + i) * p.K + j] : p.first_ts + (j - p.K));
  };
  for (int it = tid; it < n_items; it += BS_THREADS) {
    const int i = it / per_beam, j = it - i * per_beam;
    const int r = R0 + i;
    float sc = -INFINITY;
    if (j < p.K) {
      const float a = p.att[(long long)r * p.K + j];
      if (a > -INFINITY) {
        sc = a;
        if (p.w > 0.f) sc = (1.f - p.w) * a + p.w * (p.psi[(long long)r * p.K + j] - s_prev[i]);
        sc += s_run[i];
      }
    } else {
      const float x = p.proc[(long long)r * p.V + p.first_ts + (j - p.K)];
      if (x > -INFINITY) {
        const float a = x - s_lse[i];
        sc = a;
        if (p.w > 0.f) sc = (1.f - p.w) * a + p.w * (s_maxpsi[i] - s_prev[i]);
        sc += s_run[i];
      }
    }
    s_sc[it] = sc;
  }
  __syncthreads();
  float last_s = INFINITY;
  int last_flat = -1;
  for (int round = 0; round < K2; ++round) {
    float bs = -INFINITY;
    int bit = -1, bflat = 0x7fffffff;
    for (int it = tid; it < n_items; it += BS_THREADS) {
      const float sc = s_sc[it];
      if (!(sc > -INFINITY) || sc > last_s || sc < bs) continue;
      int fl = -1;
      if (sc == last_s) {  // a tie with the previous winner: only items after it in flat order remain
        fl = flat_of(it);
        if (fl <= last_flat) continue;
      }
      if (sc == bs) {
        if (fl < 0) fl = flat_of(it);
        if (bflat == 0x7fffffff && bit >= 0) bflat = flat_of(bit);
        if (fl >= bflat) continue;
        bflat = fl;
      } else {
        bflat = fl >= 0 ? fl : 0x7fffffff;
      }
      bs = sc, bit = it;
    }
    Item best{bs, bit >= 0 ? (bflat != 0x7fffffff ? bflat : flat_of(bit)) : 0x7fffffff, bit};  // slot field carries the item index here
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Item q;
      q.s = __shfl_xor_sync(0xffffffffu, best.s, o);
      q.flat = __shfl_xor_sync(0xffffffffu, best.flat, o);
      q.slot = __shfl_xor_sync(0xffffffffu, best.slot, o);
      if (item_before(q, best)) best = q;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < BS_THREADS / 32; ++w)
        if (item_before(s_red[w], best)) best = s_red[w];
      if (best.slot >= 0) {  // item index -> candidate slot (-1 for a timestamp id)
        const int j = best.slot % per_beam;
        best.slot = j < p.K ? j : -1;
      }
      s_top[round] = best;
    }
    __syncthreads();
    last_s = s_top[round].s, last_flat = s_top[round].flat;
  }
  // ---- bookkeeping (generation.py:1023-1105) ----
  if (tid == 0) {
    bool hit[MAX_K2];
    float runv[MAX_K2];
    int any_nohit = 0;
    for (int j = 0; j < K2; ++j) {
      const Item c = s_top[j];
      const bool valid = c.flat != 0x7fffffff;
      const int t = valid ? c.flat % p.V : p.pad;
      hit[j] = (t == p.eos) || (cur_len + 1 >= p.max_length) || !valid;
      runv[j] = (valid ? c.s : -INFINITY) + (hit[j] ? BEAM_NEG : 0.f);
      any_nohit |= hit[j] ? 0 : 1;
    }
    // next running beams: the NB largest of runv, ties by position
    bool used[MAX_K2] = {};
    for (int i = 0; i < NB; ++i) {
      int bj = -1;
      for (int j = 0; j < K2; ++j)
        if (!used[j] && (bj < 0 || runv[j] > runv[bj])) bj = j;
      used[bj] = true;
      s_run_src[i] = bj;
      const Item c = s_top[bj];
      const bool valid = c.flat != 0x7fffffff;
      p.parent[R0 + i] = R0 + (valid ? c.flat / p.V : 0);
      p.tok[R0 + i] = valid ? c.flat % p.V : p.pad;
      p.slot[R0 + i] = valid ? c.slot : -1;
      p.new_run[R0 + i] = runv[bj];
    }
    // finished set: old NB entries + the 2 NB continuations (only hits among the first NB count)
    const float L = (float)(cur_len + 1 - p.prompt_len);
    const float denom = powf(L, p.lp);
    bool all_full = true;
    for (int i = 0; i < NB; ++i) all_full = all_full && p.fin_flag[R0 + i] != 0;
    const bool full = all_full && p.early == 1;
    const bool unsat = p.unsat[u] != 0;
    float ms[8 + MAX_K2];
    bool mf[8 + MAX_K2];
    for (int i = 0; i < NB; ++i) ms[i] = p.fin_score[R0 + i], mf[i] = p.fin_flag[R0 + i] != 0;
    for (int j = 0; j < K2; ++j) {
      const bool did = hit[j] && j < NB && s_top[j].flat != 0x7fffffff;
      float fs = s_top[j].s / denom;
      fs += full ? BEAM_NEG : 0.f;
      fs += !unsat ? BEAM_NEG : 0.f;
      fs += !did ? BEAM_NEG : 0.f;
      ms[NB + j] = fs, mf[NB + j] = did;
    }
    bool taken[8 + MAX_K2] = {};
    float nfs[8];
    bool nff[8];
    for (int i = 0; i < NB; ++i) {
      int bi = -1;
      for (int k = 0; k < NB + K2; ++k)
        if (!taken[k] && (bi < 0 || ms[k] > ms[bi])) bi = k;
      taken[bi] = true;
      s_fin_src[i] = bi;
      nfs[i] = ms[bi], nff[i] = mf[bi];
    }
    bool nall = true;
    float worst = INFINITY;
    for (int i = 0; i < NB; ++i) {
      p.fin_score[R0 + i] = nfs[i], p.fin_flag[R0 + i] = nff[i] ? 1 : 0;
      nall = nall && nff[i];
      worst = fminf(worst, nfs[i]);
    }
    // early-stop heuristic with the incremented length (generation.py:1090-1100)
    const int Lh = (p.early == 2 && p.lp > 0.f) ? (p.max_length - p.prompt_len) : (cur_len + 1 - p.prompt_len);
    const float best_run = p.new_run[R0] / powf((float)Lh, p.lp);
    bool can = false;
    for (int i = 0; i < NB; ++i) can = can || best_run > (nff[i] ? worst : BEAM_NEG);
    const int nunsat = (unsat && can) ? 1 : 0;
    p.unsat[u] = nunsat;
    p.flags[4 * u + 0] = any_nohit, p.flags[4 * u + 1] = nall ? 1 : 0, p.flags[4 * u + 2] = nunsat;
  }
  __syncthreads();
  // ---- re-linked token sequences -> scratch (running rows, then finished rows) ----
  const int R = p.U * NB;
  for (int i = 0; i < NB; ++i) {
    const Item c = s_top[s_run_src[i]];
    const bool valid = c.flat != 0x7fffffff;
    const long long* src = p.ids + (long long)(R0 + (valid ? c.flat / p.V : 0)) * p.ids_rs;
    long long* dst = p.ids_tmp + (long long)(R0 + i) * p.ids_rs;
    for (int t = tid; t < p.ids_rs; t += BS_THREADS)
      dst[t] = t < cur_len ? src[t] : (t == cur_len ? (long long)(valid ? c.flat % p.V : p.pad) : 0);
    const int f = s_fin_src[i];
    long long* fdst = p.ids_tmp + (long long)(R + R0 + i) * p.ids_rs;
    if (f < NB) {
      const long long* fsrc = p.fin_ids + (long long)(R0 + f) * p.ids_rs;
      for (int t = tid; t < p.ids_rs; t += BS_THREADS) fdst[t] = fsrc[t];
    } else {
      const Item d = s_top[f - NB];
      const bool dv = d.flat != 0x7fffffff;
      const long long* fsrc = p.ids + (long long)(R0 + (dv ? d.flat / p.V : 0)) * p.ids_rs;
      for (int t = tid; t < p.ids_rs; t += BS_THREADS)
        fdst[t] = t < cur_len ? fsrc[t] : (t == cur_len ? (long long)(dv ? d.flat % p.V : p.pad) : (long long)p.pad);
    }
  }
}

// new hypothesis r (parent row, token, candidate slot): ancestry and CTC state -> scratch
__global__ void __launch_bounds__(128) beam_gather_kernel(const BeamParams p) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const int par = p.parent[r];
  const int cur = *p.pos;  // position of the token the step just consumed; the next one goes to cur + 1
  const int* asrc = p.anc + (long long)par * p.anc_rs;
  int* adst = p.anc_tmp + (long long)r * p.anc_rs;
  for (int t = tid; t < p.anc_rs; t += blockDim.x) adst[t] = t <= cur ? asrc[t] : r;  // own row from cur + 1 on
  if (p.w > 0.f) {
    const int slot = p.slot[r], tok = p.tok[r];
    float* dst = p.r_tmp + (long long)r * p.T * 2;
    if (tok < p.first_ts && slot >= 0) {  // update_state (decoding.py:253-260): the candidate's forward variables / score
      const float* st = p.states + (long long)par * p.T * 2 * p.K + slot;
      for (int i = tid; i < 2 * p.T; i += blockDim.x) dst[i] = st[(long long)i * p.K];
      if (tid == 0) p.new_score[r] = p.psi[(long long)par * p.K + slot];
    } else {
      const float* src = p.r_prev + (long long)par * p.T * 2;
      for (int i = tid; i < 2 * p.T; i += blockDim.x) dst[i] = src[i];
      if (tid == 0) p.new_score[r] = (tok < p.first_ts) ? BEAM_LOGZERO : p.score_prev[par];
    }
  }
}

__global__ void __launch_bounds__(128) beam_commit_kernel(const BeamParams p) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const int R = p.U * p.NB;
  long long* ids = p.ids + (long long)r * p.ids_rs;
  long long* fin = p.fin_ids + (long long)r * p.ids_rs;
  const long long* t0 = p.ids_tmp + (long long)r * p.ids_rs;
  const long long* t1 = p.ids_tmp + (long long)(R + r) * p.ids_rs;
  for (int t = tid; t < p.ids_rs; t += blockDim.x) ids[t] = t0[t], fin[t] = t1[t];
  int* a = p.anc + (long long)r * p.anc_rs;
  const int* at = p.anc_tmp + (long long)r * p.anc_rs;
  for (int t = tid; t < p.anc_rs; t += blockDim.x) a[t] = at[t];
  if (p.w > 0.f) {
    float* d = p.r_prev + (long long)r * p.T * 2;
    const float* s = p.r_tmp + (long long)r * p.T * 2;
    for (int i = tid; i < 2 * p.T; i += blockDim.x) d[i] = s[i];
    if (tid == 0) p.score_prev[r] = p.new_score[r];
  }
  if (tid == 0) p.run_score[r] = p.new_run[r];
}

}  // namespace
}  // namespace dicow

using namespace dicow;

extern "C" int dicow_beam_step(dicow_handle_t h, const dicow_beam_step_args_t* a, void* stream_) {
  if (h == nullptr) return DICOW_ERR_INVALID_ARG;
  dicow_ctx* ctx = h;
  DICOW_REQUIRE(ctx, a != nullptr && a->struct_size == sizeof(dicow_beam_step_args_t), "dicow_beam_step: bad args struct");
  DICOW_REQUIRE(ctx, a->U >= 1 && a->NB >= 1 && a->NB <= 8 && a->U * a->NB <= 64 && a->K >= 2 * a->NB && a->V > a->first_timestamp,
                "dicow_beam_step: need 1 <= beams <= 8, rows <= 64, K >= 2 beams (U=%d NB=%d K=%d)", a->U, a->NB, a->K);
  DICOW_REQUIRE(ctx, a->processed_scores && a->joint_workspace_i32 && a->joint_workspace_f32 && a->run_score && a->fin_score &&
                         a->fin_flag && a->unsat && a->ids && a->fin_ids && a->ids_tmp && a->ancestry && a->ancestry_tmp &&
                         a->pos && a->scratch_i32 && a->scratch_f32 && a->flags,
                "dicow_beam_step: null argument");
  if (a->ctc_weight > 0.f)
    DICOW_REQUIRE(ctx, a->ctc_states && a->ctc_r_prev && a->ctc_score_prev && a->ctc_r_tmp && a->T >= 2,
                  "dicow_beam_step: ctc_weight > 0 needs the CTC state buffers");
  const int R = a->U * a->NB;
  BeamParams p{};
  p.U = a->U, p.NB = a->NB, p.V = a->V, p.K = a->K, p.T = a->T;
  p.proc = a->processed_scores;
  p.cs = a->joint_workspace_i32 + 4 * R + 4;                      // layout of dicow_ctc_joint_step's workspaces
  p.lse = a->joint_workspace_f32;
  p.att = a->joint_workspace_f32 + R;
  p.psi = a->joint_workspace_f32 + R + (size_t)R * a->K;
  p.w = a->ctc_weight, p.states = a->ctc_states, p.r_prev = a->ctc_r_prev, p.score_prev = a->ctc_score_prev, p.r_tmp = a->ctc_r_tmp;
  p.run_score = a->run_score, p.fin_score = a->fin_score, p.fin_flag = a->fin_flag, p.unsat = a->unsat;
  p.ids = reinterpret_cast<long long*>(a->ids), p.fin_ids = reinterpret_cast<long long*>(a->fin_ids);
  p.ids_tmp = reinterpret_cast<long long*>(a->ids_tmp), p.ids_rs = a->ids_row_stride;
  p.anc = a->ancestry, p.anc_tmp = a->ancestry_tmp, p.anc_rs = a->ancestry_stride;
  p.pos = a->pos, p.eos = a->eos, p.pad = a->pad, p.first_ts = a->first_timestamp, p.max_length = a->max_length;
  p.prompt_len = a->prompt_len, p.lp = a->length_penalty, p.early = a->early_stopping;
  p.parent = a->scratch_i32, p.slot = a->scratch_i32 + R, p.tok = a->scratch_i32 + 2 * R;
  p.new_run = a->scratch_f32, p.new_score = a->scratch_f32 + R;
  p.flags = a->flags;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const size_t sel_smem = (size_t)a->NB * (a->K + (a->V - a->first_timestamp)) * sizeof(float);
  DICOW_REQUIRE(ctx, sel_smem <= (size_t)ctx->max_smem_optin - 4096, "dicow_beam_step: %zu bytes of candidate scores per utterance", sel_smem);
  DICOW_CUDA_OK(ctx, cudaFuncSetAttribute(beam_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  beam_select_kernel<<<a->U, BS_THREADS, sel_smem, stream>>>(p);
  beam_gather_kernel<<<R, 128, 0, stream>>>(p);
  beam_commit_kernel<<<R, 128, 0, stream>>>(p);
  DICOW_CUDA_OK(ctx, cudaGetLastError());
  return DICOW_OK;
}
