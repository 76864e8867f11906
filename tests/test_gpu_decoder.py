"""GPU parity of the decoder side (C-ABI kernels + DiCoWForConditionalGeneration) against the CPU oracle and the
committed reference outputs (tests/golden/mini_model.npz).

Tolerances: bf16 path vs fp32 oracle: max |err| <= 2e-2 x max |ref| (north_star); fp32-only kernels (losses on given
logits, logits rules): 1e-4 / exact.  Greedy decode: token-identical; where bf16 rounding flips a near-tie on these
random-weight models the test requires the oracle's own margin at that step to be below MARGIN (stated below)."""
import dataclasses
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dicow_oracle as orc
from oracle import synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
MARGIN = 0.15  # logit units; bf16 logits of these models carry ~0.05 abs error
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EOS, SOT, LANG, TASK, NOTS, TS_BEGIN, N_TS = 257, 258, 259, 260, 261, 262, 38
SUPPRESS = [1, 2, 7, 8, 9, 10, 14, 25, 258, 259, 260]
DEV = "cuda:0"


class FakeTokenizer:
    """just enough of WhisperTokenizer for SoftLabelCreator / prefix stripping (no tokenizer files offline)"""
    prefix_tokens = [SOT, LANG, TASK]
    pad_token_id = EOS

    def get_vocab(self):
        v = {f"<|{0.02 * i:.2f}|>": TS_BEGIN + i for i in range(N_TS)}
        v["Ġ"] = 220
        return v


@pytest.fixture(scope="module")
def ops():
    from ts_asr_whisper_b200 import ops as _ops
    return _ops


def build_model(dm: synth.Dims):
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    params = synth.make_params(dm)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    model.tie_weights()
    return model.to(DEV).eval(), orc.to_torch(params)


def rel_err(out, ref):
    return ((out.float().cpu() - ref).abs().max() / ref.abs().max().clamp(min=1e-6)).item()


# ----------------------------------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,epi", [(16, 1280, 1280, 0), (3, 384, 384, 1), (16, 1000, 5120, 2), (33, 51866, 128, 3),
                                       (64, 264, 256, 0), (1, 8, 32, 3)])
def test_gemm_skinny(ops, M, N, K, epi):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=DEV, generator=g)
    ref = A.float() @ W.float().t() + b
    if epi in (0, 1):
        if epi == 1:
            ref = F.gelu(ref)
        out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.gemm_skinny(A, W, out, epilogue=epi, bias=b)
    elif epi == 2:
        res = torch.randn(M, N, device=DEV, generator=g)
        ref = res + ref
        out = res.clone()
        ops.gemm_skinny(A, W, out, epilogue=epi, bias=b, resid=out)
    else:
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm_skinny(A, W, out, epilogue=epi, bias=b)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    assert rel_err(out, ref.cpu()) < (1e-2 if out.dtype == torch.bfloat16 else 1e-4)


def test_gemm_skinny_cache_append(ops):
    """KV-cache append: rows land at cache[b, *pos, :] through ldo / pos / pos_stride."""
    B, S, d, K = 5, 12, 128, 64
    g = torch.Generator(device=DEV).manual_seed(3)
    A = torch.randn(B, K, device=DEV, generator=g).bfloat16()
    W = (torch.randn(2 * d, K, device=DEV, generator=g) * 0.1).bfloat16()
    cache = torch.zeros(B, S, 2 * d, device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([7], dtype=torch.int32, device=DEV)
    ops.gemm_skinny(A, W, cache, epilogue=0, ldo=S * 2 * d, pos=pos, pos_stride=2 * d)
    torch.cuda.synchronize()
    ref = (A.float() @ W.float().t()).cpu()
    assert rel_err(cache[:, 7], ref) < 1e-2
    assert float(cache[:, :7].abs().max()) == 0.0 and float(cache[:, 8:].abs().max()) == 0.0


@pytest.mark.parametrize("B,H,Tk,use_pos", [(16, 20, 1500, False), (3, 2, 1, False), (2, 6, 37, True), (5, 4, 448, True)])
def test_decode_attention(ops, B, H, Tk, use_pos):
    d = H * 64
    g = torch.Generator(device=DEV).manual_seed(Tk + B)
    Tcap = Tk + 5
    q = (torch.randn(B, d, device=DEV, generator=g) * 0.4).bfloat16()
    kv = (torch.randn(B, Tcap, 2 * d, device=DEV, generator=g) * 1.1).bfloat16()
    out = torch.full((B, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([Tk - 1], dtype=torch.int32, device=DEV) if use_pos else None
    ops.decode_attention(q, kv, kv[:, :, d:], out, B=B, H=H, Tk=0 if use_pos else Tk, kv_row_stride=2 * d,
                         kv_batch_stride=Tcap * 2 * d, pos=pos)
    torch.cuda.synchronize()
    qf = q.float().view(B, H, 1, 64)
    kf = kv[:, :Tk, :d].float().view(B, Tk, H, 64).transpose(1, 2)
    vf = kv[:, :Tk, d:].float().view(B, Tk, H, 64).transpose(1, 2)
    ref = (torch.softmax(qf @ kf.transpose(-1, -2), -1) @ vf).reshape(B, d)
    assert rel_err(out, ref.cpu()) < 2e-2


# (M, N, K, epilogue, LayerNorm prologue): covers one CTA per tile, several tiles per CTA (N = 51866), K split over a
# cluster of 2 / 4 / 8 CTAs (N = 1280 with K = 1280 / 5120 / 4096+), row tiles MT = 1..4, ragged N and tiny shapes
DL_CASES = [(16, 3840, 1280, 0, True), (16, 1280, 1280, 2, False), (16, 1280, 5120, 2, False), (16, 5120, 1280, 1, True),
            (16, 51866, 1280, 3, True), (33, 1000, 384, 3, True), (64, 264, 256, 0, False), (1, 8, 32, 3, False),
            (5, 1280, 1280, 0, True), (48, 384, 1536, 2, False), (2, 128, 128, 1, True), (16, 640, 8192, 3, False),
            # LayerNorm prologue with K split over a cluster pair and several tiles per pass (the cross-attention q
            # projection of the step), 17..32 rows with the prologue, a ragged last tile group
            (16, 1280, 1280, 2, True), (20, 2560, 1280, 0, True), (32, 5120, 1280, 1, True), (16, 3000, 640, 3, True),
            # the beam-search step at 60 hypotheses: two 32-row CTAs per tile group, several tiles per pass, K split 1 / 2 / 4
            (60, 5120, 1280, 1, False), (60, 1280, 5120, 2, False), (60, 1280, 1280, 2, False), (60, 3840, 1280, 0, True)]


@pytest.mark.parametrize("M,N,K,epi,ln", DL_CASES)
def test_decode_linear(ops, M, N, K, epi, ln):
    """fused [LayerNorm ->] Linear of the decode step vs fp32 torch on the same bf16-rounded operands"""
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N + K)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=DEV, generator=g)
    kw = {}
    if ln:
        x = torch.randn(M, K, device=DEV, generator=g) * 1.7 + 0.3
        gamma = torch.rand(K, device=DEV, generator=g) + 0.5
        beta = torch.randn(K, device=DEV, generator=g) * 0.1
        A = F.layer_norm(x, (K,), gamma, beta, 1e-5).bfloat16()
        kw = dict(x=x, gamma=gamma, beta=beta)
    else:
        A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
        kw = dict(A=A)
    ref = A.float() @ W.float().t() + b
    if epi in (0, 1):
        if epi == 1:
            ref = F.gelu(ref)
        out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.decode_linear(W, out, epilogue=epi, bias=b, **kw)
    elif epi == 2:
        res = torch.randn(M, N, device=DEV, generator=g)
        ref = res + ref
        out = res.clone()
        ops.decode_linear(W, out, epilogue=epi, bias=b, resid=out, **kw)
    else:
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.decode_linear(W, out, epilogue=epi, bias=b, **kw)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    # the LayerNorm prologue rounds A to bf16 like the reference's autocast; a 1-ulp bf16 difference of A against the
    # torch LayerNorm shows up at the 1e-3 level of the fp32 output
    tol = 1e-2 if out.dtype == torch.bfloat16 else (2e-3 if ln else 1e-4)
    assert rel_err(out, ref.cpu()) < tol


def test_decode_linear_fused_qkv_cache_append(ops):
    """LayerNorm -> [q | k,v]: q to its buffer, k,v appended to cache[b, *pos, :] (one kernel of the decode step)"""
    B, S, d = 16, 12, 1280
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(B, d, device=DEV, generator=g)
    gamma, beta = torch.rand(d, device=DEV, generator=g) + 0.5, torch.randn(d, device=DEV, generator=g) * 0.1
    W = (torch.randn(3 * d, d, device=DEV, generator=g) * 0.03).bfloat16()
    bias = torch.randn(3 * d, device=DEV, generator=g)
    q = torch.full((B, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    cache = torch.zeros(B, S, 2 * d, device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([7], dtype=torch.int32, device=DEV)
    ops.decode_linear(W, q, x=x, gamma=gamma, beta=beta, epilogue=0, bias=bias, out2=cache, n_split=d, ldo2=S * 2 * d,
                      pos=pos, pos_stride=2 * d)
    torch.cuda.synchronize()
    A = F.layer_norm(x, (d,), gamma, beta, 1e-5).bfloat16().float()
    ref = (A @ W.float().t() + bias).cpu()
    assert rel_err(q, ref[:, :d]) < 1e-2
    assert rel_err(cache[:, 7], ref[:, d:]) < 1e-2
    assert float(cache[:, :7].abs().max()) == 0.0 and float(cache[:, 8:].abs().max()) == 0.0


def test_logits_rules_cluster_matches_single_cta(ops):
    """Whisper-sized vocabulary: the row scan split over a cluster of 8 CTAs (production path) picks the same tokens as
    the single-CTA path that also materialises the processed scores"""
    B, V, P, ngen = 16, 51866, 3, 9
    ts_begin, eos, nots = 50365, 50257, 50364
    rng = np.random.default_rng(4)
    ids = torch.zeros(B, P + ngen + 1, dtype=torch.int64)
    ids[:, :P] = torch.tensor([50258, 50259, 50360])
    for b in range(B):
        hist, t = [], ts_begin
        while len(hist) < ngen:
            t = min(t + int(rng.integers(0, 40)), V - 1)
            hist.append(t)
            hist += [int(x) for x in rng.integers(300, 40000, size=int(rng.integers(0, 4)))]
            if rng.integers(0, 2):
                hist.append(t)
        ids[b, P:P + ngen] = torch.tensor(hist[:ngen])
    logits = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 3.0
    logits[::3, ts_begin:] += 5.0
    bitmap = ops.suppress_bitmap([220, 50256, 1, 2, 7, 359, 503], V, DEV)
    kw = dict(begin_index=P, eos=eos, pad=eos, no_timestamps=nots, ts_begin=ts_begin, cur_len=P + ngen, suppress_bitmap=bitmap)
    out = []
    for proc in (torch.empty(B, V, device=DEV), None):
        d_ids = ids.to(DEV)
        unf = torch.ones(B, dtype=torch.int32, device=DEV)
        unf[2] = 0
        ops.logits_rules_argmax(logits.to(DEV), d_ids, unf, processed_scores=proc, **kw)
        torch.cuda.synchronize()
        out.append((d_ids[:, P + ngen].cpu().tolist(), unf.cpu().tolist()))
    assert out[0] == out[1]
    assert len(set(out[0][0])) > 3  # not a degenerate comparison


@pytest.mark.parametrize("graphs", [False, True])
def test_fused_decode_step_matches_unfused(graphs):
    """large-v3-turbo decoder dims, B = 16: the fused step (LayerNorm prologues, fused q|k,v, cluster split-K, PDL)
    generates the same tokens and first-step logits as the one-kernel-per-op step"""
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    cfg = DiCoWConfig(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=1, encoder_attention_heads=20,
                      decoder_layers=4, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120,
                      max_source_positions=1500, max_target_positions=448, use_fddt=True, ctc_weight=0.0,
                      pad_token_id=50257, eos_token_id=50257)
    torch.manual_seed(0)
    with torch.device(DEV):
        model = DiCoWForConditionalGeneration(cfg)
    model.eval()
    model.use_cuda_graphs = graphs
    B, T = 16, 1500
    g = torch.Generator(device=DEV).manual_seed(1)
    enc = (torch.randn(B, T, cfg.d_model, device=DEV, generator=g) * 0.5).bfloat16()
    prompt = torch.tensor([[50258, 50259, 50360]] * B, device=DEV)
    rules = dict(eos=50257, pad=50257, no_timestamps=50364, ts_begin=50365, max_initial_timestamp_index=None,
                 timestamp_rules=True, suppress_bitmap=model._suppress_bitmap([220, 50256], torch.device(DEV)))
    res = {}
    for fused in (False, True, "ln_prologue"):
        model.fused_decode_step = fused
        ids, first = model.greedy_decode_window(enc, prompt, 3 + 20, rules, return_first_logits=True)
        torch.cuda.synchronize()
        res[fused] = (ids.cpu(), first.float().cpu())
    for mode in (True, "ln_prologue"):
        err = ((res[mode][1] - res[False][1]).abs().max() / res[False][1].abs().max()).item()
        same = (res[mode][0] == res[False][0]).float().mean().item()
        print(f"fused ({mode}) vs unfused: first-step logits rel err {err:.3e}, token agreement {same:.3f}")
        assert err < 5e-3
        # the steps round intermediate activations identically (bf16 LayerNorm output, bf16 q/k/v/ctx/h) and differ only
        # in fp32 summation order, so tokens agree except where a random-weight argmax is a near tie that then diverges
        assert torch.equal(res[mode][0][:, :4], res[False][0][:, :4])
        assert same > 0.9


def test_decode_attention_head_major_cache(ops):
    """cross-attention cache layout of the decode step: projection rows [B*T, k | v] -> [B, H, T, 128] (kv_to_head_major),
    then one query per (batch, head) over the contiguous per-head stream (kv_head_stride)"""
    B, H, T = 16, 20, 1500
    d = H * 64
    g = torch.Generator(device=DEV).manual_seed(5)
    q = (torch.randn(B, d, device=DEV, generator=g) * 0.4).bfloat16()
    rows = (torch.randn(B * T, 2 * d, device=DEV, generator=g) * 1.1).bfloat16()
    hm = torch.full((B, H, T, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.kv_to_head_major(rows, hm, B=B, T=T, H=H)
    torch.cuda.synchronize()
    r4 = rows.view(B, T, 2, H, 64)
    assert torch.equal(hm[..., :64], r4[:, :, 0].transpose(1, 2)) and torch.equal(hm[..., 64:], r4[:, :, 1].transpose(1, 2))
    out = torch.full((B, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.decode_attention(q, hm, hm[..., 64:], out, B=B, H=H, Tk=T, kv_row_stride=128, kv_batch_stride=H * T * 128,
                         kv_head_stride=T * 128)
    ref = torch.full((B, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.decode_attention(q, rows, rows[:, d:], ref, B=B, H=H, Tk=T, kv_row_stride=2 * d, kv_batch_stride=T * 2 * d)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    assert rel_err(out, ref.float().cpu()) < 1e-3


def _random_history(rng, n):
    """a plausible generated sequence: text runs separated by timestamp pairs, random phase at the end"""
    seq, t = [], TS_BEGIN
    while len(seq) < n:
        t = min(t + int(rng.integers(0, 4)), TS_BEGIN + N_TS - 1)
        seq.append(t)
        for _ in range(int(rng.integers(0, 4))):
            seq.append(int(rng.integers(11, 250)))
        t = min(t + int(rng.integers(0, 3)), TS_BEGIN + N_TS - 1)
        seq.append(t)
    return seq[:n]


@pytest.mark.parametrize("ngen", [0, 1, 2, 3, 7, 12])
def test_logits_rules_match_oracle(ops, ngen):
    """processed scores and argmax vs oracle.timestamp_rules (pinned to the reference's processors by the golden test)"""
    B, V, P = 6, 300, 3
    rng = np.random.default_rng(ngen)
    ids = torch.zeros(B, P + ngen + 1, dtype=torch.int64)
    ids[:, :P] = torch.tensor([SOT, LANG, TASK])
    for b in range(B):
        ids[b, P:P + ngen] = torch.tensor(_random_history(rng, ngen), dtype=torch.int64) if ngen else ids[b, P:P]
    logits = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 3.0
    logits[0, TS_BEGIN:] += 4.0  # force the "timestamp mass > best text" rule on one row
    s = logits.clone()
    s[:, SUPPRESS] = -float("inf")
    ref = orc.timestamp_rules(ids[:, :P + ngen], s, begin_index=P, eos=EOS, no_timestamps=NOTS, ts_begin=TS_BEGIN)
    ref_tok = ref.argmax(-1)
    d_ids, d_logits = ids.to(DEV), logits.to(DEV)
    unf = torch.ones(B, dtype=torch.int32, device=DEV)
    unf[1] = 0  # finished row -> pad
    proc = torch.empty(B, V, device=DEV)
    ops.logits_rules_argmax(d_logits, d_ids, unf, begin_index=P, eos=EOS, pad=EOS, no_timestamps=NOTS, ts_begin=TS_BEGIN,
                            cur_len=P + ngen, suppress_bitmap=ops.suppress_bitmap(SUPPRESS, V, DEV), processed_scores=proc)
    torch.cuda.synchronize()
    assert torch.equal(torch.isinf(proc.cpu()), torch.isinf(ref)), "masked set differs"
    fin = ~torch.isinf(ref)
    assert torch.equal(proc.cpu()[fin], ref[fin])
    tok = d_ids[:, P + ngen].cpu()
    expect = ref_tok.clone()
    expect[1] = EOS
    assert tok.tolist() == expect.tolist()
    exp_unf = [(0 if b == 1 else int(expect[b] != EOS)) for b in range(B)]
    assert unf.cpu().tolist() == exp_unf


def test_softlabel_ce_and_ctc(ops):
    rng = np.random.default_rng(0)
    R, V = 40, 300
    logits = torch.from_numpy(rng.normal(size=(R, V)).astype(np.float32)) * 2
    labels = torch.from_numpy(synth.make_labels("ce", 4, 10, V, EOS, TS_BEGIN, prefix=(LANG, TASK))).reshape(-1)
    upp = labels.clone()
    upp[::5] = torch.where(upp[::5] >= 0, (upp[::5] + 3) % 250, upp[::5])
    ref_soft = orc.decoder_loss(logits.view(4, 10, V), labels.view(4, 10), upp.view(4, 10), TS_BEGIN, N_TS)
    ref_hard = orc.decoder_loss(logits.view(4, 10, V), labels.view(4, 10), upp.view(4, 10))
    sm = orc.timestamp_smoothing(N_TS).to(DEV)
    got_soft = ops.softlabel_ce(logits.to(DEV), labels.to(DEV), upp.to(DEV), ts_begin=TS_BEGIN, smoothing=sm, soft_mode=True)
    got_hard = ops.softlabel_ce(logits.to(DEV), labels.to(DEV), upp.to(DEV), soft_mode=False)
    got_hard1 = ops.softlabel_ce(logits.to(DEV), labels.to(DEV), None, soft_mode=False)
    assert abs(got_soft.item() - ref_soft.item()) < 1e-4 * max(1, abs(ref_soft.item()))
    assert abs(got_hard.item() - ref_hard.item()) < 1e-4 * max(1, abs(ref_hard.item()))
    assert abs(got_hard1.item() - orc.decoder_loss(logits.view(4, 10, V), labels.view(4, 10), None).item()) < 1e-4
    # CTC vs torch's own ctc_loss on the same fp32 logits (the reference's call, encoder.py:123-134)
    B, T, V1 = 5, 40, 301
    lg = torch.from_numpy(rng.normal(size=(B, T, V1)).astype(np.float32)) * 2
    lab = torch.full((B, 14), -100, dtype=torch.int64)
    for b, n in enumerate([14, 3, 0, 9, 14]):
        lab[b, :n] = torch.from_numpy(rng.integers(0, 20, size=n))  # small alphabet: repeated labels occur
    lab[4, :14] = 5  # all-repeats: needs 2 L + 1 = 29 <= T frames, fine; also try an infeasible one below
    for red in ("mean", "sum"):
        ref = orc.ctc_loss(lg, lab, reduction=red)
        got = ops.ctc_loss(lg.to(DEV), lab.to(DEV), reduction=red)
        assert abs(got.item() - ref.item()) < 2e-4 * max(1, abs(ref.item())), (red, got.item(), ref.item())
    short = lg[:, :10].contiguous()  # 14 labels do not fit 10 frames -> inf -> zero_infinity
    ref = orc.ctc_loss(short, lab)
    got = ops.ctc_loss(short.to(DEV), lab.to(DEV))
    assert abs(got.item() - ref.item()) < 2e-4 * max(1, abs(ref.item()))


# ----------------------------------------------------------------------------------------------------------------
# model level
# ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mini():
    dm = synth.GOLDEN_MINI
    dmp = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
    model, p = build_model(dmp)
    feats = torch.from_numpy(synth.make_features("g0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", 2, dm.T, "soft", pad_tail=7))
    g = np.load(os.path.join(GOLD, "mini_model.npz"))
    return g, dmp, model, p, feats, stno


def test_forward_losses_match_oracle_and_golden(mini):
    g, dmp, model, p, feats, stno = mini
    labels = torch.from_numpy(g["labels"])
    upp = torch.from_numpy(g["upp_labels"])
    model.tokenizer, model.soft_label_creator = None, None
    model.ctc_prefix_tokens = (SOT, LANG, TASK)
    out = model(feats.to(DEV), stno_mask=stno.to(DEV), labels=labels.to(DEV), upp_labels=upp.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(out.logits, torch.from_numpy(g["fwd_logits"])) < BF16_TOL
    assert abs(out.loss.item() - float(g["fwd_hard_loss"])) < BF16_TOL * max(1.0, abs(float(g["fwd_hard_loss"])))
    model.set_tokenizer(FakeTokenizer())
    out2 = model(feats.to(DEV), stno_mask=stno.to(DEV), labels=labels.to(DEV), upp_labels=upp.to(DEV))
    assert abs(out2.loss.item() - float(g["fwd_soft_loss"])) < BF16_TOL * max(1.0, abs(float(g["fwd_soft_loss"])))
    with torch.no_grad():
        ref_loss, ref_logits, _ = orc.model_forward(p, dmp, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK),
                                                    ts_begin=TS_BEGIN, n_ts=N_TS)
    assert rel_err(out2.logits, ref_logits) < BF16_TOL
    print(f"mini forward: logits rel err {rel_err(out2.logits, ref_logits):.3e}; loss {out2.loss.item():.5f} vs {ref_loss.item():.5f}")


def _check_greedy(ids, ref_ids, ref_raw_logits, P, dm, max_flips=1):
    """token identity against the oracle's own greedy run; a divergence is only tolerated where the oracle's top-2 margin at
    that step is below MARGIN, it ends the comparison of that row (the continuations legitimately differ afterwards), and
    the NUMBER of rows that diverge is bounded.  (tests/test_gpu_turbo_parity.py checks every step of every row instead, by
    teacher-forcing the oracle on the CUDA tokens.)"""
    ids, ref_ids = ids.cpu(), ref_ids.cpu()
    n = min(ids.shape[1], ref_ids.shape[1])
    flips, compared = 0, 0
    for b in range(ids.shape[0]):
        for t in range(P, n):
            compared += 1
            if int(ids[b, t]) == int(ref_ids[b, t]):
                continue
            raw = ref_raw_logits[t - P][b:b + 1].clone()
            raw[:, SUPPRESS] = -float("inf")
            proc = orc.timestamp_rules(ref_ids[b:b + 1, :t], raw, begin_index=P, eos=dm.eos_token_id, no_timestamps=NOTS,
                                       ts_begin=TS_BEGIN)[0]
            margin = float(proc[int(ref_ids[b, t])] - proc[int(ids[b, t])])
            assert margin < MARGIN, f"row {b} step {t - P}: token {int(ids[b, t])} vs {int(ref_ids[b, t])}, margin {margin:.3f}"
            flips += 1
            break  # after a tolerated near-tie flip the continuations legitimately differ
    print(f"greedy: {compared} tokens compared, {flips} row(s) diverged at a near-tie (oracle margin < {MARGIN})")
    assert flips <= max_flips, f"{flips} rows diverged from the oracle (bound {max_flips})"


@pytest.fixture(params=[False, True], ids=["kernel-per-op", "megakernel"])
def decode_path(request, mini):
    """both decode-step implementations: the default kernel-per-operation sequence and the experimental persistent
    decode-layers kernel (csrc/decode_mega.cu)"""
    model = mini[2]
    old = model.decode_megakernel
    model.decode_megakernel = request.param
    model.clear_decode_cache()
    yield request.param
    model.decode_megakernel = old
    model.clear_decode_cache()


@pytest.mark.parametrize("graphs", [False, True])
def test_greedy_decode_matches_oracle_and_golden(mini, graphs, decode_path):
    g, dmp, model, p, feats, stno = mini
    model.use_cuda_graphs = graphs
    prompt = torch.tensor([[SOT, LANG, TASK]] * 2)
    enc = model.get_encoder()(feats.to(DEV), stno_mask=stno.to(DEV)).last_hidden_state
    rules = dict(eos=EOS, pad=EOS, no_timestamps=NOTS, ts_begin=TS_BEGIN, max_initial_timestamp_index=None,
                 timestamp_rules=True, suppress_bitmap=model._suppress_bitmap(SUPPRESS, torch.device(DEV)))
    ids, first = model.greedy_decode_window(enc, prompt.to(DEV), 3 + 24, rules, return_first_logits=True)
    torch.cuda.synchronize()
    assert rel_err(first, torch.from_numpy(g["greedy_first_logits"])) < BF16_TOL
    with torch.no_grad():
        ref_enc = orc.encoder_forward(p, dmp, feats, stno)
        ref_ids, ref_lg = orc.greedy_decode(p, dmp, ref_enc, prompt, 24, suppress=SUPPRESS, no_timestamps=NOTS,
                                            ts_begin=TS_BEGIN, return_logits=True)
    assert ref_ids.tolist() == g["greedy_ids"].tolist()
    print("greedy ids (cuda):", ids.cpu().tolist())
    _check_greedy(ids, ref_ids, ref_lg, 3, dmp)
    # second call reuses buffers / graphs and must reproduce itself exactly
    ids2 = model.greedy_decode_window(enc, prompt.to(DEV), 3 + 24, rules)
    assert torch.equal(ids, ids2)


def test_forward_tiny_dims():
    """whisper-tiny dims (BASELINE configs[0]) with a small vocabulary: teacher-forced logits + combined loss"""
    dm = dataclasses.replace(synth.WHISPER_TINY, vocab=2047, pad_token_id=257, eos_token_id=257,
                             decoder_start_token_id=258, max_target=64)
    model, p = build_model(dm)
    model.ctc_prefix_tokens = (SOT, LANG, TASK)
    feats = torch.from_numpy(synth.make_features("t0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("t0", 2, dm.T, "soft", pad_tail=30))
    labels = torch.from_numpy(synth.make_labels("t0", 2, 24, 250, EOS, TS_BEGIN, prefix=(LANG, TASK)))
    out = model(feats.to(DEV), stno_mask=stno.to(DEV), labels=labels.to(DEV), upp_labels=labels.to(DEV))
    with torch.no_grad():
        ref_loss, ref_logits, ref_enc = orc.model_forward(p, dm, feats, stno, labels, labels,
                                                          ctc_prefix_tokens=(SOT, LANG, TASK))
    e = rel_err(out.logits, ref_logits)
    print(f"tiny forward: logits rel err {e:.3e}; loss {out.loss.item():.5f} vs {ref_loss.item():.5f}")
    assert e < BF16_TOL
    assert abs(out.loss.item() - ref_loss.item()) < BF16_TOL * max(1.0, abs(ref_loss.item()))


def test_generate_long_form_runs_and_segments(mini):
    """two recordings of different length through the seek loop; checks the seek bookkeeping invariants"""
    g, dmp, model, p, feats, stno = mini
    model.tokenizer, model.soft_label_creator = None, None
    F2 = 2 * dmp.T
    long_feats = torch.cat([feats, torch.from_numpy(synth.make_features("g1", 2, dmp.n_mels, F2))], dim=-1)
    long_stno = torch.cat([stno, torch.from_numpy(synth.make_stno("g1", 2, dmp.T, "hard"))], dim=-1)
    attn = torch.ones(2, 2 * F2, dtype=torch.long)
    attn[1, F2 + 31:] = 0  # second recording is shorter
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = NOTS, EOS, EOS
    gc.suppress_tokens, gc.return_timestamps, gc.max_new_tokens, gc.num_beams = SUPPRESS, True, 24, 1
    out = model.generate(long_feats.to(DEV), attention_mask=attn.to(DEV), stno_mask=long_stno.to(DEV),
                         forced_decoder_ids=torch.tensor([[SOT, LANG, TASK]] * 2), return_segments=True)
    assert out["sequences"].shape[0] == 2 and len(out["segments"]) == 2
    for segs in out["segments"]:
        last = 0.0
        for s in segs:
            assert float(s["start"]) <= float(s["end"]) and float(s["end"]) >= last - 1e-9
            last = float(s["end"])
    seqs = model.generate(long_feats.to(DEV), attention_mask=attn.to(DEV), stno_mask=long_stno.to(DEV),
                          forced_decoder_ids=torch.tensor([[SOT, LANG, TASK]] * 2))
    assert torch.equal(seqs, out["sequences"])


@pytest.mark.parametrize("timestamps", [True, False], ids=["timestamps", "no-timestamps"])
def test_generate_speculative_next_window_is_transparent(mini, timestamps):
    """with ``speculate_next_window`` the window at seek + 3000 is encoded on a second stream (part of the SMs)
    under the decode steps of the current window.  Sequences and segments must be exactly those of the plain seek loop; without
    timestamps every window advances by a full window, so every speculated window must be used."""
    g, dmp, model, p, feats, stno = mini
    model.tokenizer, model.soft_label_creator = None, None
    F2 = 2 * dmp.T
    more = [torch.from_numpy(synth.make_features(f"g{k}", 2, dmp.n_mels, F2)) for k in (1, 3, 4)]
    long_feats = torch.cat([feats] + more, dim=-1).to(DEV)
    long_stno = torch.cat([stno] + [torch.from_numpy(synth.make_stno(f"g{k}", 2, dmp.T, "hard")) for k in (1, 3, 4)], dim=-1).to(DEV)
    attn = torch.ones(2, 4 * F2, dtype=torch.long)
    attn[1, 2 * F2 + 31:] = 0  # second recording is shorter
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = NOTS, EOS, EOS
    gc.suppress_tokens, gc.return_timestamps, gc.max_new_tokens, gc.num_beams = SUPPRESS, timestamps, 24, 1
    prompt = torch.tensor([[SOT, LANG, TASK] + ([] if timestamps else [NOTS])] * 2)
    kw = dict(attention_mask=attn.to(DEV), stno_mask=long_stno, forced_decoder_ids=prompt, return_segments=True)
    try:
        model.speculate_next_window = False
        plain = model.generate(long_feats, **kw)
        model.speculate_next_window = True
        model.speculation_sms = 64
        spec = model.generate(long_feats, **kw)
        stats = dict(model.speculation_stats)
    finally:
        model.speculate_next_window = False
        model.speculation_sms = None  # automatic budget (exercised by the SE-DiCoW long-form test)
        gc.return_timestamps = True
    print("speculation:", stats)
    assert torch.equal(plain["sequences"], spec["sequences"])
    assert [[(s["tokens"].tolist(), float(s["start"]), float(s["end"])) for s in r] for r in plain["segments"]] == \
           [[(s["tokens"].tolist(), float(s["start"]), float(s["end"])) for s in r] for r in spec["segments"]]
    assert stats["hits"] + stats["misses"] >= 1
    if not timestamps:
        assert stats["misses"] == 0 and stats["hits"] >= 3


def test_generate_joint_ctc_long_form(mini):
    """generate(ctc_weight > 0): the long-form seek loop with joint CTC / attention selection per window (encoder CTC
    logits -> rescoring inside the CUDA-graphed step); shrinking batch (second recording shorter) re-allocates the CTC state"""
    g, dmp, model, p, feats, stno = mini
    model.tokenizer, model.soft_label_creator = None, None
    F2 = 2 * dmp.T
    long_feats = torch.cat([feats, torch.from_numpy(synth.make_features("g1", 2, dmp.n_mels, F2))], dim=-1)
    long_stno = torch.cat([stno, torch.from_numpy(synth.make_stno("g1", 2, dmp.T, "hard"))], dim=-1)
    attn = torch.ones(2, 2 * F2, dtype=torch.long)
    attn[1, F2 + 31:] = 0
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = NOTS, EOS, EOS
    gc.suppress_tokens, gc.return_timestamps, gc.max_new_tokens, gc.num_beams = SUPPRESS, True, 24, 1
    kw = dict(attention_mask=attn.to(DEV), stno_mask=long_stno.to(DEV), forced_decoder_ids=torch.tensor([[SOT, LANG, TASK]] * 2),
              return_segments=True)
    plain = model.generate(long_feats.to(DEV), **kw)
    joint = model.generate(long_feats.to(DEV), ctc_weight=0.3, ctc_tokens_to_score=40, **kw)
    again = model.generate(long_feats.to(DEV), ctc_weight=0.3, ctc_tokens_to_score=40, **kw)
    assert torch.equal(joint["sequences"], again["sequences"])
    assert model.encoder_logits is None  # generation.py:559
    assert joint["sequences"].shape[0] == 2
    # the CTC evidence changes what is decoded on this model (attention alone repeats one token until max_new_tokens)
    assert not torch.equal(joint["sequences"], plain["sequences"])
    beams = model.generate(long_feats.to(DEV), ctc_weight=0.3, ctc_tokens_to_score=40, num_beams=3, length_penalty=0.1, **kw)
    beams2 = model.generate(long_feats.to(DEV), ctc_weight=0.3, ctc_tokens_to_score=40, num_beams=3, length_penalty=0.1, **kw)
    assert beams["sequences"].shape[0] == 2 and torch.equal(beams["sequences"], beams2["sequences"])


def test_detect_language_and_prompt_without_forced_ids(mini):
    """generation.py:151-221: one decoder step on <|sot|> over the first window, non-language logits masked, argmax --
    against the oracle's forward; generate() without forced_decoder_ids builds <|sot|> <|lang|> <|task|> from it"""
    g, dmp, model, p, feats, stno = mini
    model.tokenizer, model.soft_label_creator = None, None
    lang_to_id = {"<|aa|>": 259, "<|bb|>": 30, "<|cc|>": 77, "<|dd|>": 201}
    gc = model.generation_config
    gc.lang_to_id, gc.task_to_id = lang_to_id, {"transcribe": TASK, "translate": 5}
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = NOTS, EOS, EOS
    gc.suppress_tokens, gc.return_timestamps, gc.max_new_tokens, gc.num_beams = SUPPRESS, True, 6, 1
    gc.decoder_start_token_id = SOT
    ids = model.detect_language(input_features=feats.to(DEV), stno_mask=stno.to(DEV))
    with torch.no_grad():
        ref_enc = orc.encoder_forward(p, dmp, feats, stno)
        hid = orc.decoder_forward(p, dmp, torch.full((2, 1), SOT), ref_enc)
        logits = torch.nn.functional.linear(hid[:, -1], p["proj_out.weight"])
    keep = torch.tensor(sorted(lang_to_id.values()))
    ref = keep[logits[:, keep].argmax(-1)]
    top2 = logits[:, keep].topk(2, dim=-1).values
    for b in range(2):
        if float(top2[b, 0] - top2[b, 1]) > MARGIN:
            assert int(ids[b]) == int(ref[b])
    gc.task = "transcribe"  # as the reference's container sets it (src/models/containers.py:59): <|sot|><|lang|><|transcribe|>
    out = model.generate(feats.to(DEV), stno_mask=stno.to(DEV), return_segments=True)
    assert out["sequences"].shape[0] == 2
    forced = model.generate(feats.to(DEV), stno_mask=stno.to(DEV), return_segments=True,
                            forced_decoder_ids=torch.stack([torch.full((2,), SOT), ids.cpu(), torch.full((2,), TASK)], 1))
    assert torch.equal(out["sequences"], forced["sequences"])
    with pytest.raises(ValueError):
        model.generate(feats.to(DEV), stno_mask=stno.to(DEV), language="zz")


# ---- SE-DiCoW generate(): enrollments, and the enrollment key/value cache of the long-form loop ---------------------------
@pytest.fixture(scope="module")
def mini_se():
    dm = synth.GOLDEN_MINI  # use_enrollments, 2 speaker communication blocks
    model, p = build_model(dm)
    feats = torch.from_numpy(synth.make_features("g0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", 2, dm.T, "soft", pad_tail=7))
    enr = {"input_features": torch.from_numpy(synth.make_features("g0e", 2, dm.n_mels, 2 * dm.T)),
           "stno_mask": torch.from_numpy(synth.make_stno("g0e", 2, dm.T, "hard"))}
    return dm, model, p, feats, stno, enr


def _gen_cfg(model, max_new):
    model.tokenizer, model.soft_label_creator = None, None
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = NOTS, EOS, EOS
    gc.suppress_tokens, gc.return_timestamps, gc.max_new_tokens, gc.num_beams = SUPPRESS, True, max_new, 1


def test_encoder_enrollment_kv_cache_is_exact(mini_se):
    """the enrollment stream never reads the target stream (layers.py:145-170): its projected keys / values captured on one
    window reproduce the full two-stream forward of ANOTHER window bit for bit"""
    dm, model, p, feats, stno, enr = mini_se
    enc = model.get_encoder()
    enr_d = {k: v.to(DEV) for k, v in enr.items()}
    cap = []
    first = enc(feats.to(DEV), stno_mask=stno.to(DEV), enrollments=enr_d, capture_enrollment_kv=cap).last_hidden_state
    assert len(cap) == dm.scb_layers and cap[0].shape == (2, dm.T, 2 * dm.d)
    feats2 = torch.from_numpy(synth.make_features("g2", 2, dm.n_mels, 2 * dm.T)).to(DEV)
    stno2 = torch.from_numpy(synth.make_stno("g2", 2, dm.T, "soft")).to(DEV)
    full = enc(feats2, stno_mask=stno2, enrollments=enr_d).last_hidden_state
    cached = enc(feats2, stno_mask=stno2, enrollment_kv=cap).last_hidden_state
    assert torch.equal(full, cached) and not torch.equal(full, first)
    swapped = enc(feats2, stno_mask=stno2, enrollment_kv=[c.flip(0) for c in cap]).last_hidden_state
    assert not torch.equal(swapped, full)  # the cache really is what conditions the targets
    with pytest.raises(ValueError, match="either"):
        enc(feats2, stno_mask=stno2, enrollments=enr_d, enrollment_kv=cap)


def test_generate_se_dicow_matches_oracle_and_cache_is_transparent(mini_se):
    dm, model, p, feats, stno, enr = mini_se
    _gen_cfg(model, 16)
    prompt = torch.tensor([[SOT, LANG, TASK]] * 2)
    enr_d = {k: v.to(DEV) for k, v in enr.items()}
    # one window: token parity with the oracle's two-stream encoder + greedy loop
    out = model.generate(feats.to(DEV), stno_mask=stno.to(DEV), enrollments=enr_d, forced_decoder_ids=prompt,
                         return_segments=True)
    with torch.no_grad():
        ref_enc = orc.encoder_forward(p, dm, feats, stno, enrollments=enr)
        ref_ids, ref_lg = orc.greedy_decode(p, dm, ref_enc, prompt, 16, suppress=SUPPRESS, no_timestamps=NOTS,
                                            ts_begin=TS_BEGIN, return_logits=True)
    enc_dev = model.get_encoder()(feats.to(DEV), stno_mask=stno.to(DEV), enrollments=enr_d).last_hidden_state
    assert rel_err(enc_dev, ref_enc) < BF16_TOL
    rules = dict(eos=EOS, pad=EOS, no_timestamps=NOTS, ts_begin=TS_BEGIN, max_initial_timestamp_index=None,
                 timestamp_rules=True, suppress_bitmap=model._suppress_bitmap(SUPPRESS, torch.device(DEV)))
    ids = model.greedy_decode_window(enc_dev, prompt.to(DEV), 3 + 16, rules)
    _check_greedy(ids, ref_ids, ref_lg, 3, dm)
    assert out["sequences"].shape[0] == 2
    # three windows per recording (second one shorter): the cached long-form loop returns exactly what the uncached one does
    F2 = 2 * dm.T
    more = [torch.from_numpy(synth.make_features(f"g{k}", 2, dm.n_mels, F2)) for k in (1, 3)]
    long_feats = torch.cat([feats] + more, dim=-1).to(DEV)
    long_stno = torch.cat([stno] + [torch.from_numpy(synth.make_stno(f"g{k}", 2, dm.T, "soft")) for k in (1, 3)], dim=-1).to(DEV)
    attn = torch.ones(2, 3 * F2, dtype=torch.long)
    attn[1, F2 + 31:] = 0
    kw = dict(attention_mask=attn.to(DEV), stno_mask=long_stno, enrollments=enr_d, forced_decoder_ids=prompt, return_segments=True)
    calls = []
    enc = model.get_encoder()
    orig = enc._forward_inference

    def spy(*a, **k):
        calls.append(a[9] is not None if len(a) > 9 else k.get("enrollment_kv") is not None)
        return orig(*a, **k)
    enc._forward_inference = spy
    try:
        model.cache_enrollment_kv = True
        with_cache = model.generate(long_feats, **kw)
        used = list(calls)
        calls.clear()
        model.cache_enrollment_kv = False
        without = model.generate(long_feats, **kw)
    finally:
        enc._forward_inference = orig
        model.cache_enrollment_kv = True
    assert used[0] is False and all(used[1:]) and len(used) >= 2 and not any(calls)
    assert torch.equal(with_cache["sequences"], without["sequences"])
    try:  # speculative next-window encoding on top of the enrollment cache: same result
        model.speculate_next_window = True
        spec = model.generate(long_feats, **kw)
    finally:
        model.speculate_next_window = False
    print("SE-DiCoW speculation:", model.speculation_stats)
    assert torch.equal(with_cache["sequences"], spec["sequences"])
    assert [[s["tokens"].tolist() for s in r] for r in with_cache["segments"]] == \
           [[s["tokens"].tolist() for s in r] for r in without["segments"]]
