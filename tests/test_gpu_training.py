"""GPU parity of the training step (hand-scheduled backward on the C-ABI kernels) against torch autograd through the
fp32 oracle (oracle/dicow_oracle.py) on the same seeded weights / inputs.

Both BASELINE training configurations are covered at miniature and whisper-tiny dimensions:
  * CTC encoder pre-training (configs[4]): everything frozen but the CTC head (src/pretrain_encoder.py:42-51), loss through
    ``encoder(return_logits=True)`` + ``encoder.get_loss`` exactly as src/utils/trainers.py:76-103 calls them;
  * DiCoW fine-tuning (configs[2]): ``DiCoWForConditionalGeneration.forward(labels, upp_labels)`` -> 0.7 CE + 0.3 CTC, with
    the decoder frozen (the recipe) and with every parameter trainable.

SE-DiCoW (enrollment streams through the speaker communication blocks, src/models/dicow/layers.py:145-193) is covered by
``test_se_dicow_finetune_step`` / ``test_se_dicow_ctc_pretrain_step`` (gate off its zero init so every SCB gradient is live).

Tolerance: the path computes with bf16 operands (fp32 accumulation / residual stream / statistics) against an fp32
reference, so per-parameter gradients are required to agree to max |err| <= GRAD_TOL x max |ref| (north_star: 2e-2 bf16;
gradients accumulate one bf16 rounding per layer on the way back and the max-norm is taken over up to 10^6 entries of a
tensor, so the bound used here is 5e-2; measured worst case 4.2e-2 on a decoder k_proj weight of the 2-layer miniature,
whose few-term sums are the worst conditioned) and cosine >= 0.999.  The full-size model (tests/test_gpu_turbo_parity.py:
32 + 4 layers at d = 1280) is held to per-tensor-class bounds of 3.5e-2 / cosine 0.9995 (FDDT tables 1e-1)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import dicow_oracle as orc
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 5e-2
# The SCB gate gradient is ONE scalar = sum over B*T*d signed terms G*upd.  At miniature dims the sum is well conditioned
# (|sum| ~ 6 x the l2 norm of its terms) and the 5e-2 bound on the sum itself applies (scale = |sum|); at whisper-tiny
# dims it cancels to 0.3 % of sum|terms| (measured on the fp32 oracle: 0.098 vs 31.8, l2 norm 0.13), where bf16 errors of
# the heavy-tailed stream gradient G (bounded relative to max|G|, not per term) reach ~0.1 x the l2 norm.  The bound is
# therefore taken against max(|sum|, l2 norm of the terms); tests/test_gpu_backward.py::test_gate_bwd pins the kernel's
# arithmetic exactly on identical inputs.
GATE_TOL = 0.25
EOS, SOT, LANG, TASK, NOTS, TS_BEGIN, N_TS = 257, 258, 259, 260, 261, 262, 38
HEAD = ("model.encoder.additional_self_attention_layer", "model.encoder.additional_layer", "model.encoder.subsample_conv",
        "model.encoder.lm_head")


class FakeTokenizer:
    prefix_tokens = [SOT, LANG, TASK]
    pad_token_id = EOS

    def get_vocab(self):
        v = {f"<|{0.02 * i:.2f}|>": TS_BEGIN + i for i in range(N_TS)}
        v["Ġ"] = 220
        return v


def _build(dm: synth.Dims):
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    params = synth.make_params(dm)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    model.tie_weights()
    model = model.to(DEV).train()
    p = orc.to_torch(params, device=DEV)
    p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]  # tied (src/train.py:109-113)
    return model, p


def _compare(model, p, names, label, scalar_refs=None, tol=None):
    """scalar_refs: {name: (gradient, scale)} for scalar parameters whose gradient is a sum of many signed terms (the
    SCB gate): the error is measured against the l2 norm of the terms (the random-walk size of the sum) when the sum
    itself cancels below it -- a relative error on a cancelling sum measures the conditioning, not the kernels."""
    worst = 0.0
    named = dict(model.named_parameters())
    checked = 0
    for n in names:
        got = named[n].grad
        if p[n].grad is None and not (scalar_refs and n in scalar_refs):
            # a module the forward does not use (additional_self_attention_layer next to additional_layer,
            # encoder.py:88-100): autograd leaves .grad unset in the reference, and so does the B200 step
            assert got is None, f"{label}: gradient for {n}, which the oracle's forward never touches"
            continue
        assert got is not None, f"{label}: no gradient for {n}"
        if scalar_refs and n in scalar_refs:
            ref, scale = scalar_refs[n]
            err = abs(got.item() - ref) / scale
            print(f"{label}: {n}: got {got.item():.4e} ref {ref:.4e} cancellation scale {scale:.4e} err/scale {err:.3e}")
            tol = GRAD_TOL if abs(ref) >= scale else GATE_TOL  # well conditioned: the usual bound on the sum itself
            assert err < tol, f"{label}: {n}: got {got.item():.4e} ref {ref:.4e} scale {scale:.4e}"
            checked += 1
            continue
        ref = p[n].grad
        assert ref is not None, f"{label}: oracle has no gradient for {n}"
        scale = ref.abs().max().item()
        if scale < 1e-12:
            assert got.abs().max().item() < 1e-6, n
            continue
        err = (got.float() - ref).abs().max().item() / scale
        cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.flatten(), dim=0).item()
        worst = max(worst, err)
        assert err < (tol or GRAD_TOL) and cos > 0.999, f"{label}: {n}: rel err {err:.3e}, cos {cos:.5f}"
        checked += 1
    print(f"{label}: {checked} parameter gradients, worst rel err {worst:.3e}")
    return worst


def _inputs(dm, B, tag):
    feats = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T)).to(DEV)
    stno = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=5)).to(DEV)
    return feats, stno


MINI = synth.Dims(**{**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0})
# whisper-tiny widths (d 384, 6 heads) with a short window so the fp32 autograd reference stays small
TINY_SHORT = dataclasses.replace(synth.WHISPER_TINY, T=200, vocab=1000, max_target=64, pad_token_id=257, eos_token_id=257,
                                 decoder_start_token_id=258, enc_layers=2, dec_layers=2)


# additional_layer=True: a whole encoder layer in front of the CTC head instead of the bare self-attention (encoder.py:17-18, 88-93)
MINI_LAYER = dataclasses.replace(MINI, additional_layer=True)
TINY_LAYER = dataclasses.replace(TINY_SHORT, additional_layer=True)


@pytest.mark.parametrize("dm,B", [(MINI, 2), (TINY_SHORT, 3), (MINI_LAYER, 2), (TINY_LAYER, 3)],
                         ids=["mini", "tiny-short", "mini-additional-layer", "tiny-short-additional-layer"])
@pytest.mark.parametrize("body", [False, True], ids=["head-only", "all-encoder"])
def test_ctc_pretrain_step(dm, B, body):
    """configs[4]: encoder(return_logits=True) -> get_loss -> backward, as CustomTrainerEncoder.compute_loss does"""
    model, p = _build(dm)
    enc = model.get_encoder()
    for n, q in model.named_parameters():
        q.requires_grad_(n.startswith(HEAD) or (body and n.startswith("model.encoder.") and "embed_positions" not in n))
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    assert trainable
    feats, stno = _inputs(dm, B, "tr0")
    rng = np.random.default_rng(5)
    L = min(12, dm.T // 8 - 2)
    labels = torch.full((B, L), -100, dtype=torch.int64)
    for b in range(B):
        n = L - 2 * b
        labels[b, :n] = torch.from_numpy(rng.integers(0, 200, size=n))
    labels = labels.to(DEV)
    out = enc(feats, stno_mask=stno, return_logits=True)
    assert out.logits.requires_grad
    loss = enc.get_loss(out.logits, labels)
    loss.backward()
    for n in trainable:
        p[n].requires_grad_(True)
    ref_logits = orc.encoder_forward(p, dm, feats, stno, return_logits=True)
    ref_loss = orc.ctc_loss(ref_logits, labels)
    ref_loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    _compare(model, p, trainable, f"ctc-pretrain[{'body' if body else 'head'}]")
    frozen = [q for n, q in model.named_parameters() if not q.requires_grad]
    assert all(q.grad is None for q in frozen)


@pytest.mark.parametrize("dm,B,S", [(MINI, 2, 11), (TINY_SHORT, 3, 24)], ids=["mini", "tiny-short"])
@pytest.mark.parametrize("mode", ["decoder-frozen", "all", "fddt-only"])
def test_finetune_step(dm, B, S, mode):
    _finetune_case(dm, B, S, mode, "tr1")


# FDDT(bias_only=True) (src/models/dicow/FDDT.py:10-13, 43-51): the per-class parameter is a bias vector only.
# The "tr1" draw is ill-conditioned for these tables at whisper-tiny widths: on the fp32 CPU oracle alone, rounding the
# weight matrices to bf16 moves d(loss)/d(encoder output) by 28 % and conv1.weight's gradient by 11 % (loss 22.2329 ->
# 22.2296), against 0.9 % / 0.8 % for "tr2" (tests/cond_cpu.py) -- a bf16 path cannot be held to 5 % there, so these
# cases use the well-conditioned draw.  The decoder does not see the FDDT variant, so its parameters stay frozen here
# ("all" is covered above; on "tr2" the near-uniform cross-attention makes the q/k weight gradients of BOTH variants
# cancel to 6-8 % bf16 noise with cos 0.9995, measured on the diagonal variant too).
MINI_BIAS = dataclasses.replace(MINI, fddt_bias_only=True)
TINY_BIAS = dataclasses.replace(TINY_SHORT, fddt_bias_only=True)


@pytest.mark.parametrize("dm,B,S", [(MINI_BIAS, 2, 11), (TINY_BIAS, 3, 24)], ids=["mini", "tiny-short"])
@pytest.mark.parametrize("mode", ["decoder-frozen", "fddt-only"])
def test_finetune_step_bias_only_fddt(dm, B, S, mode):
    _finetune_case(dm, B, S, mode, "tr2")


# full-matrix FDDT: a d x d CustomLinear per class (src/models/dicow/layers.py:7-47, FDDT.py:52-62)
MINI_FULL = dataclasses.replace(MINI, fddt_is_diagonal=False)
TINY_FULL = dataclasses.replace(TINY_SHORT, fddt_is_diagonal=False)


@pytest.mark.parametrize("dm,B,S,tag", [(MINI_FULL, 2, 11, "tr1"), (TINY_FULL, 3, 24, "tr1")], ids=["mini", "tiny-short"])
@pytest.mark.parametrize("mode", ["decoder-frozen", "fddt-only"])
def test_finetune_step_full_matrix_fddt(dm, B, S, tag, mode):
    _finetune_case(dm, B, S, mode, tag)


@pytest.mark.parametrize("over", [{"apply_fddt_to_n_layers": 1}, {"use_pre_pos_fddt": False}], ids=["first-layer-only", "no-pre-pos"])
def test_finetune_step_fddt_placement_variants(over):
    """FDDT in the first layer only / no FDDT in front of the positions: gradients of what is left (encoder.py:45-47, 62)"""
    _finetune_case(dataclasses.replace(TINY_SHORT, enc_layers=3, **over), 3, 24, "decoder-frozen", "tr2")


def test_finetune_step_remove_timestamps_from_ctc():
    """remove_timestamps_from_ctc=True (encoder.py:76, 111-113): the CTC branch is trained on the labels below the first task
    token only (vocab 1700 -> ids < 193); loss and gradients vs the oracle, whose filter is pinned to the reference live"""
    _finetune_case(dataclasses.replace(MINI, remove_timestamps_from_ctc=True, vocab=1700), 2, 11, "decoder-frozen", "tr1")


@pytest.mark.parametrize("dm,B,S", [(MINI_LAYER, 2, 11), (TINY_LAYER, 3, 24)], ids=["mini", "tiny-short"])
def test_finetune_step_additional_layer(dm, B, S):
    """the CTC branch's gradient reaches the encoder body through the whole extra layer (accumulated into the decoder's)"""
    _finetune_case(dm, B, S, "decoder-frozen", "tr1")


def _finetune_case(dm, B, S, mode, tag):
    """configs[2]: loss = 0.7 soft-label CE + 0.3 CTC through DiCoWForConditionalGeneration.forward, loss.backward()"""
    model, p = _build(dm)
    model.set_tokenizer(FakeTokenizer())
    for n, q in model.named_parameters():
        if mode == "decoder-frozen":
            q.requires_grad_(n.startswith("model.encoder.") and "embed_positions" not in n)
        elif mode == "fddt-only":  # the recipe's first 2000 steps train the FDDT tables only (prefixes_to_preheat)
            q.requires_grad_("fddt" in n)
        else:
            q.requires_grad_("encoder.embed_positions" not in n)
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    feats, stno = _inputs(dm, B, tag)
    labels = torch.from_numpy(synth.make_labels(tag, B, S, min(dm.vocab, 300), EOS, TS_BEGIN, prefix=(LANG, TASK)))
    upp = labels.clone()
    upp[:, ::3] = torch.where(upp[:, ::3] >= 0, (upp[:, ::3] + 3) % 250, upp[:, ::3])
    labels, upp = labels.to(DEV), upp.to(DEV)
    out = model(feats, stno_mask=stno, labels=labels, upp_labels=upp)
    assert out.loss.requires_grad
    (2.0 * out.loss).backward()  # a non-unit upstream gradient (gradient accumulation / loss scaling) must flow through
    for n in trainable:
        p[n].requires_grad_(True)
    ref_loss, ref_logits, _ = orc.model_forward(p, dm, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK),
                                                ts_begin=TS_BEGIN, n_ts=N_TS)
    (2.0 * ref_loss).backward()
    torch.cuda.synchronize()
    assert abs(out.loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    err = ((out.logits.float() - ref_logits).abs().max() / ref_logits.abs().max()).item()
    assert err < 2e-2, f"logits rel err {err:.3e}"
    _compare(model, p, trainable, f"finetune[{mode}]")


@pytest.mark.parametrize("dm,B,S", [(MINI, 2, 11), (TINY_SHORT, 3, 24)], ids=["mini", "tiny-short"])
def test_finetune_step_with_decoder_lora(dm, B, S):
    """use_lora (src/models/containers.py:69-90): rank-16 adapters on the decoder projections train together with the
    encoder, the decoder's base weights stay frozen.  Loss, logits and the gradients of every lora_A / lora_B (and of the
    encoder) against torch autograd through the oracle's restatement y = x W^T + b + (alpha / r) (x A^T) B^T."""
    model, p = _build(dm)
    model.set_tokenizer(FakeTokenizer())
    model.add_lora(r=16, lora_alpha=32, seed=3)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():  # B off its zero init so that the gradients of A are live too
        for n, q in model.named_parameters():
            if n.endswith("lora_B"):
                q.copy_((torch.randn(q.shape, generator=g) * 0.02).to(DEV))
    for n, q in model.named_parameters():
        q.requires_grad_("lora_" in n or (n.startswith("model.encoder.") and "embed_positions" not in n))
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    assert sum("lora_" in n for n in trainable) == 20 * dm.dec_layers
    for n, q in model.named_parameters():
        if "lora_" in n:
            p[n] = q.detach().clone().float()
    feats, stno = _inputs(dm, B, "tr1")
    labels = torch.from_numpy(synth.make_labels("tr1", B, S, min(dm.vocab, 300), EOS, TS_BEGIN, prefix=(LANG, TASK))).to(DEV)
    out = model(feats, stno_mask=stno, labels=labels, upp_labels=labels)
    out.loss.backward()
    for n in trainable:
        p[n].requires_grad_(True)
    ref_loss, ref_logits, _ = orc.model_forward(p, dm, feats, stno, labels, labels, ctc_prefix_tokens=(SOT, LANG, TASK),
                                                ts_begin=TS_BEGIN, n_ts=N_TS)
    ref_loss.backward()
    torch.cuda.synchronize()
    assert abs(out.loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    err = ((out.logits.float() - ref_logits).abs().max() / ref_logits.abs().max()).item()
    assert err < 2e-2, f"logits rel err {err:.3e}"
    # adapters: [*, 16] matrices whose entries are sums over few (B * S tokens, or 16) terms; the cross-attention k_proj
    # adapter inherits the cancelling q/k gradient of the near-uniform cross attention of these random models (see the
    # note above test_finetune_step_bias_only_fddt): measured worst case 5.2e-2 (cos 0.99939) -> bound 7e-2, cosine 0.999
    _compare(model, p, trainable, "finetune[lora]", tol=7e-2)
    assert all(q.grad is None for n, q in model.named_parameters() if not q.requires_grad)
    # generate() decodes with the merged weights: merging the adapters must not change the teacher-forced logits
    with torch.no_grad():
        before = model(feats, stno_mask=stno, labels=labels, upp_labels=labels).logits
        model.merge_lora()
        after = model(feats, stno_mask=stno, labels=labels, upp_labels=labels).logits
    assert torch.equal(before, after)


def test_training_step_updates_and_second_step():
    """two optimizer steps: the prepared bf16 weight copies follow the parameter updates (versioned cache), the loss
    moves, and evaluation under no_grad takes the inference path with identical loss"""
    dm = MINI
    model, _ = _build(dm)
    model.set_tokenizer(FakeTokenizer())
    opt = torch.optim.AdamW([q for q in model.parameters() if q.requires_grad], lr=1e-3)
    feats, stno = _inputs(dm, 2, "tr2")
    labels = torch.from_numpy(synth.make_labels("tr2", 2, 9, 300, EOS, TS_BEGIN, prefix=(LANG, TASK))).to(DEV)
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        out = model(feats, stno_mask=stno, labels=labels, upp_labels=labels)
        out.loss.backward()
        opt.step()
        losses.append(out.loss.item())
    assert losses[2] < losses[0], losses
    with torch.no_grad():
        ev = model(feats, stno_mask=stno, labels=labels, upp_labels=labels)
    out = model(feats, stno_mask=stno, labels=labels, upp_labels=labels)
    assert abs(ev.loss.item() - out.loss.item()) < 1e-3 * max(1.0, abs(ev.loss.item()))
    # from the second optimizer step on the prepared weights are refreshed IN PLACE by a captured CUDA graph: what it left
    # behind must be exactly what an eager rebuild from the current parameters gives
    from ts_asr_whisper_b200 import modeling

    def flat(o, pre=""):
        if isinstance(o, torch.Tensor):
            yield pre, o
        elif isinstance(o, dict):
            for k, v in o.items():
                yield from flat(v, f"{pre}.{k}")
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                yield from flat(v, f"{pre}[{i}]")

    enc = model.get_encoder()
    if modeling.prepare_graphs:
        assert enc in modeling._PREPARE_GRAPHS, "the graph-captured refresh was not used"
    snap = {k: v.clone() for k, v in flat(enc.prepare())}
    enc.invalidate_cache()
    fresh = dict(flat(enc.prepare()))
    assert snap.keys() == fresh.keys() and len(snap) > 20
    for k, v in snap.items():
        assert torch.equal(v, fresh[k]), k
    import copy
    copy.deepcopy(model)  # stays copyable (the graph object lives outside the module)


def test_unfreeze_after_fddt_warmup_phase():
    """CustomTrainer.training_step (src/utils/trainers.py:116-139) flips requires_grad in the middle of training: the first
    steps train the FDDT tables only, then everything except the frozen keywords.  The step after the flip must see the
    new trainable set (no stale gradient buckets / tapes) and match the oracle's autograd on it."""
    dm, B, S = TINY_SHORT, 3, 24
    model, p = _build(dm)
    model.set_tokenizer(FakeTokenizer())
    feats, stno = _inputs(dm, B, "tr1")
    labels = torch.from_numpy(synth.make_labels("tr1", B, S, min(dm.vocab, 300), EOS, TS_BEGIN, prefix=(LANG, TASK))).to(DEV)
    for n, q in model.named_parameters():
        q.requires_grad_("fddt" in n)
    model(feats, stno_mask=stno, labels=labels, upp_labels=labels).loss.backward()
    with_grad = {n for n, q in model.named_parameters() if q.grad is not None}
    assert with_grad and all("fddt" in n for n in with_grad)
    model.zero_grad(set_to_none=True)
    frozen_keywords = ("decoder", "embed_positions")  # the flip of trainers.py:121-131
    for n, q in model.named_parameters():
        q.requires_grad_(not any(k in n for k in frozen_keywords))
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    out = model(feats, stno_mask=stno, labels=labels, upp_labels=labels)
    out.loss.backward()
    for n in trainable:
        p[n].requires_grad_(True)
    ref_loss, _, _ = orc.model_forward(p, dm, feats, stno, labels, labels, ctc_prefix_tokens=(SOT, LANG, TASK),
                                       ts_begin=TS_BEGIN, n_ts=N_TS)
    ref_loss.backward()
    assert abs(out.loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    _compare(model, p, trainable, "after-unfreeze")
    assert all(q.grad is None for n, q in model.named_parameters() if not q.requires_grad)


# ---- SE-DiCoW: enrollment streams + speaker communication blocks in the training step -------------------------------
SE_MINI = synth.GOLDEN_MINI  # 3 layers, 2 of them with an SCB, odd T
SE_TINY = dataclasses.replace(TINY_SHORT, enc_layers=3, use_enrollments=True, scb_layers=2)


def _expand_gates(p, names, B, dm):
    """give the oracle one gate entry per element of the gated update (all equal to the scalar), so that autograd returns
    the individual terms of the scalar gate gradient: their sum is the reference, their l2 norm its conditioning scale"""
    gates = [n for n in names if n.endswith("cross_gate.gate")]
    for n in gates:
        p[n] = p[n].detach().reshape(1, 1, 1).expand(B, dm.T, dm.d).clone().requires_grad_(True)
    return gates


def _gate_refs(p, gates):
    out = {}
    for n in gates:
        terms = p[n].grad.double()
        total = terms.sum().item()
        out[n] = (total, max(abs(total), terms.norm().item()))
    return out


def _enrollment_inputs(dm, B, tag):
    return {"input_features": torch.from_numpy(synth.make_features(tag + "e", B, dm.n_mels, 2 * dm.T)).to(DEV),
            "stno_mask": torch.from_numpy(synth.make_stno(tag + "e", B, dm.T, "hard")).to(DEV)}


@pytest.mark.parametrize("dm,B,S", [(SE_MINI, 2, 11), (SE_TINY, 2, 20)], ids=["mini", "tiny-short"])
@pytest.mark.parametrize("mode", ["decoder-frozen", "scb-only"])
def test_se_dicow_finetune_step(dm, B, S, mode):
    """the se_dicow recipe: DiCoWForConditionalGeneration.forward(..., enrollments=...) -> loss.backward()"""
    model, p = _build(dm)
    model.set_tokenizer(FakeTokenizer())
    for n, q in model.named_parameters():
        if mode == "decoder-frozen":
            q.requires_grad_(n.startswith("model.encoder.") and "embed_positions" not in n)
        else:  # configs/train/se_dicow.yaml prefixes_to_preheat: the new blocks first
            q.requires_grad_("ca_enrolls" in n)
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    assert any("ca_enrolls" in n and "cross_gate" in n for n in trainable)
    feats, stno = _inputs(dm, B, "se1")
    enr = _enrollment_inputs(dm, B, "se1")
    labels = torch.from_numpy(synth.make_labels("se1", B, S, min(dm.vocab, 300), EOS, TS_BEGIN, prefix=(LANG, TASK)))
    labels = labels.to(DEV)
    out = model(feats, stno_mask=stno, labels=labels, upp_labels=labels, enrollments=enr)
    assert out.loss.requires_grad
    out.loss.backward()
    for n in trainable:
        p[n].requires_grad_(True)
    gates = _expand_gates(p, trainable, B, dm)
    ref_loss, ref_logits, _ = orc.model_forward(p, dm, feats, stno, labels, labels, enrollments=enr,
                                                ctc_prefix_tokens=(SOT, LANG, TASK), ts_begin=TS_BEGIN, n_ts=N_TS)
    ref_loss.backward()
    torch.cuda.synchronize()
    assert abs(out.loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    err = ((out.logits.float() - ref_logits).abs().max() / ref_logits.abs().max()).item()
    assert err < 2e-2, f"logits rel err {err:.3e}"
    _compare(model, p, trainable, f"se-finetune[{mode}]", _gate_refs(p, gates))
    # the no_grad evaluation (interleaved inference path) agrees with the stacked training path
    with torch.no_grad():
        ev = model(feats, stno_mask=stno, labels=labels, upp_labels=labels, enrollments=enr)
    assert abs(ev.loss.item() - out.loss.item()) < 1e-2 * max(1.0, abs(ev.loss.item()))


def test_se_dicow_ctc_pretrain_step():
    """encoder(return_logits=True, enrollments=...) -> get_loss -> backward with the whole encoder trainable"""
    dm, B = SE_MINI, 2
    model, p = _build(dm)
    enc = model.get_encoder()
    for n, q in model.named_parameters():
        q.requires_grad_(n.startswith("model.encoder.") and "embed_positions" not in n)
    trainable = [n for n, q in model.named_parameters() if q.requires_grad]
    feats, stno = _inputs(dm, B, "se2")
    enr = _enrollment_inputs(dm, B, "se2")
    labels = torch.tensor([[5, 9, 17, 3], [8, 2, -100, -100]], dtype=torch.int64, device=DEV)
    out = enc(feats, stno_mask=stno, return_logits=True, enrollments=enr)
    loss = enc.get_loss(out.logits, labels)
    loss.backward()
    for n in trainable:
        p[n].requires_grad_(True)
    gates = _expand_gates(p, trainable, B, dm)
    ref_logits = orc.encoder_forward(p, dm, feats, stno, enr, return_logits=True)
    ref_loss = orc.ctc_loss(ref_logits, labels)
    ref_loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, abs(ref_loss.item()))
    _compare(model, p, trainable, "se-ctc-pretrain", _gate_refs(p, gates))


def test_adamw_single_launch_matches_torch():
    """ts_asr_whisper_b200.optim.AdamW (one launch over all tensors, csrc/optimizer.cu) against torch.optim.AdamW(fused=True) on
    ragged tensor sizes, two parameter groups (the reference's get_optimizer: a higher learning rate and no weight decay for the
    second one), several steps; state_dict round trip into torch's optimizer"""
    from ts_asr_whisper_b200.optim import AdamW
    g = torch.Generator(device=DEV).manual_seed(3)
    shapes = [(1280, 1280), (5120,), (1, ), (3, 77), (16385,), (4, 1280), (51866, 8), (33,)]
    a = [torch.nn.Parameter(torch.randn(*s, device=DEV, generator=g)) for s in shapes]
    b = [torch.nn.Parameter(x.detach().clone()) for x in a]
    kw = dict(lr=3e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    ours = AdamW([{"params": a[:5]}, {"params": a[5:], "lr": 3e-2, "weight_decay": 0.0}], **kw)
    ref = torch.optim.AdamW([{"params": b[:5]}, {"params": b[5:], "lr": 3e-2, "weight_decay": 0.0}], fused=True, **kw)
    for step in range(4):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, device=DEV, generator=g) * (0.1 if step else 3.0)
            x.grad, y.grad = gr.clone(), gr.clone()
        if step == 2:  # a parameter without a gradient is skipped by both
            a[1].grad = b[1].grad = None
        ours.step()
        ref.step()
    torch.cuda.synchronize()
    for x, y, s in zip(a, b, shapes):
        err = (x - y).abs().max().item() / max(y.abs().max().item(), 1e-12)
        assert err < 2e-6, (s, err)
    for x, y in zip(a, b):
        assert (ours.state[x]["exp_avg_sq"] - ref.state[y]["exp_avg_sq"]).abs().max().item() <= 1e-6 * ref.state[y]["exp_avg_sq"].abs().max().item() + 1e-12
    fresh = torch.optim.AdamW([{"params": b[:5]}, {"params": b[5:]}], **kw)
    fresh.load_state_dict(ours.state_dict())  # same state layout
