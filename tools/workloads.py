"""The secondary workloads of BASELINE.json (configs[2], [3], [4]) as functions, shared by bench.py (which carries them in
its one JSON line under "secondary", each with its own bracketed clock sample) and by the tools/bench_*.py command lines.

  finetune_step      configs[2]  DiCoW-v3 fine-tune step (enc + dec + CTC, bf16 operands / fp32 master weights), B = 8 / GPU,
                                 gradient all-reduce (parallel.GradientExchange: per-layer flat buckets on a side stream)
                                 reference: scripts/submit_slurm.sh:34, configs/train/dicow_v3.yaml:56-66, src/train.py:227-259
  ctc_pretrain_step  configs[4]  CTC encoder pre-train step, B = 16 / GPU (src/pretrain_encoder.py:42-51)
  se_dicow_greedy    configs[3]  SE-DiCoW (FDDT + enrollment cross-attention) greedy decode, B = 16 windows
                                 (configs/decode/se_dicow_greedy.yaml, src/models/dicow/generation.py:707-782)

Timing rules as bench.py: CUDA events on the launching stream, barrier + synchronize on both sides, >= 3 warm-up steps, MAX
over ranks, inputs rotate over 3 batches and the per-step working set (GBs of activations) is far above the 126 MB L2."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SOT, LANG, TASK, EOS, TS_BEGIN, N_TS = 50258, 50259, 50360, 50257, 50365, 1501
# fwd 2273.8 (encoder) + 92.0 (CTC head) + ~84 (decoder at S=64 incl. cross K/V and proj_out) GFLOP / utt;
# backward = 2x forward for trained parts; decoder wgrads skipped when frozen
GFLOP_FINETUNE = 3 * 2273.8 + 3 * 92.0 + 2 * 84.0
GFLOP_CTC_PRETRAIN = 2273.8 + 3 * 92.0
GFLOP_SE_ENCODER = 3577.0  # SE-DiCoW encoder forward per target utterance (8 SCB layers, 2 streams for 8 layers)


def peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


class WhisperIds:
    """the token-id facts of the multilingual Whisper tokenizer the loss needs (no tokenizer files offline;
    export_sources/generation_config.json pins the ids)"""
    prefix_tokens = [SOT, LANG, TASK]
    pad_token_id = EOS

    def get_vocab(self):
        return {f"<|{0.02 * i:.2f}|>": TS_BEGIN + i for i in range(N_TS)}


def turbo_config(**over):
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    kw = dict(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, encoder_attention_heads=20,
              decoder_layers=4, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120,
              max_source_positions=1500, max_target_positions=448, use_fddt=True, use_pre_pos_fddt=True,
              fddt_is_diagonal=True, non_target_fddt_value=0.5, fddt_init="suppressive", ctc_weight=0.0,
              activation_function="gelu")
    kw.update(over)
    return DiCoWConfig(**kw)


def train_config(ctc_weight=0.3):
    return turbo_config(ctc_weight=ctc_weight, additional_self_attention_layer=True, pre_ctc_sub_sample=True,
                        pad_token_id=EOS, eos_token_id=EOS, bos_token_id=EOS, decoder_start_token_id=SOT)


def perturb_(enc, gen_device):
    """Move FDDT / LayerNorm parameters off their identity init -- seeded."""
    g = torch.Generator(device=gen_device).manual_seed(1234)
    with torch.no_grad():
        for name, p in enc.named_parameters():
            if "fddt" in name:
                if name.endswith("weight"):
                    p.copy_(torch.rand(p.shape, generator=g, device=gen_device) + 0.5)
                else:
                    p.copy_(torch.randn(p.shape, generator=g, device=gen_device) * 0.1)
            elif "layer_norm.weight" in name:
                p.copy_(torch.rand(p.shape, generator=g, device=gen_device) * 0.4 + 0.8)
            elif "embed_positions" in name:
                p.copy_(torch.randn(p.shape, generator=g, device=gen_device) * 0.1)


def make_inputs(B, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    feats = (torch.randn(B, 128, 3000, generator=g) * 0.4 - 0.3).clamp_(-1.0, 1.5)
    stno = torch.softmax(3.0 * torch.randn(B, 4, 1500, generator=g), dim=1)
    if pin:
        feats, stno = feats.pin_memory(), stno.pin_memory()
    return feats.to(device), stno.to(device)


def make_train_batch(B, S, seed, dev):
    g = torch.Generator().manual_seed(seed)
    feats = (torch.randn(B, 128, 3000, generator=g) * 0.4 - 0.3).clamp_(-1.0, 1.5)
    stno = torch.softmax(3.0 * torch.randn(B, 4, 1500, generator=g), dim=1)
    labels = torch.full((B, S), -100, dtype=torch.int64)
    for b in range(B):
        n = int(S * (0.6 + 0.4 * torch.rand((), generator=g)))
        row = [LANG, TASK, TS_BEGIN] + torch.randint(0, 50257, (n - 6,), generator=g).tolist() + [TS_BEGIN + 100 + b, EOS]
        labels[b, :len(row)] = torch.tensor(row[:S])
    upp = labels.clone()
    flip = (torch.rand(B, S, generator=g) < 0.05) & (labels >= 0) & (labels < 50257)
    upp[flip] = (upp[flip] + 7) % 50257
    return feats.to(dev), stno.to(dev), labels.to(dev), upp.to(dev)


def _barrier():
    from ts_asr_whisper_b200 import parallel
    parallel.barrier()
    torch.cuda.synchronize()


def _timed(fn, steps, dev):
    """K calls of fn(i) between barrier + synchronize, CUDA events on the current stream; ms = MAX over ranks"""
    from ts_asr_whisper_b200 import parallel
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    _barrier()
    return parallel.max_over_ranks([e0.elapsed_time(e1)], dev)[0]


def train_step(workload: str, dev, rank: int, world: int, steps: int = 5, warmup: int = 3, batch: int = 0,
               label_len: int = 64, overlap: bool = True, sampler=None) -> dict:
    """One of the two training workloads.  Returns a dict for the JSON line.  At world > 1 the step is timed twice: with
    the gradient exchange (the number reported) and with it switched off (each rank stepping on its local gradients), so that
    exposed_comm_ms = step(exchange) - step(local) is measured in the same process on the same clocks."""
    from ts_asr_whisper_b200 import ops, parallel, training
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    fine = workload == "finetune"
    B = batch or (8 if fine else 16)
    torch.manual_seed(1234)
    with torch.device(dev):
        model = DiCoWForConditionalGeneration(train_config())
    model.tie_weights()
    model.set_tokenizer(WhisperIds())
    model.train()
    enc = model.get_encoder()
    head = ("model.encoder.additional_self_attention_layer", "model.encoder.subsample_conv", "model.encoder.lm_head")
    for n, p in model.named_parameters():
        if fine:  # decoder frozen, sinusoidal positions frozen (the recipe: 720 M trainable parameters)
            p.requires_grad_(n.startswith("model.encoder.") and "embed_positions" not in n)
        else:  # src/pretrain_encoder.py:42-51: everything frozen but the CTC head
            p.requires_grad_(n.startswith(head))
    params = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in params)
    if os.environ.get("DICOW_TORCH_ADAMW") == "1":
        opt, opt_name = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.0, fused=True), "torch AdamW(fused=True)"
    else:  # what containers.get_optimizer returns: torch.optim.AdamW's update in one launch (optim.py)
        from ts_asr_whisper_b200.optim import AdamW
        opt, opt_name = AdamW(params, lr=1e-5, weight_decay=0.0), "ts_asr_whisper_b200.optim.AdamW (single launch)"
    exchange = parallel.GradientExchange()
    batches = [make_train_batch(B, label_len, 100 + 10 * rank + i, dev) for i in range(3)]
    state = {"exchange": True, "loss": None}

    def step(i):
        feats, stno, labels, upp = batches[i % 3]
        opt.zero_grad(set_to_none=True)
        training.gradient_exchange = exchange if (state["exchange"] and overlap) else None
        if fine:
            loss = model(feats, stno_mask=stno, labels=labels, upp_labels=upp).loss
        else:
            out = enc(feats, stno_mask=stno, return_logits=True)
            lab = labels[:, 3:].clone()  # src/utils/trainers.py:76-103: prompt tokens stripped, eos -> -100
            lab[lab == EOS] = -100
            loss = enc.get_loss(out.logits, lab)
        loss.backward()
        if state["exchange"] and not overlap:
            parallel.allreduce_gradients(params)
        opt.step()
        state["loss"] = loss.detach()

    try:
        for i in range(max(3, warmup)):
            step(i)
        torch.cuda.reset_peak_memory_stats()
        c0, b0 = exchange.n_collectives, exchange.bytes
        l0 = ops.launch_count
        m0 = sampler.mark() if sampler else 0
        ms = _timed(step, steps, dev)
        m1 = sampler.mark() if sampler else 0
        launches = ops.launch_count - l0
        n_coll, n_bytes = (exchange.n_collectives - c0) // steps, (exchange.bytes - b0) // steps
        ms_local = None
        if world > 1:
            state["exchange"] = False
            step(0)
            ms_local = _timed(step, steps, dev)
            state["exchange"] = True
    finally:
        training.gradient_exchange = None
    pk = peaks()
    peak_tf = pk.get("bf16_tflops_sustained") or 1400.0
    gf = GFLOP_FINETUNE if fine else GFLOP_CTC_PRETRAIN
    value = world * B * steps / (ms * 1e-3)
    tf = value / world * gf / 1e3
    out = {
        "workload": ("BASELINE configs[2]: DiCoW-v3 fine-tune step (enc+dec+CTC, decoder frozen), large-v3-turbo" if fine else
                     "BASELINE configs[4]: CTC encoder pre-train step (CTC head trained), large-v3-turbo"),
        "value": value, "unit": "utt/s", "ms_per_step": ms / steps, "steps": steps, "warmup": max(3, warmup),
        "batch_per_gpu": B, "label_len": label_len, "n_gpus": world, "trainable_params": n_train,
        "optimizer": opt_name + ", fp32 master weights", "dtype": "bf16",
        "gflop_per_utt": gf, "tflops_per_gpu": tf,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                     "floor_ms": B * gf / peak_tf, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"
                     if pk else "fallback"},
        "exchange": {"mode": "overlapped per-layer flat fp32 buckets, NCCL all-reduce(AVG) on a side stream" if overlap
                     else "after backward", "collectives_per_step": n_coll, "allreduce_bytes_per_step": n_bytes,
                     "ms_per_step_local_gradients_only": (ms_local / steps) if ms_local else None,
                     "exposed_comm_ms": ((ms - ms_local) / steps) if ms_local else 0.0},
        "gpu_launches": launches, "loss": float(state["loss"]),
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }
    if sampler:
        out["clocks"] = sampler.window(m0, m1)
    del model, opt, batches, params
    training._scalar_cache.clear()
    torch.cuda.empty_cache()
    return out


def se_dicow_model(dev, ctc_weight: float = 0.0):
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    cfg = turbo_config(use_enrollments=True, scb_layers=8, pad_token_id=EOS, eos_token_id=EOS, decoder_start_token_id=SOT)
    if ctc_weight > 0:  # the CTC head of the recipe (configs/base.yaml:5,19-21)
        cfg.ctc_weight, cfg.additional_self_attention_layer, cfg.pre_ctc_sub_sample = 0.3, True, True
    torch.manual_seed(4321)
    with torch.device(dev):
        model = DiCoWForConditionalGeneration(cfg)
    perturb_(model.get_encoder(), dev)
    with torch.no_grad():
        for blk in model.get_encoder().ca_enrolls:
            blk.cae.cross_gate.gate.fill_(0.5)
    return model.eval()


def decode_rules(model, dev):
    return dict(eos=EOS, pad=EOS, no_timestamps=50364, ts_begin=TS_BEGIN, max_initial_timestamp_index=None,
                timestamp_rules=True, suppress_bitmap=model._suppress_bitmap([EOS, 220, 50256], dev))


def decode_floor_bytes(cfg, B: int, T: int = 1500) -> int:
    """HBM bytes one greedy step must move: the decoder layers' weights + proj_out once, and every
    window's cross-attention K/V"""
    d, L, ffn = cfg.d_model, cfg.decoder_layers, cfg.decoder_ffn_dim
    wbytes = L * (8 * d * d + 2 * d * ffn) * 2 + cfg.vocab_size * d * 2
    return wbytes + B * L * T * 2 * d * 2


def se_dicow_greedy(dev, rank: int, world: int, batch: int = 16, new_tokens: int = 128, reps: int = 5, sampler=None,
                    beams: int = 1, ctc_weight: float = 0.0, graphs: bool = True, fused=True) -> dict:
    """BASELINE configs[3]: one pass = DiCoWEncoder.forward over B target + B enrollment windows (8 speaker
    communication blocks) + cross-K/V projection + prompt + ``new_tokens`` greedy tokens per window through the CUDA-graphed
    decode step (EOS suppressed so every run decodes the same number of tokens).  Also times the decode steps alone against
    the HBM floor of a step."""
    from ts_asr_whisper_b200 import ops, parallel
    model = se_dicow_model(dev, ctc_weight)
    model.use_cuda_graphs = graphs
    model.fused_decode_step = fused
    cfg = model.config
    B = batch
    batches = []
    for i in range(3):
        f, s = make_inputs(2 * B, 20 + 7 * rank + i, device=dev)
        batches.append((f[:B], s[:B], {"input_features": f[B:], "stno_mask": s[B:]}))
    prompt = torch.tensor([[SOT, LANG, TASK]] * B, device=dev)
    rules = decode_rules(model, dev)
    n = 3 + new_tokens
    enc = model.get_encoder()

    def decode(hidden):
        ctc = None
        if ctc_weight > 0:
            ctc = {"logits": model.get_enc_logits(hidden), "weight": ctc_weight, "prefix_len": 3, "bos": SOT}
        if beams > 1:  # configs/decode/se_dicow_beam_joint.yaml: 5 beams, ctc 0.2, length_penalty 0.1
            return model.beam_decode_window(hidden, prompt, n, rules, num_beams=beams, length_penalty=0.1, ctc=ctc)
        return model.greedy_decode_window(hidden, prompt, n, rules, ctc=ctc)

    ids = None
    with torch.no_grad():
        for i in range(3):
            f, s, e = batches[i % 3]
            ids = decode(enc(f, stno_mask=s, enrollments=e).last_hidden_state)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * reps)]
        l0 = ops.launch_count
        m0 = sampler.mark() if sampler else 0
        _barrier()
        for i in range(reps):
            f, s, e = batches[i % 3]
            ev[3 * i].record()
            hidden = enc(f, stno_mask=s, enrollments=e).last_hidden_state
            ev[3 * i + 1].record()
            ids = decode(hidden)
            ev[3 * i + 2].record()
        _barrier()
        m1 = sampler.mark() if sampler else 0
        launches = (ops.launch_count - l0) // reps
        t_enc = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(reps)) / reps
        t_all = sum(ev[3 * i].elapsed_time(ev[3 * i + 2]) for i in range(reps)) / reps
    ms_all, ms_enc = parallel.max_over_ranks([t_all, t_enc], dev)
    ms_dec = ms_all - ms_enc
    pk = peaks()
    hbm = pk.get("hbm_gbs") or 6650.0
    peak_tf = pk.get("bf16_tflops_sustained") or 1400.0
    bytes_step = decode_floor_bytes(cfg, B * beams if beams > 1 else B)
    floor_ms = bytes_step / (hbm * 1e9) * 1e3
    ms_step = ms_dec / (n - 1)  # prompt (2 steps) + new tokens; includes the per-window cross-K/V projection (4 GEMMs)
    out = {
        "workload": "BASELINE configs[3]: SE-DiCoW (FDDT + enrollment cross-attn, 8 SCB layers) greedy decode, "
                    f"batch={B} windows + {B} enrollment windows, large-v3-turbo",
        "value": world * B / (ms_all * 1e-3), "unit": "windows/s", "tokens_per_s": world * B * new_tokens / (ms_all * 1e-3),
        "n_gpus": world, "batch_per_gpu": B, "beams": beams, "ctc_weight": ctc_weight, "new_tokens_per_window": new_tokens,
        "reps": reps, "ms_per_batch": ms_all, "ms_encoder": ms_enc, "ms_decode": ms_dec, "ms_per_decode_step": ms_step,
        "decode_steps": n - 1, "cuda_graphs": bool(model.use_cuda_graphs), "gpu_launches_per_batch": launches,
        "encoder_tflops": B * GFLOP_SE_ENCODER / ms_enc,
        "roofline": {"bound": "hbm", "kernel": "decode step (all kernels of one token step)", "achieved":
                     bytes_step / (ms_step * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": floor_ms / ms_step, "bytes_per_step": bytes_step, "floor_ms": floor_ms,
                     "encoder_frac_of_sustained_tensor_peak": B * GFLOP_SE_ENCODER / ms_enc / peak_tf,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if pk else "fallback"},
        "dtype": "bf16", "generated_tail": ids[0, -4:].tolist(),
    }
    if sampler:
        out["clocks"] = sampler.window(m0, m1)
    model.clear_decode_cache() if hasattr(model, "clear_decode_cache") else None
    del model, batches, ids
    torch.cuda.empty_cache()
    return out


def longform_speculation(dev, rank: int, world: int, batch: int = 16, windows: int = 6, new_tokens: int = 64, reps: int = 2,
                         sampler=None) -> dict:
    """long-form generate() over synthetic multi-window recordings (large-v3-turbo + FDDT, timestamps
    off so that every window advances by a full window, EOS suppressed: every window decodes ``new_tokens`` tokens), the plain
    seek loop against ``speculate_next_window`` (the window at seek + 3000 encoded on part of the SMs under the decode steps
    of the current one), interleaved in one process; the sequences must be identical."""
    from ts_asr_whisper_b200 import parallel
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    torch.manual_seed(4321)
    with torch.device(dev):
        model = DiCoWForConditionalGeneration(turbo_config(pad_token_id=EOS, eos_token_id=EOS,
                                                           decoder_start_token_id=SOT)).eval()
    perturb_(model.get_encoder(), dev)
    B, W = batch, windows
    g = torch.Generator().manual_seed(5 + rank)
    feats = (torch.randn(B, 128, 3000 * W, generator=g) * 0.4 - 0.3).clamp_(-1.0, 1.5).to(dev)
    stno = torch.softmax(3.0 * torch.randn(B, 4, 1500 * W, generator=g), dim=1).to(dev)
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = 50364, EOS, EOS
    gc.suppress_tokens, gc.begin_suppress_tokens = [EOS, 220, 50256], None
    gc.return_timestamps, gc.max_new_tokens, gc.num_beams = False, new_tokens, 1
    prompt = torch.tensor([[SOT, LANG, TASK, 50364]] * B)
    kw = dict(stno_mask=stno, forced_decoder_ids=prompt, return_segments=True)

    def run(spec):
        model.speculate_next_window = spec
        _barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model.generate(feats, **kw)
        e1.record()
        _barrier()
        return e0.elapsed_time(e1), out["sequences"], dict(model.speculation_stats)

    with torch.no_grad():
        run(False), run(True)  # graphs, lazy loads
        m0 = sampler.mark() if sampler else 0
        ms = {False: [], True: []}
        ref, identical, stats = None, True, None
        for _ in range(reps):
            for spec in (False, True):
                t, seq, st = run(spec)
                ms[spec].append(t)
                ref = seq if ref is None else ref
                identical = identical and bool(torch.equal(ref, seq))
                stats = st if spec else stats
        m1 = sampler.mark() if sampler else 0
    med = {k: sorted(v)[len(v) // 2] for k, v in ms.items()}
    plain, spec = parallel.max_over_ranks([med[False], med[True]], dev)
    out = {
        "workload": f"SURVEY 8(f).3: long-form generate(), {B} recordings x {W} windows per GPU, {new_tokens} tokens / window, "
                    "large-v3-turbo + FDDT, timestamps off",
        "value": world * B * W / (spec * 1e-3), "unit": "windows/s", "n_gpus": world,
        "ms_per_call_plain_seek_loop": plain, "ms_per_call_speculative": spec, "speedup": plain / spec,
        "windows_per_s_plain": world * B * W / (plain * 1e-3), "speculation_sms": model.speculation_sms if model.speculation_sms is not None else "auto",
        "speculation_stats": stats, "identical_sequences": identical, "reps": reps, "dtype": "bf16",
    }
    if sampler:
        out["clocks"] = sampler.window(m0, m1)
    model.clear_decode_cache()
    del model, feats, stno
    torch.cuda.empty_cache()
    return out
