#!/usr/bin/env python
"""Training-step benchmarks of the hot path (BASELINE.json configs[2] and configs[4]); bench.py stays the encoder-forward
headline.  One JSON line per run, same timing rules as bench.py (CUDA events, barrier + synchronize on both sides, max
over ranks, >= 3 warm-up steps, inputs rotate over 3 batches, working set >> L2).

    python tools/bench_train.py --workload finetune      [--batch 8]  [--steps K] [--warmup W]
    python tools/bench_train.py --workload ctc_pretrain  [--batch 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_train.py --workload finetune --gpus N

  finetune      DiCoW-v3 fine-tune step: large-v3-turbo encoder + FDDT + CTC head + 4-layer decoder, loss = 0.7 soft-label CE
                + 0.3 CTC, decoder frozen (the recipe: 720 M trainable parameters), bf16 operands / fp32 master weights,
                forward + backward + gradient all-reduce (overlapped, parallel.GradientExchange) + torch fused AdamW.
  ctc_pretrain  CTC encoder pre-training step: everything frozen but the CTC head (src/pretrain_encoder.py:42-51).

A step shards over utterances: B per GPU fixed ("weak" scaling), one exchange step (the gradient all-reduce)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SOT, LANG, TASK, EOS, TS_BEGIN, N_TS = 50258, 50259, 50360, 50257, 50365, 1501
# fwd 2273.8 (encoder) + 92.0 (CTC head) + ~84 (decoder at S=64 incl. cross K/V and proj_out) GFLOP / utt;
# backward = 2x forward for trained parts; decoder wgrads skipped when frozen
GFLOP_FINETUNE = 3 * 2273.8 + 3 * 92.0 + 2 * 84.0
GFLOP_CTC_PRETRAIN = 2273.8 + 3 * 92.0


class WhisperIds:
    """the token-id facts of the multilingual Whisper tokenizer the loss needs (no tokenizer files offline;
    export_sources/generation_config.json pins the ids)"""
    prefix_tokens = [SOT, LANG, TASK]
    pad_token_id = EOS

    def get_vocab(self):
        return {f"<|{0.02 * i:.2f}|>": TS_BEGIN + i for i in range(N_TS)}


def config(ctc_weight=0.3):
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    return DiCoWConfig(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, encoder_attention_heads=20,
                       decoder_layers=4, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120,
                       max_source_positions=1500, max_target_positions=448, use_fddt=True, use_pre_pos_fddt=True,
                       fddt_is_diagonal=True, non_target_fddt_value=0.5, fddt_init="suppressive", ctc_weight=ctc_weight,
                       additional_self_attention_layer=True, pre_ctc_sub_sample=True, activation_function="gelu",
                       pad_token_id=EOS, eos_token_id=EOS, bos_token_id=EOS, decoder_start_token_id=SOT)


def make_batch(B, S, seed, dev):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="finetune", choices=["finetune", "ctc_pretrain"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--labels", type=int, default=64)
    ap.add_argument("--no-overlap", action="store_true", help="exchange gradients after the backward instead of during it")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from ts_asr_whisper_b200 import ops, parallel, training
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    parallel.init_process_group("nccl", dev)
    fine = args.workload == "finetune"
    B = args.batch or (8 if fine else 16)
    torch.manual_seed(1234)
    with torch.device(dev):
        model = DiCoWForConditionalGeneration(config())
    model.tie_weights()
    model.set_tokenizer(WhisperIds())
    model.train()
    enc = model.get_encoder()
    head = ("model.encoder.additional_self_attention_layer", "model.encoder.subsample_conv", "model.encoder.lm_head")
    for n, p in model.named_parameters():
        if fine:  # decoder frozen, sinusoidal positions frozen
            p.requires_grad_(n.startswith("model.encoder.") and "embed_positions" not in n)
        else:
            p.requires_grad_(n.startswith(head))
    params = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in params)
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.0, fused=True)
    exchange = parallel.GradientExchange()
    if not args.no_overlap:
        training.gradient_exchange = exchange
    batches = [make_batch(B, args.labels, 100 + 10 * rank + i, dev) for i in range(3)]

    def step(i):
        feats, stno, labels, upp = batches[i % 3]
        opt.zero_grad(set_to_none=True)
        if fine:
            loss = model(feats, stno_mask=stno, labels=labels, upp_labels=upp).loss
        else:
            out = enc(feats, stno_mask=stno, return_logits=True)
            lab = labels[:, 3:].clone()  # src/utils/trainers.py:76-103: prompt tokens stripped, eos -> -100
            lab[lab == EOS] = -100
            loss = enc.get_loss(out.logits, lab)
        loss.backward()
        if args.no_overlap:
            parallel.allreduce_gradients(params)
        opt.step()
        return loss

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        loss = step(i)
    l0 = ops.launch_count
    torch.cuda.reset_peak_memory_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    barrier()
    ms = parallel.max_over_ranks([e0.elapsed_time(e1)], dev)[0]
    launches = ops.launch_count - l0
    if rank == 0:
        utts = world * B * args.steps
        gf = GFLOP_FINETUNE if fine else GFLOP_CTC_PRETRAIN
        value = utts / (ms * 1e-3)
        print(json.dumps({
            "metric": ("DiCoW-v3 fine-tune step" if fine else "CTC encoder pre-train step") + " utterances/sec, large-v3-turbo",
            "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[{2 if fine else 4}]", "batch_per_gpu": B, "label_len": args.labels,
                       "trainable_params": n_train, "optimizer": "torch AdamW(fused=True), fp32 master weights",
                       "exchange": ("after backward" if args.no_overlap else "overlapped per-layer buckets") +
                                   f", {exchange.n_collectives // max(1, args.steps + args.warmup)} collectives/step"},
            "tflops_per_gpu": value / world * gf / 1e3, "gflop_per_utt": gf, "gpu_launches": launches, "loss": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
