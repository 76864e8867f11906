#!/usr/bin/env python
"""Diagnostic: per-kernel GPU time of one training step (torch.profiler / CUPTI, kernels not serialised).
    python tools/profile_train.py [--workload finetune|ctc_pretrain] [--batch B]
Not a benchmark: numbers taken under a profiler are never reported as bench values."""
import argparse
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads as bt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="finetune")
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dev = torch.device("cuda", 0)
    fine = args.workload == "finetune"
    B = args.batch or (8 if fine else 16)
    with torch.device(dev):
        model = DiCoWForConditionalGeneration(bt.train_config())
    model.tie_weights()
    model.set_tokenizer(bt.WhisperIds())
    model.train()
    enc = model.get_encoder()
    head = ("model.encoder.additional_self_attention_layer", "model.encoder.subsample_conv", "model.encoder.lm_head")
    for n, p in model.named_parameters():
        p.requires_grad_((n.startswith("model.encoder.") and "embed_positions" not in n) if fine else n.startswith(head))
    params = [p for p in model.parameters() if p.requires_grad]
    from ts_asr_whisper_b200.optim import AdamW
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.0, fused=True) if os.environ.get("DICOW_TORCH_ADAMW") == "1" \
        else AdamW(params, lr=1e-5, weight_decay=0.0)
    batch = bt.make_train_batch(B, 64, 5, dev)

    def step():
        feats, stno, labels, upp = batch
        opt.zero_grad(set_to_none=True)
        if fine:
            loss = model(feats, stno_mask=stno, labels=labels, upp_labels=upp).loss
        else:
            lab = labels[:, 3:].clone()
            lab[lab == bt.EOS] = -100
            loss = enc.get_loss(enc(feats, stno_mask=stno, return_logits=True).logits, lab)
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    total = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            name = name[:90]
            a = agg.setdefault(name, [0.0, 0])
            a[0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            a[1] += 1
            total += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    # idle gaps of the device inside the step: the largest ones and the kernel that ends each
    evs = sorted(((ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
                  if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
    if evs:
        span = (max(e[1] for e in evs) - evs[0][0]) / 1e3
        gaps, busy_until, prev = [], evs[0][1], evs[0][2]
        for st, en, nm in evs[1:]:
            if st > busy_until:
                gaps.append((st - busy_until, nm + "   <- after " + prev[:60]))
            if en >= busy_until:
                prev = nm
            busy_until = max(busy_until, en)
        print(f"device span of the step {span:.2f} ms; idle {sum(g[0] for g in gaps) / 1e3:.2f} ms in {len(gaps)} gaps; largest:")
        for us, nm in sorted(gaps, reverse=True)[:12]:
            print(f"   {us:8.1f} us before {nm[:170]}")
    print(f"total GPU kernel time {total / 1e3:.2f} ms over {sum(a[1] for a in agg.values())} launches")
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:80]:
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  x{n:<5d} {name}")


if __name__ == "__main__":
    main()
