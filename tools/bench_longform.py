#!/usr/bin/env python
"""Diagnostic: long-form generate() over synthetic multi-window recordings with and without the speculative
next-window encoder pass.  large-v3-turbo + FDDT (optionally SE-DiCoW), random-init weights, EOS suppressed so that every
window decodes --tokens new tokens.  One JSON line per setting; interleaved A/B/A/B in one process.

    python tools/bench_longform.py [--batch 16] [--windows 4] [--tokens 64] [--se] [--timestamps] [--sms 96]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--windows", type=int, default=4)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--sms", type=int, default=0, help="0 = automatic")
    ap.add_argument("--se", action="store_true")
    ap.add_argument("--timestamps", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    if args.se:
        model = wl.se_dicow_model(dev)
    else:
        torch.manual_seed(4321)
        with torch.device(dev):
            model = DiCoWForConditionalGeneration(wl.turbo_config(pad_token_id=wl.EOS, eos_token_id=wl.EOS,
                                                                  decoder_start_token_id=wl.SOT)).eval()
        wl.perturb_(model.get_encoder(), dev)
    B, W = args.batch, args.windows
    g = torch.Generator().manual_seed(5)
    feats = (torch.randn(B, 128, 3000 * W, generator=g) * 0.4 - 0.3).clamp_(-1.0, 1.5).to(dev)
    stno = torch.softmax(3.0 * torch.randn(B, 4, 1500 * W, generator=g), dim=1).to(dev)
    enr = None
    if args.se:
        f, s = wl.make_inputs(B, 77, device=dev)
        enr = {"input_features": f, "stno_mask": s}
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = 50364, wl.EOS, wl.EOS
    gc.suppress_tokens = [wl.EOS, 220, 50256]  # EOS suppressed: fixed decode length
    gc.begin_suppress_tokens = None  # src/utils/general.py:26
    gc.return_timestamps, gc.max_new_tokens, gc.num_beams = bool(args.timestamps), args.tokens, 1
    prompt = torch.tensor([[wl.SOT, wl.LANG, wl.TASK] + ([] if args.timestamps else [50364])] * B)
    kw = dict(stno_mask=stno, forced_decoder_ids=prompt, return_segments=True)
    if enr is not None:
        kw["enrollments"] = enr
    model.speculation_sms = args.sms or None

    def run(spec):
        model.speculate_next_window = spec
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model.generate(feats, **kw)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out, dict(model.speculation_stats)

    run(False), run(True)  # warm-up: graphs, lazy loads
    res = {False: [], True: []}
    ref = None
    for _ in range(args.reps):
        for spec in (False, True):
            ms, out, stats = run(spec)
            res[spec].append(ms)
            if ref is None:
                ref = out["sequences"]
            assert torch.equal(ref, out["sequences"]), "speculation changed the result"
            last = stats
    for spec in (False, True):
        ms = sorted(res[spec])[len(res[spec]) // 2]
        print(json.dumps({"workload": f"long-form generate, {B} recordings x {W} windows, {args.tokens} tokens / window, "
                                      f"{'SE-DiCoW' if args.se else 'DiCoW'} turbo, timestamps {'on' if args.timestamps else 'off'}",
                          "speculate_next_window": spec, "speculation_sms": args.sms if spec else None,
                          "ms_per_call_median": round(ms, 2), "windows_per_s": round(B * W / ms * 1e3, 1),
                          "runs_ms": [round(x, 1) for x in res[spec]], "stats": last if spec else None}))


if __name__ == "__main__":
    main()
