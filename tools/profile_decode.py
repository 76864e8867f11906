#!/usr/bin/env python
"""Diagnostic: GPU timeline of the greedy decode step (torch.profiler / CUPTI): per-kernel durations inside the replayed
CUDA graph and the idle gaps between them.  Not a benchmark (numbers under a profiler are never reported as bench values).
    python tools/profile_decode.py [--unfused] [--no-graphs] [--batch 16]"""
import argparse
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import turbo_config  # noqa: E402
from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--ln-prologue", action="store_true")
ap.add_argument("--no-graphs", action="store_true")
ap.add_argument("--beams", type=int, default=1)
ap.add_argument("--ctc-weight", type=float, default=0.0)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = turbo_config()
cfg.encoder_layers = 1
cfg.pad_token_id = cfg.eos_token_id = 50257
with torch.device(dev):
    model = DiCoWForConditionalGeneration(cfg)
model.eval()
model.use_cuda_graphs = not args.no_graphs
model.fused_decode_step = False if args.unfused else ("ln_prologue" if args.ln_prologue else True)
B, T, d = args.batch, 1500, cfg.d_model
enc = (torch.randn(B, T, d, device=dev) * 0.5).bfloat16()
prompt = torch.tensor([[50258, 50259, 50360]] * B, device=dev)
rules = dict(eos=50257, pad=50257, no_timestamps=50364, ts_begin=50365, max_initial_timestamp_index=None,
             timestamp_rules=True, suppress_bitmap=model._suppress_bitmap([50257, 220, 50256], dev))
n = 3 + args.steps
ctc = None
if args.ctc_weight > 0:
    ctc = {"logits": torch.randn(B, 375, cfg.vocab_size + 1, device=dev) * 2.0, "weight": args.ctc_weight, "prefix_len": 3,
           "bos": 50258}


def decode():
    if args.beams > 1:
        return model.beam_decode_window(enc, prompt, n, rules, num_beams=args.beams, length_penalty=0.1, ctc=ctc)
    return model.greedy_decode_window(enc, prompt, n, rules, ctc=ctc)


for _ in range(2):
    decode()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    decode()
    torch.cuda.synchronize()
evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("dicow::", ""))
        evs.append((ev.time_range.start, ev.time_range.end, name))
evs.sort()
# the decode steps: everything from the first embed_kernel on
first = next(i for i, e in enumerate(evs) if "embed_kernel" in e[2])
steps = evs[first:]
span = steps[-1][1] - steps[0][0]
busy = sum(e[1] - e[0] for e in steps)
gaps = [max(0.0, steps[i + 1][0] - steps[i][1]) for i in range(len(steps) - 1)]
overlap = sum(max(0.0, steps[i][1] - steps[i + 1][0]) for i in range(len(steps) - 1))
nsteps = sum(1 for e in steps if "embed_kernel" in e[2])
print(f"{len(steps)} kernels over {nsteps} steps: span {span / nsteps:.1f} us/step, kernel time {busy / nsteps:.1f} us/step, "
      f"idle gaps {sum(gaps) / nsteps:.1f} us/step, overlap (PDL) {overlap / nsteps:.1f} us/step")
agg = collections.OrderedDict()
for s, e, name in steps:
    a = agg.setdefault(name[:60], [0.0, 0])
    a[0] += e - s
    a[1] += 1
for name, (us, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us / nsteps:8.1f} us/step  x{k / nsteps:5.1f}  avg {us / k:6.1f} us  {name}")
print("one step, in order (start offset us, duration us):")
i0 = [i for i, e in enumerate(steps) if "embed_kernel" in e[2]][min(10, nsteps - 1)]
t0 = steps[i0][0]
for s, e, name in steps[i0:i0 + 60]:
    if "embed_kernel" in name and s != t0:
        break
    print(f"  {s - t0:8.1f} {e - s:7.1f}  {name[:50]}")
