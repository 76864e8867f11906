"""Greedy decode throughput (BASELINE configs[3] shape: large-v3-turbo decoder, B=16 windows, fixed step count with EOS
suppressed so runs are comparable).  Reports ms / step against the HBM floor of the step
(decoder weights 344 MB + B x 30.7 MB cross-KV per step)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import turbo_config  # noqa: E402
from ts_asr_whisper_b200 import ops  # noqa: E402
from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--no-graphs", action="store_true")
ap.add_argument("--unfused", action="store_true", help="one kernel per operation (the pre-fusion decode step)")
ap.add_argument("--ln-prologue", action="store_true", help="LayerNorm as the prologue of the linear kernel")
ap.add_argument("--beams", type=int, default=1, help="se_dicow workload: beam search with this many beams")
ap.add_argument("--ctc-weight", type=float, default=0.0, help="se_dicow workload: joint CTC / attention decoding weight")
ap.add_argument("--workload", default="decoder", choices=["decoder", "se_dicow"],
                help="decoder: token steps on synthetic encoder states; se_dicow: BASELINE configs[3] end to end "
                     "(SE-DiCoW encoder with enrollment streams + 8 SCB layers, then the greedy loop)")
args = ap.parse_args()


def se_dicow_e2e():
    """BASELINE configs[3] end to end: tools/workloads.se_dicow_greedy (the function bench.py's "secondary" runs)"""
    from bench import ClockSampler
    from tools import workloads
    dev = torch.device("cuda:0")
    sampler = ClockSampler(0)
    sampler.start()
    out = workloads.se_dicow_greedy(dev, 0, 1, batch=args.batch, new_tokens=args.steps, sampler=sampler, beams=args.beams,
                                    ctc_weight=args.ctc_weight, graphs=not args.no_graphs,
                                    fused=False if args.unfused else ("ln_prologue" if args.ln_prologue else True))
    sampler.stop()
    print(json.dumps(out))


if args.workload == "se_dicow":
    se_dicow_e2e()
    sys.exit(0)
dev = torch.device("cuda:0")
cfg = turbo_config()
cfg.encoder_layers = 1  # the encoder is not what is measured here; hidden states are synthetic
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
for _ in range(2):
    model.greedy_decode_window(enc, prompt, n, rules)
torch.cuda.synchronize()
l0 = ops.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ids = model.greedy_decode_window(enc, prompt, n, rules)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
wbytes = sum(p.numel() for p in model.model.decoder.layers.parameters()) * 2 + cfg.vocab_size * d * 2
kvbytes = B * cfg.decoder_layers * T * 2 * d * 2
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except OSError:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
step_ms = ms / (args.steps + 2)
floor_ms = (wbytes + kvbytes) / (hbm * 1e9) * 1e3
print(json.dumps({"metric": "greedy decode, turbo decoder", "batch": B, "steps": args.steps, "ms_total": ms,
                  "ms_per_step": step_ms, "tokens_per_s": B * args.steps / (ms * 1e-3), "cuda_graphs": model.use_cuda_graphs,
                  "fused_step": model.fused_decode_step, "pdl": os.environ.get("DICOW_PDL", "0") == "1",
                  "bytes_per_step": wbytes + kvbytes, "hbm_floor_ms": floor_ms, "frac_of_hbm_roofline": floor_ms / step_ms,
                  "achieved_gbs": (wbytes + kvbytes) / (step_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                  "note": "includes the per-window cross-K/V projection (4 GEMMs) and 2 prompt steps",
                  "generated_tail": ids[0, -4:].tolist()}))
