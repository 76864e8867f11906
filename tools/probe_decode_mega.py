"""Per-phase timeline of the persistent decode-layers kernel (csrc/decode_mega.cu): %globaltimer stamps of CTA 0 at every
phase boundary (work done / barrier passed), averaged over a few steps.  Debug aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.workloads import turbo_config, decode_rules  # noqa: E402
from ts_asr_whisper_b200 import lib  # noqa: E402
from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration  # noqa: E402

dev = torch.device("cuda:0")
cfg = turbo_config()
cfg.encoder_layers = 1
cfg.pad_token_id = cfg.eos_token_id = 50257
with torch.device(dev):
    model = DiCoWForConditionalGeneration(cfg).eval()
model.use_cuda_graphs = True  # steady state: the stamps are those of the LAST step of a long graph-replayed decode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
enc = (torch.randn(B, 1500, cfg.d_model, device=dev) * 0.5).bfloat16()
prompt = torch.tensor([[50258, 50259, 50360]] * B, device=dev)
rules = decode_rules(model, dev)
prof = torch.zeros(256, dtype=torch.int64, device=dev)
h = lib.handle(0)
lib.load_library().dicow_debug_set_attention_profile(h, prof.data_ptr())  # before capture: the pointer is baked into the graph
model.greedy_decode_window(enc, prompt, 3 + 8, rules)
names = ["embed"]
for li in range(cfg.decoder_layers):
    names += [f"L{li} ln1+qkv", f"L{li} self-attn", f"L{li} out_proj", f"L{li} ln2+q", f"L{li} cross-attn", f"L{li} out_proj2",
              f"L{li} ln3+fc1", f"L{li} fc2"]
acc = torch.zeros(len(names), 2, dtype=torch.float64)
n = 0
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.greedy_decode_window(enc, prompt, 3 + 96, rules)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / 98
    p = prof.cpu().double()
    for i in range(len(names)):
        acc[i, 0] += p[1 + 2 * i] - p[2 * i]      # work
        acc[i, 1] += p[2 + 2 * i] - p[1 + 2 * i]  # barrier
    n += 1
lib.load_library().dicow_debug_set_attention_profile(h, None)
acc /= n * 1e3
tot = acc.sum().item()
print(f"whole decode step (graph replay, incl. cross-K/V projection amortised over 98 steps): {ms_step * 1e3:.1f} us")
print(f"decode-layers kernel, B={B}: {tot:.1f} us per step on CTA 0 (work {acc[:, 0].sum():.1f} us, barriers {acc[:, 1].sum():.1f} us)")
for nm, (w, b) in zip(names[:9], acc[:9].tolist()):
    print(f"  {nm:16s} work {w:6.2f} us   barrier wait {b:6.2f} us")
kinds = {}
for nm, (w, b) in zip(names[1:], acc[1:].tolist()):
    k = nm.split(" ", 1)[1]
    kinds.setdefault(k, [0.0, 0.0])
    kinds[k][0] += w
    kinds[k][1] += b
print("  per phase kind, summed over the layers:")
for k, (w, b) in kinds.items():
    print(f"    {k:12s} work {w:6.1f} us   barrier wait {b:6.1f} us")
