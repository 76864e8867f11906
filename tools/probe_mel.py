"""Timing probe for dicow_logmel: B 30 s windows, M=128."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops
from ts_asr_whisper_b200.feature_extraction import DiCoWFeatureExtractor
dev = torch.device("cuda:0")
fe = DiCoWFeatureExtractor(feature_size=128, device=dev)
filt = fe._filters(dev)
for B in (1, 32, 128):
    audio = torch.randn(B, 480000, device=dev) * 0.1
    for _ in range(3):
        ops.logmel(audio, filt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.logmel(audio, filt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = B * (480000 * 4 + 3 * 128 * 3000 * 4) / 1e9  # audio in, features out (written, re-read, re-written)
    gf = B * 3000 * (2 * 2 * 201 * 200 + 2 * 402) / 1e9
    print(f"logmel B={B}: {ms:.3f} ms  {ms * 1e3 / B:.2f} us/window  {gf / ms:.1f} GFLOP/ms fp32  {gb / ms * 1e3:.0f} GB/s")
