"""In-kernel clock64 profile of one attention CTA (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops, lib
dev = torch.device("cuda:0")
B, H, T = 32, 20, 1500
d = H * 64
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
prof = torch.zeros(16, 8, dtype=torch.int64, device=dev)
h = lib.handle(0)
for variant in [int(a) for a in sys.argv[1:]] or [8, 10, 0]:
    lib.load_library().dicow_debug_set_attention_profile(h, prof.data_ptr())
    for _ in range(2):
        ops.attention(qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d,
                      q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d,
                      o_batch_stride=T * d, variant=variant)
    torch.cuda.synchronize()
    lib.load_library().dicow_debug_set_attention_profile(h, None)
    p = prof.cpu()
    t0 = p[0, 0].item()
    print(f"variant {variant}: per step [wait_S, ld+max, rescale, exp+P, arrive | mma: P seen->PV issued] (clk); step total")
    for j in range(12):
        r = p[j].tolist()
        nxt = p[j + 1, 0].item() if j < 11 else r[5]
        print(f"  j={j:2d} t={r[0]-t0:7d}  waitS {r[1]-r[0]:5d}  ld+max {r[2]-r[1]:5d}  resc {r[3]-r[2]:5d}  exp+P {r[4]-r[3]:5d}  "
              f"arr {r[5]-r[4]:4d} | mma sawP@{r[6]-t0:7d} (+{r[6]-r[5]:4d} after arrive) issued +{r[7]-r[6]:4d} | step {nxt-r[0]:5d}")
