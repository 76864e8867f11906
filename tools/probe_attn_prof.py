"""In-kernel clock64 profile of one CTA of the ping-pong attention kernel (debug aid, not part of the data path)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops, lib
dev = torch.device("cuda:0")
B, H, T = 32, 20, 1500
d = H * 64
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
prof = torch.zeros(64, 8, dtype=torch.int64, device=dev)
h = lib.handle(0)
for variant in [int(a) for a in sys.argv[1:]] or [0]:
    lib.load_library().dicow_debug_set_attention_profile(h, prof.data_ptr())
    for _ in range(2):
        ops.attention(qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d,
                      q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d,
                      o_batch_stride=T * d, variant=variant)
    torch.cuda.synchronize()
    lib.load_library().dicow_debug_set_attention_profile(h, None)
    p = prof.cpu()
    t0 = p[0, 0].item()
    nj = int(os.environ.get("NJ", "12"))
    print(f"variant {variant}: softmax thread (row 0) of each tile, clk")
    for t in range(2):
        for j in range(nj):
            r = p[t * 16 + j].tolist()
            print(f"  tile {t} j={j:2d} t={r[0]-t0:7d}  waitS {r[1]-r[0]:5d}  ld {r[2]-r[1]:5d}  max/resc {r[3]-r[2]:5d}  "
                  f"exp {r[4]-r[3]:5d}  stwait {r[5]-r[4]:4d}  arrive {r[6]-r[5]:4d} | total {r[6]-r[0]:5d}")
    print("  MMA thread: [sawP0 PV0-issued S0'-issued | sawP1 PV1-issued S1'-issued] relative to t0")
    for j in range(nj):
        r = [x - t0 for x in p[32 + j].tolist()[:6]]
        print(f"  j={j:2d} " + " ".join(f"{x:7d}" for x in r))
