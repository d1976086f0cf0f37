"""Phase timeline (SM clock cycles) of CTA 0 of the tcgen05 attention forward: where a tile's time goes.
    python scripts/trace_attn.py [B S H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops, lib
B, S, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 257, 16)
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
ops.attn_set_mode(5)
for _ in range(2):
    ops.attn_fwd(qkv, B, S, H, False)
buf = torch.zeros(3 * 256, dtype=torch.int64, device="cuda")
lib.load().uniir_attn_set_trace(buf.data_ptr())
ops.attn_fwd(qkv, B, S, H, False)
torch.cuda.synchronize()
lib.load().uniir_attn_set_trace(None)
t = buf.cpu().view(3, 256)
t0 = int(t[t > 0].min())
names = ["wait s_full", "s_full", "pass1 done", "xmax bar", "chunk0", "chunk1", "chunk2", "chunk3", "chunk4", "xsum bar", "o_full", "epilogue done"]
print("softmax warp 2 (cycles since first event):")
for tile in range(6):
    row = t[0, tile * 12:(tile + 1) * 12] - t0
    print(f" tile {tile}: " + "  ".join(f"{n}={int(v)}" for n, v in zip(names, row) if v > -t0))
mn = ["wait q_full", "q_full", "s_free", "QK issued", "o_free", "p_full0", "p_full1", "p_full2", "p_full3", "p_full4"]
print("MMA thread:")
for tile in range(6):
    row = t[1, tile * 12:tile * 12 + 10] - t0
    print(f" tile {tile}: " + "  ".join(f"{n}={int(v)}" for n, v in zip(mn, row) if v > -t0))
print("tail warp (per head: wait kv, kv_full, done):")
for hd in range(4):
    row = t[2, hd * 3:hd * 3 + 3] - t0
    print(f" head {hd}: " + "  ".join(str(int(v)) for v in row))
