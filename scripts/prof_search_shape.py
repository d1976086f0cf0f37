"""One retrieval launch of a given shape for ncu counter captures:  prof_search_shape.py Q N [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops
Q, N = int(sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
d = 768
g = torch.Generator(device="cuda").manual_seed(2023)
c = torch.empty(N, d, device="cuda", dtype=torch.float16)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20)); c[s:e] = torch.randn(e - s, d, generator=g, device="cuda").half()
q = torch.randn(Q, d, generator=g, device="cuda").half()
qi, ci = ops.row_inv_norm(q), ops.row_inv_norm(c)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
ops.ip_topk_partial(q, qi, c, ci, k)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
print(f"Q {Q} N {N} ms {ms:.2f} TFLOPs {2.0 * Q * N * d / ms / 1e9:.1f}")
