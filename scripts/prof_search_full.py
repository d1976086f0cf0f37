"""One full-size retrieval launch (190k x 5.6M x 768, top-50) for an ncu DRAM-traffic capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops
Q, N, d, k = 190_000, 5_600_000, 768, 50
g = torch.Generator(device="cuda").manual_seed(2023)
c = torch.empty(N, d, device="cuda", dtype=torch.float16)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20)); c[s:e] = torch.randn(e - s, d, generator=g, device="cuda").half()
q = torch.randn(Q, d, generator=g, device="cuda").half()
qi, ci = ops.row_inv_norm(q), ops.row_inv_norm(c)
ops.ip_topk_partial(q, qi, c, ci, k)
torch.cuda.synchronize()
print("done")
