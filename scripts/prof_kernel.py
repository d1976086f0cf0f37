"""Launch one kernel configuration a few times for an `ncu --set full` capture.
    python scripts/prof_kernel.py gemm_ags | gemm_fc1 | gemm_out | search"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops
which = sys.argv[1]
M = 16448
base = dict(M=M, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=False, act=0,
            residual=None, aux=False, ags=False, accumulate=False, split_k=1)
if which == "gemm_ags":
    print(ops.time_gemm_config({**base, "N": 4096, "K": 1024, "b_mn": True, "act": 1, "ags": True}, iters=3))
elif which == "gemm_fc1":
    print(ops.time_gemm_config({**base, "N": 4096, "K": 1024, "bias": True, "act": 1, "aux": True}, iters=3))
elif which == "gemm_fc2":
    print(ops.time_gemm_config({**base, "N": 1024, "K": 4096, "bias": True, "out_dtype": torch.float32, "residual": torch.float32}, iters=3))
elif which == "gemm_out":
    print(ops.time_gemm_config({**base, "N": 1024, "K": 1024, "bias": True, "out_dtype": torch.float32, "residual": torch.float32}, iters=3))
elif which == "search":
    Q, N, d, k = 32768, 1_400_000, 768, 50
    q = torch.randn(Q, d, device="cuda").half(); c = torch.randn(N, d, device="cuda").half()
    qi = ops.row_inv_norm(q); ci = ops.row_inv_norm(c)
    for _ in range(2):
        ops.ip_topk_partial(q, qi, c, ci, k)
    torch.cuda.synchronize()
if which == "attn":
    B, S, H = 64, 257, 16
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3):
        out, lse = ops.attn_fwd(qkv, B, S, H, False)
        ops.attn_bwd(qkv, out, torch.randn_like(out), lse, B, S, H, False)
    torch.cuda.synchronize()
print("done")
