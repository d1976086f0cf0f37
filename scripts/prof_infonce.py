"""In-batch InfoNCE kernels at the BASELINE configs[1] shapes, for ncu / CUDA-event timing:
per rank at 8 GPUs (bs 512 x 4096 gathered candidates) and the whole batch on one GPU (4096 x 4096), d = 768.
Algorithmic bytes: bs*d*4 (q) + n_cols*d*4 (all_p) + outputs; FLOPs 2*bs*n_cols*d forward.

    python scripts/prof_infonce.py [--json out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uniir_b200 import ops

dev = "cuda"
E = 768
res = []
for bs, n_cols in ((512, 4096), (4096, 4096)):
    g = torch.Generator(device=dev).manual_seed(bs)
    emb = torch.randn(2 * bs, E, device=dev, generator=g)
    q_idx = (torch.arange(bs, device=dev, dtype=torch.int32) * 2).contiguous()
    p_idx = q_idx + 1
    nce = ops.InfoNCE(bs, n_cols, E, dev)
    nce.all_p.copy_(torch.nn.functional.normalize(torch.randn(n_cols, E, device=dev, generator=g), dim=-1))
    ls = torch.full((), 4.6052, device=dev)
    dls = torch.zeros((), device=dev)
    d_emb = torch.zeros(2 * bs, E, device=dev)

    def fwd():
        nce.normalize(emb, q_idx, p_idx, 0)
        return nce.forward(ls, 0)

    def bwd():
        d_qn, d_all_p = nce.backward(ls, 0, dls)
        nce.normalize_bwd(q_idx, p_idx, d_all_p[:bs], d_emb, 0)

    for _ in range(3):
        fwd(); bwd()
    torch.cuda.synchronize()
    row = dict(bs=bs, n_cols=n_cols, d=E)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        row[f"{name}_us"] = e0.elapsed_time(e1) / 20 * 1e3
    row["algorithmic_bytes_fwd"] = (bs + n_cols) * E * 4 + bs * 8
    row["fwd_gbps"] = row["algorithmic_bytes_fwd"] / (row["fwd_us"] * 1e-6) / 1e9
    row["fwd_tflops"] = 2.0 * bs * n_cols * E / (row["fwd_us"] * 1e-6) / 1e12
    res.append(row)
    print(row, flush=True)
if "--json" in sys.argv:
    json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
