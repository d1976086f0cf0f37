"""A/B timing of retrieval-kernel variants in ONE process on ONE box: the configurations (environment overrides read
at every launch) are run interleaved, `reps` times each, so box-to-box and thermal drift cancel.
    python scripts/bench_search_modes.py Q N reps "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512" "UNIIR_SEARCH_MODE=1" ..."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops
Q, N, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfgs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[4:]]
d, k = 768, 50
g = torch.Generator(device="cuda").manual_seed(2023)
c = torch.empty(N, d, device="cuda", dtype=torch.float16)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20)); c[s:e] = torch.randn(e - s, d, generator=g, device="cuda").half()
q = torch.randn(Q, d, generator=g, device="cuda").half()
qi, ci = ops.row_inv_norm(q), ops.row_inv_norm(c)
KEYS = ("UNIIR_SEARCH_MODE", "UNIIR_SEARCH_TPC", "UNIIR_SEARCH_ROT", "UNIIR_SEARCH_2CTA", "UNIIR_SEARCH_HINT", "UNIIR_SEARCH_EPIPIPE", "UNIIR_SEARCH_SELECT")
times = [[] for _ in cfgs]
ref = None
for r in range(reps + 1):
    for i, cfg in enumerate(cfgs):
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update(cfg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ps, pi = ops.ip_topk_partial(q, qi, c, ci, k); e1.record(); torch.cuda.synchronize()
        if r > 0:
            times[i].append(e0.elapsed_time(e1))
        if r == 0:      # all variants must return the same lists
            s_, i_ = ops.topk_merge(ps, pi, k)
            if ref is None:
                ref = (s_, i_)
            else:
                same = (i_ == ref[1]).float().mean().item()
                assert same > 0.9999 and (s_ - ref[0]).abs().max().item() < 1e-5, (cfg, same)
for cfg, t in zip(cfgs, times):
    t = sorted(t)
    print(json.dumps({"cfg": cfg, "ms_min": round(t[0], 1), "ms_med": round(t[len(t) // 2], 1),
                      "tflops_med": round(2.0 * Q * N * d / t[len(t) // 2] / 1e9, 1), "tflops_best": round(2.0 * Q * N * d / t[0] / 1e9, 1)}))
