"""HBM-roofline check of the row kernels at the training step's shapes (ViT-L/14 micro-batch of 128
items: 32896 x 1024 tokens; text tower: 9856 x 768).  CUDA events on the launch stream, inputs of one
launch >> 126 MB L2 for the vision shapes; a rotating set of buffers defeats L2 for the small ones.

    python scripts/bench_rowops.py [--json gpurun_out/rowops.json] [--once]

Algorithmic bytes per element:  LN fwd 4 (x fp32) + 2 (y bf16);  LN bwd 2 (dy) + 4 (x)
+ 4 (dx in) + 4 (dx out) + 2 (dx bf16);  colsum 2.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uniir_b200 import ops

dev = "cuda"
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]


ONCE = "--once" in sys.argv       # one launch per case, no warm-up: the ncu capture mode


def timeit(fn, nbuf, iters=20):
    if ONCE:
        fn(0)
        torch.cuda.synchronize()
        return float("nan")
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = []
    for name, T, D in (("vision 128 x 257 x 1024", 32896, 1024), ("text 128 x 77 x 768", 9856, 768)):
        nbuf = max(2, int(600e6 // (T * D * 4)) + 1)      # rotate through > 600 MB so nothing is served from L2
        xs = [torch.randn(T, D, device=dev) for _ in range(nbuf)]
        dys = [torch.randn(T, D, device=dev).bfloat16() for _ in range(nbuf)]
        dxs = [torch.zeros(T, D, device=dev) for _ in range(nbuf)]
        dxbs = [torch.zeros(T, D, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty(T, D, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
        mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
        dg, db, dc = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        ops.layernorm_fwd(xs[0], g, b, 1e-5, out=ys[0], stats=(mean, rstd))
        n = T * D
        cases = [
            ("layernorm_fwd", 6 * n, lambda i: ops.layernorm_fwd(xs[i], g, b, 1e-5, out=ys[i], stats=(mean, rstd))),
            ("layernorm_bwd", 16 * n, lambda i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, g, dxs[i], dg, db, dx_bf16=dxbs[i])),
            ("layernorm_bwd+colsum", 16 * n, lambda i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, g, dxs[i], dg, db,
                                                                          dx_bf16=dxbs[i], dcolsum=dc)),
            ("colsum [T,D]", 2 * n, lambda i: ops.colsum_accum(dys[i], dc)),
        ]
        for cname, nbytes, fn in cases:
            ms = timeit(fn, nbuf)
            gbs = nbytes / ms / 1e6
            out.append(dict(kernel=cname, shape=name, ms=ms, algorithmic_bytes=nbytes, gbs=gbs, frac_of_measured_hbm=gbs / PEAK))
            print(f"{cname:24s} {name:26s} {ms * 1e3:8.1f} us  {gbs:8.0f} GB/s  {gbs / PEAK:5.2f} of measured {PEAK:.0f} GB/s")
        del xs, dxs, dxbs, ys
        for mult, nm in ((3, "d_qkv [T,3D]"), (4, "d_pre [T,4D]")):
            wide = [torch.randn(T, mult * D, device=dev).bfloat16() for _ in range(2)]
            acc = torch.zeros(mult * D, device=dev)
            ms = timeit(lambda i: ops.colsum_accum(wide[i], acc), 2)
            gbs = 2 * T * mult * D / ms / 1e6
            out.append(dict(kernel=f"colsum {nm}", shape=name, ms=ms, algorithmic_bytes=2 * T * mult * D, gbs=gbs,
                            frac_of_measured_hbm=gbs / PEAK))
            print(f"{'colsum ' + nm:24s} {name:26s} {ms * 1e3:8.1f} us  {gbs:8.0f} GB/s  {gbs / PEAK:5.2f} of measured {PEAK:.0f} GB/s")
            del wide
        del dys
        torch.cuda.empty_cache()
    if "--json" in sys.argv:
        p = sys.argv[sys.argv.index("--json") + 1]
        os.makedirs(os.path.dirname(p) or ".", exist_ok=True)
        json.dump(dict(peak_hbm_gbs=PEAK, results=out), open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
