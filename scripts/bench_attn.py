"""Attention forward and backward: mma.sync head-resident kernels (mode 0) vs tcgen05 kernels (mode bits 1 =
forward, 2 = backward) at the tower shapes.  Algorithmic FLOPs = 4 * B * H * Sq * Skv * 64 forward, 2.5 x that
backward (halved for the causal text tower).  CUDA events, 20 launches.

    python scripts/bench_attn.py [--json gpurun_out/attn.json] [--once]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uniir_b200 import ops

dev = "cuda"
ONCE = "--once" in sys.argv


def timeit(fn, iters=20):
    if ONCE:
        fn(); torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = []
    shapes = [("ViT-L/14 vision, 128 items", 128, 257, 16, False), ("CLIP text, 128 items", 128, 77, 12, True),
              ("BLIP ViT-L/16, 64 items", 64, 197, 16, False), ("BERT self-attn, 64 items", 64, 100, 12, False)]
    for name, B, S, H, causal in shapes:
        qkv = (torch.randn(B * S, 3 * H * 64, device=dev) * 1.0).bfloat16()
        out = torch.empty(B * S, H * 64, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, S, device=dev)
        flops = 4.0 * B * H * S * S * 64 * (0.5 if causal else 1.0)
        row = dict(shape=name, B=B, S=S, H=H, causal=causal, algorithmic_flops=flops)
        outs = {}
        for mode, m in ((0, 0), (1, 5)):     # 5 = tcgen05 forward forced (also for the causal text shape)
            prev = ops.attn_set_mode(m)
            try:
                ms = timeit(lambda: ops.attn_fwd(qkv, B, S, H, causal, out=out, lse=lse))
                outs[mode] = out.clone()
            finally:
                ops.attn_set_mode(prev)
            row[f"ms_mode{mode}"] = ms
            row[f"tflops_mode{mode}"] = flops / ms / 1e9
        row["max_abs_diff"] = (outs[0].float() - outs[1].float()).abs().max().item()
        # backward (delta + dQ + dK/dV; tcgen05: prep + one fused kernel)
        dout = torch.randn(B * S, H * 64, device=dev).bfloat16()
        dqkv = torch.empty_like(qkv)
        ws = torch.empty(ops.attn_bwd_ws_floats(B, S, S, H), device=dev)
        grads = {}
        for mode in (0, 2):
            prev = ops.attn_set_mode(mode)
            try:
                ms = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, B, S, H, causal, dqkv=dqkv, delta=ws))
                grads[mode] = dqkv.clone()
            finally:
                ops.attn_set_mode(prev)
            row[f"bwd_ms_mode{mode}"] = ms
            row[f"bwd_tflops_mode{mode}"] = 2.5 * flops / ms / 1e9
        row["bwd_rel_diff"] = ((grads[0].float() - grads[2].float()).norm() / grads[0].float().norm()).item()
        print(f"{name:28s} bwd: mma.sync {row['bwd_ms_mode0'] * 1e3:8.1f} us ({row['bwd_tflops_mode0']:6.1f} TF/s)   "
              f"tcgen05 {row['bwd_ms_mode2'] * 1e3:8.1f} us ({row['bwd_tflops_mode2']:6.1f} TF/s)   rel diff {row['bwd_rel_diff']:.3e}")
        res.append(row)
        print(f"{name:28s} mma.sync {row['ms_mode0'] * 1e3:8.1f} us ({row['tflops_mode0']:6.1f} TF/s)   "
              f"tcgen05 {row['ms_mode1'] * 1e3:8.1f} us ({row['tflops_mode1']:6.1f} TF/s)   max|diff| {row['max_abs_diff']:.3e}")
    if "--json" in sys.argv:
        p = sys.argv[sys.argv.index("--json") + 1]
        os.makedirs(os.path.dirname(p) or ".", exist_ok=True)
        json.dump(res, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
