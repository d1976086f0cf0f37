"""tcgen05 attention backward (mode bit 2) against the fp32 reference and the mma.sync kernels, one shape per
subprocess (a device trap kills the CUDA context, the next shape still runs).  Prints per-tensor errors.

    python scripts/check_attn_bwd.py            # all shapes
    python scripts/check_attn_bwd.py one B S H causal [Skv]

Variants are selected by environment (read once by the library): UNIIR_ATTN_BWD_EW=16 (16 elementwise warps),
UNIIR_ATTN_BWD_FOLD=1 (tail column / row inside the main kernel: correct on hardware, 2.2x slower, opt-in only).
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(1, 64, 1, 0), (1, 128, 1, 0), (2, 100, 2, 0), (1, 192, 1, 0), (2, 197, 2, 0), (1, 256, 1, 0), (1, 257, 1, 0),
          (3, 257, 4, 0), (5, 77, 3, 1), (1, 130, 2, 1), (1, 128, 1, 1), (20, 257, 16, 0), (40, 77, 12, 1),
          (3, 100, 12, 0, 197)]


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def one(B, S, H, causal, Skv=None):
    import torch
    from uniir_b200 import ops
    dev = "cuda"
    g = torch.Generator().manual_seed(S + B)
    D = H * 64
    if Skv is None:
        qkv = (torch.randn(B * S, 3 * D, generator=g) * 1.5).bfloat16()
        dout = torch.randn(B * S, D, generator=g).bfloat16()
        out, lse = ops.attn_fwd(qkv.to(dev), B, S, H, bool(causal))
        qr = qkv.float().requires_grad_(True)
        q, k, v = qr.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        att = q @ k.transpose(-1, -2) * 0.125
        if causal:
            att = att + torch.full((S, S), float("-inf")).triu_(1)
        ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B * S, D)
        ref.backward(dout.float())
        gref = qr.grad
        res = {}
        for mode in (0, 2):
            prev = ops.attn_set_mode(mode)
            try:
                res[mode] = ops.attn_bwd(qkv.to(dev), out, dout.to(dev), lse, B, S, H, bool(causal)).clone()
                torch.cuda.synchronize()
            finally:
                ops.attn_set_mode(prev)
        names = ("dq", "dk", "dv")
        msg = []
        worst = 0.0
        for i, n in enumerate(names):
            a = res[2][:, i * D:(i + 1) * D]; b = gref[:, i * D:(i + 1) * D]; c = res[0][:, i * D:(i + 1) * D]
            e_ref, e_mma, e0 = rel(a, b), rel(a, c), rel(c, b)
            worst = max(worst, e_ref)
            msg.append(f"{n}: vs fp32 {e_ref:.2e} (mma.sync {e0:.2e}) vs mma {e_mma:.2e}")
            if e_ref > 1e-2:   # where?
                d = (a.float().cpu() - b).view(B, S, H, 64).abs().amax(dim=(0, 2, 3))
                bad = (d > 0.05 * b.abs().max()).nonzero().flatten().tolist()
                msg.append(f"   bad rows ({len(bad)}): {bad[:12]}{'...' if len(bad) > 12 else ''}")
        print(f"B={B} S={S} H={H} causal={causal}: " + "; ".join(msg), "OK" if worst < 1e-2 else "FAIL", flush=True)
    else:
        q = torch.randn(B * S, D, generator=g).bfloat16().to(dev)
        kv = torch.randn(B * Skv, 2 * D, generator=g).bfloat16().to(dev)
        dout = torch.randn(B * S, D, generator=g).bfloat16().to(dev)
        kvl = torch.tensor([Skv, max(1, Skv // 3), 1][:B] + [Skv] * max(0, B - 3), dtype=torch.int32, device=dev)
        lse = torch.empty(B, H, S, device=dev)
        out = ops.attn_fwd_ex(q, kv[:, :D], kv[:, D:], B, S, Skv, H, kv_len=kvl, lse=lse)
        res = {}
        for mode in (0, 2):
            prev = ops.attn_set_mode(mode)
            try:
                dq = torch.zeros(B * S, D, device=dev, dtype=torch.bfloat16)
                dkv = torch.zeros(B * Skv, 2 * D, device=dev, dtype=torch.bfloat16)
                ops.attn_bwd_ex(q, kv[:, :D], kv[:, D:], out, dout, lse, B, S, Skv, H, dq, dkv[:, :D], dkv[:, D:], kv_len=kvl)
                torch.cuda.synchronize()
                res[mode] = (dq, dkv)
            finally:
                ops.attn_set_mode(prev)
        e = [rel(res[2][0], res[0][0]), rel(res[2][1][:, :D], res[0][1][:, :D]), rel(res[2][1][:, D:], res[0][1][:, D:])]
        print(f"cross B={B} Sq={S} Skv={Skv} H={H} kv_len: dq {e[0]:.2e} dk {e[1]:.2e} dv {e[2]:.2e} vs mma.sync",
              "OK" if max(e) < 1e-2 else "FAIL", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "from":
        for idx in range(int(sys.argv[2]), len(SHAPES)):
            print(f"#{idx}", flush=True)
            one(*SHAPES[idx])
    elif len(sys.argv) > 1 and sys.argv[1] == "one":
        one(*[int(x) for x in sys.argv[2:]])
    else:
        # one process for as many shapes as survive; after a crash / hang continue with the next shape
        start = 0
        while start < len(SHAPES):
            try:
                r = subprocess.run([sys.executable, __file__, "from", str(start)], capture_output=True, text=True, timeout=240)
                out, rc = r.stdout, r.returncode
                err = r.stderr
            except subprocess.TimeoutExpired as e:
                out, rc, err = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), -999, "TIMEOUT"
            done = [int(l[1:]) for l in out.splitlines() if l.startswith("#")]
            print("\n".join(l for l in out.splitlines() if not l.startswith("#")), flush=True)
            if rc == 0:
                break
            failed = done[-1] if done else start
            print(f"shape {SHAPES[failed]}: process exit {rc}: " + " | ".join(err.strip().splitlines()[-4:]), flush=True)
            start = failed + 1
