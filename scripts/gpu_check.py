"""Diagnostic GPU checks (one case per process so a trap in one does not poison the rest).
    python scripts/gpu_check.py <case>
Checker here is torch fp32 on the GPU; the pytest parity tests use the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uniir_b200 import ops, lib as L

dev = "cuda"
torch.manual_seed(0)


def rel_err(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item()


def gemm_case(M, N, K, a_mn, b_mn, dtype=torch.bfloat16, **kw):
    A = torch.randn(K, M, device=dev).to(dtype) if a_mn else torch.randn(M, K, device=dev).to(dtype)
    B = torch.randn(K, N, device=dev).to(dtype) if b_mn else torch.randn(N, K, device=dev).to(dtype)
    Af = (A.t() if a_mn else A).float()
    Bf = (B.t() if b_mn else B).float()
    ref = Af @ Bf.t()
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, **kw)
    torch.cuda.synchronize()
    r, m = rel_err(out, ref)
    ok = r < 1e-3
    print(f"gemm M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} {dtype} rel={r:.3e} max={m:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        # locate the error pattern
        d = (out - ref).abs()
        bad = (d > 1e-2 * ref.abs().max()).nonzero()
        print("  first bad:", bad[:8].tolist(), " n_bad", len(bad), "of", d.numel())
        print("  out[0,:8]", out[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
    return ok


def case_gemm_kk():
    ok = True
    for (M, N, K) in [(128, 128, 64), (128, 256, 64), (256, 256, 128), (128, 128, 256), (1000, 768, 1024), (4096, 3072, 1024), (333, 4096, 768)]:
        ok &= gemm_case(M, N, K, False, False)
    ok &= gemm_case(512, 512, 512, False, False, dtype=torch.float16)
    return ok


def case_gemm_kmn():
    ok = True
    for (M, N, K) in [(128, 128, 64), (256, 256, 128), (1000, 768, 1024)]:
        ok &= gemm_case(M, N, K, False, True)
    return ok


def case_gemm_mnk():
    ok = True
    for (M, N, K) in [(128, 128, 64), (256, 256, 128), (1000, 768, 1024)]:
        ok &= gemm_case(M, N, K, True, False)
    return ok


def case_gemm_mnmn():
    ok = True
    for (M, N, K) in [(128, 128, 64), (256, 256, 128), (1024, 768, 1000), (4096, 1024, 16448)]:
        ok &= gemm_case(M, N, K, True, True)
    ok &= gemm_case(1024, 1024, 16448, True, True, split_k=8)
    return ok


def case_gemm_epi():
    ok = True
    M, N, K = 777, 1024, 512
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16() * 0.05
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    pre = A.float() @ B.float().t() + bias
    # bias + quickgelu + aux
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = ops.gemm(A, B, bias=bias, act=L.ACT_QUICKGELU, aux_out=aux)
    ref = pre * torch.sigmoid(1.702 * pre)
    r1, _ = rel_err(out, ref); r2, _ = rel_err(aux, pre)
    print(f"epi bias+quickgelu rel={r1:.3e} aux rel={r2:.3e}"); ok &= r1 < 1e-2 and r2 < 1e-2
    out = ops.gemm(A, B, bias=bias, act=L.ACT_GELU)
    r1, _ = rel_err(out, torch.nn.functional.gelu(pre)); print(f"epi gelu rel={r1:.3e}"); ok &= r1 < 1e-2
    # bias + residual fp32 -> fp32 (in place on residual)
    x = res.clone()
    ops.gemm(A, B, bias=bias, residual=x, out=x)
    r1, _ = rel_err(x, pre + res); print(f"epi bias+residual(inplace fp32) rel={r1:.3e}"); ok &= r1 < 1e-3
    # act grad
    src = torch.randn(M, N, device=dev).bfloat16()
    out = ops.gemm(A, B, act=L.ACT_QUICKGELU, act_grad_src=src, out_dtype=torch.float32)
    s = src.float(); sg = torch.sigmoid(1.702 * s)
    ref = (A.float() @ B.float().t()) * (sg * (1 + 1.702 * s * (1 - sg)))
    r1, _ = rel_err(out, ref); print(f"epi quickgelu-grad rel={r1:.3e}"); ok &= r1 < 1e-3
    # accumulate
    acc = torch.randn(M, N, device=dev); acc0 = acc.clone()
    ops.gemm(A, B, out=acc, accumulate=True, alpha=0.5)
    r1, _ = rel_err(acc, acc0 + 0.5 * (A.float() @ B.float().t())); print(f"epi accumulate rel={r1:.3e}"); ok &= r1 < 1e-3
    return ok


def search_ref(q, c, k):
    qn = torch.nn.functional.normalize(q.float(), dim=-1)
    cn = torch.nn.functional.normalize(c.float(), dim=-1)
    best_s = None
    chunk = 262144
    for s in range(0, c.shape[0], chunk):
        sc = qn @ cn[s:s + chunk].t()
        ts, ti = sc.topk(min(k, sc.shape[1]), dim=1)
        ti = ti + s
        if best_s is None:
            best_s, best_i = ts, ti
        else:
            cs = torch.cat([best_s, ts], 1); ci = torch.cat([best_i, ti], 1)
            o = cs.argsort(dim=1, descending=True, stable=True)[:, :k]
            best_s = cs.gather(1, o); best_i = ci.gather(1, o)
    return best_s, best_i


def search_case(Q, N, d, k):
    q = torch.randn(Q, d, device=dev).half(); c = torch.randn(N, d, device=dev).half()
    qi = ops.row_inv_norm(q); ci = ops.row_inv_norm(c)
    r, _ = rel_err(qi, 1 / q.float().norm(dim=1))
    t0 = time.time()
    ps, pi = ops.ip_topk_partial(q, qi, c, ci, k)
    s, i = ops.topk_merge(ps, pi, k)
    torch.cuda.synchronize(); t1 = time.time()
    rs, ri = search_ref(q, c, k)
    kk = min(k, N)
    match = (i[:, :kk] == ri[:, :kk]).float().mean().item()
    serr = (s[:, :kk] - rs[:, :kk]).abs().max().item()
    # set-level agreement
    inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i[:64, :kk], ri[:64, :kk])) / (64 * kk) if Q >= 64 else -1
    ok = match > 0.999 and serr < 1e-5
    print(f"search Q={Q} N={N} d={d} k={k} splits={ps.shape[1]} invnorm_rel={r:.2e} idx_match={match:.5f} set_match64={inter:.4f} score_maxerr={serr:.2e} t={t1-t0:.3f}s {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        print("  got ", i[0, :8].tolist(), s[0, :8].tolist()); print("  ref ", ri[0, :8].tolist(), rs[0, :8].tolist())
    return ok


def case_search_small():
    ok = search_case(100, 1000, 64, 10)
    ok &= search_case(300, 5000, 128, 10)
    ok &= search_case(130, 40, 128, 50)
    return ok


def case_search_mid():
    ok = search_case(1000, 100000, 768, 50)
    ok &= search_case(4096, 300001, 768, 50)
    return ok


def case_search_perf():
    Q, N, d, k = 16384, 1000000, 768, 50
    q = torch.randn(Q, d, device=dev).half(); c = torch.randn(N, d, device=dev).half()
    qi = ops.row_inv_norm(q); ci = ops.row_inv_norm(c)
    splits, wsb = ops.ip_topk_plan(Q, N, k)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    ps = torch.empty(Q, splits, k, device=dev); pi = torch.empty(Q, splits, k, device=dev, dtype=torch.int32)
    for _ in range(2):
        ops.ip_topk_partial(q, qi, c, ci, k, workspace=ws, out=(ps, pi))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.ip_topk_partial(q, qi, c, ci, k, workspace=ws, out=(ps, pi))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = 2.0 * Q * N * d
    print(f"search perf Q={Q} N={N}: {ms:.2f} ms  {fl/ms/1e9:.1f} TFLOP/s  splits={splits}")
    return True


def case_gemm_perf():
    for (M, N, K) in [(16448, 3072, 1024), (16448, 1024, 4096), (16448, 4096, 1024), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3): ops.gemm(A, B, out=out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.gemm(A, B, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10): torch.matmul(A, B.t(), out=out)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        print(f"gemm perf M={M} N={N} K={K}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.1f} TF/s | cublas {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.1f} TF/s")
    return True


if __name__ == "__main__":
    name = sys.argv[1]
    ok = globals()["case_" + name]()
    print(f"CASE {name}: {'PASS' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)
