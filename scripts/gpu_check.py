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
    o0 = torch.zeros(M, N, device=dev) if kw.get("split_k", 1) > 1 else None
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=o0, out_dtype=torch.float32, **kw)
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




def case_rowops():
    ok = True
    F = torch.nn.functional
    for D, rows in [(1024, 1000), (768, 333), (128, 50)]:
        x = torch.randn(rows, D, device=dev) * 2 + 0.5
        g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, save_stats=True)
        ref = F.layer_norm(x, (D,), g, b, 1e-5)
        r, _ = rel_err(y, ref); print(f"ln fwd D={D} rel={r:.2e}"); ok &= r < 5e-3
        y32 = ops.layernorm_fwd(x, g, b, 1e-5, out_dtype=torch.float32)
        r, _ = rel_err(y32, ref); print(f"ln fwd fp32 rel={r:.2e}"); ok &= r < 1e-5
        # backward
        dy = torch.randn(rows, D, device=dev).bfloat16()
        xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy.float())
        dx = torch.ones(rows, D, device=dev); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
        ops.layernorm_bwd(dy, x, mean, rstd, g, dx, dg, db)
        r1, _ = rel_err(dx - 1, xr.grad); r2, _ = rel_err(dg, gr.grad); r3, _ = rel_err(db, br.grad)
        print(f"ln bwd D={D} dx={r1:.2e} dgamma={r2:.2e} dbeta={r3:.2e}"); ok &= max(r1, r2, r3) < 1e-4
    # gathered rows
    x = torch.randn(10, 7, 128, device=dev); g = torch.randn(128, device=dev); b = torch.randn(128, device=dev)
    idx = (torch.arange(10, device=dev) * 7 + torch.randint(0, 7, (10,), device=dev)).int()
    y = ops.layernorm_fwd(x, g, b, 1e-5, row_index=idx, out_dtype=torch.float32)
    r, _ = rel_err(y, F.layer_norm(x.view(70, 128)[idx.long()], (128,), g, b, 1e-5)); print(f"ln gather rel={r:.2e}"); ok &= r < 1e-5
    # im2col + embed
    img = torch.randn(3, 3, 56, 56, device=dev)
    col = ops.patch_im2col(img, 14, 592)
    ref = F.unfold(img, 14, stride=14).transpose(1, 2).reshape(3 * 16, 588)
    r, _ = rel_err(col[:, :588], ref); z = col[:, 588:].abs().max().item(); print(f"im2col rel={r:.2e} pad={z}"); ok &= r < 5e-3 and z == 0
    patch = torch.randn(3 * 16, 128, device=dev).bfloat16(); cls = torch.randn(128, device=dev); pos = torch.randn(17, 128, device=dev)
    x0 = ops.vision_embed_fwd(patch, cls, pos, 3, 17, 128)
    ref = torch.cat([cls.expand(3, 1, 128), patch.float().view(3, 16, 128)], 1) + pos
    r, _ = rel_err(x0, ref); print(f"vision_embed rel={r:.2e}"); ok &= r < 1e-6
    dx0 = torch.randn(3, 17, 128, device=dev); dpatch = torch.empty(48, 128, device=dev, dtype=torch.bfloat16)
    dcls = torch.zeros(128, device=dev); dpos = torch.zeros(17, 128, device=dev)
    ops.vision_embed_bwd(dx0, 3, 17, 128, dpatch, dcls, dpos)
    r1, _ = rel_err(dpatch, dx0[:, 1:].reshape(48, 128)); r2, _ = rel_err(dcls, dx0[:, 0].sum(0)); r3, _ = rel_err(dpos, dx0.sum(0))
    print(f"vision_embed bwd {r1:.2e} {r2:.2e} {r3:.2e}"); ok &= r1 < 5e-3 and r2 < 1e-6 and r3 < 1e-6
    ids = torch.randint(1, 500, (5, 77), device=dev).int(); ids[:, 0] = 510; ids[torch.arange(5), torch.tensor([3, 10, 76, 5, 1])] = 511
    tok = torch.randn(512, 64, device=dev); tpos = torch.randn(77, 64, device=dev)
    x, eot = ops.text_embed_fwd(ids, tok, tpos)
    r, _ = rel_err(x, tok[ids.long()] + tpos); e_ok = (eot.long() == torch.arange(5, device=dev) * 77 + ids.argmax(-1)).all().item()
    print(f"text_embed rel={r:.2e} eot_ok={e_ok}"); ok &= r < 1e-6 and e_ok
    dx = torch.randn(5, 77, 64, device=dev); dtok = torch.zeros(512, 64, device=dev); dtpos = torch.zeros(77, 64, device=dev)
    ops.text_embed_bwd(dx, ids, dtok, dtpos)
    ref = torch.zeros(512, 64, device=dev).index_add_(0, ids.long().flatten(), dx.view(-1, 64))
    r1, _ = rel_err(dtok, ref); r2, _ = rel_err(dtpos, dx.sum(0)); print(f"text_embed bwd {r1:.2e} {r2:.2e}"); ok &= r1 < 1e-5 and r2 < 1e-6
    # fuse / colsum / pad / adamw
    t = torch.randn(9, 64, device=dev); v = torch.randn(9, 64, device=dev)
    mt = torch.randint(0, 2, (9,), device=dev); mv = torch.randint(0, 2, (9,), device=dev)
    f = ops.score_fuse_fwd(t, v, mt, mv); r, _ = rel_err(f, t * mt[:, None] + v * mv[:, None]); print(f"fuse rel={r:.2e}"); ok &= r < 1e-6
    dy = torch.randn(3000, 200, device=dev).bfloat16(); cs = torch.ones(200, device=dev)
    ops.colsum_accum(dy, cs); r, _ = rel_err(cs - 1, dy.float().sum(0)); print(f"colsum rel={r:.2e}"); ok &= r < 1e-4
    w = torch.randn(20, 588, device=dev); wp = ops.pad_cast_bf16(w, 640); r, _ = rel_err(wp[:, :588], w); ok &= r < 5e-3 and wp[:, 588:].abs().max().item() == 0
    acc = torch.ones(20, 588, device=dev); ops.unpad_accum(torch.randn(20, 640, device=dev), acc)
    n = 100003
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); vv = torch.zeros(n, device=dev)
    pr = p.clone().requires_grad_(True); opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    p16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        gg = torch.randn(n, device=dev); pr.grad = gg.clone(); opt.step()
        g.copy_(gg); ops.adamw_step(p, g, m, vv, p16, 1e-3, (0.9, 0.98), 1e-6, 0.2, step)
    r, mx = rel_err(p, pr.detach()); r2, _ = rel_err(p16, p); print(f"adamw rel={r:.2e} max={mx:.2e} g_zero={g.abs().max().item()} p16={r2:.2e}"); ok &= r < 1e-5 and g.abs().max().item() == 0
    return ok


def attn_ref(qkv, B, S, H, causal):
    D = H * 64
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = q @ k.transpose(-1, -2) * 0.125
    if causal:
        att = att + torch.full((S, S), float("-inf"), device=qkv.device).triu_(1)
    p = att.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B * S, D), torch.logsumexp(att, -1)


def case_attn():
    ok = True
    for (B, S, H, causal) in [(2, 64, 1, False), (3, 257, 4, False), (5, 77, 3, True), (2, 197, 2, False), (2, 50, 2, False), (1, 130, 2, True)]:
        qkv = (torch.randn(B * S, 3 * H * 64, device=dev) * 1.5).bfloat16().requires_grad_(False)
        out, lse = ops.attn_fwd(qkv, B, S, H, causal)
        qr = qkv.float().requires_grad_(True)
        ref, rlse = attn_ref(qr, B, S, H, causal)
        r1, _ = rel_err(out, ref); r2, m2 = rel_err(lse, rlse)
        dout = torch.randn(B * S, H * 64, device=dev).bfloat16()
        ref.backward(dout.float())
        dqkv = ops.attn_bwd(qkv, out, dout, lse, B, S, H, causal)
        D = H * 64
        rq, _ = rel_err(dqkv[:, :D], qr.grad[:, :D]); rk, _ = rel_err(dqkv[:, D:2 * D], qr.grad[:, D:2 * D]); rv, _ = rel_err(dqkv[:, 2 * D:], qr.grad[:, 2 * D:])
        good = r1 < 1e-2 and m2 < 1e-2 and max(rq, rk, rv) < 2e-2
        print(f"attn B={B} S={S} H={H} causal={causal}: out={r1:.2e} lse_max={m2:.2e} dq={rq:.2e} dk={rk:.2e} dv={rv:.2e} {'OK' if good else 'FAIL'}", flush=True)
        ok &= good
    # timing at model shape
    B, S, H = 64, 257, 16
    qkv = torch.randn(B * S, 3 * H * 64, device=dev).bfloat16()
    out, lse = ops.attn_fwd(qkv, B, S, H, False); dout = torch.randn_like(out); dqkv = torch.empty_like(qkv)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.attn_fwd(qkv, B, S, H, False, out=out, lse=lse)
    e1.record()
    for _ in range(10): ops.attn_bwd(qkv, out, dout, lse, B, S, H, False, dqkv=dqkv)
    e2.record(); torch.cuda.synchronize()
    fl = 4.0 * B * H * S * S * 64
    print(f"attn perf B=64 S=257 H=16: fwd {e0.elapsed_time(e1)/10:.3f} ms ({fl/(e0.elapsed_time(e1)/10)/1e9:.1f} TF/s)  bwd {e1.elapsed_time(e2)/10:.3f} ms ({2.5*fl/(e1.elapsed_time(e2)/10)/1e9:.1f} TF/s)")
    return ok


def case_infonce():
    ok = True
    F = torch.nn.functional
    for (bs, W, E, rank) in [(4, 1, 64, 0), (100, 1, 512, 0), (64, 4, 768, 2), (512, 8, 768, 5)]:
        n = bs * W
        emb = torch.randn(2 * bs, E, device=dev)
        q_idx = torch.arange(0, 2 * bs, 2, device=dev).int(); p_idx = q_idx + 1
        ls = torch.tensor(4.6052, device=dev)
        nce = ops.InfoNCE(bs, n, E, dev)
        others = F.normalize(torch.randn(n, E, device=dev), dim=-1)
        nce.all_p.copy_(others)
        off = rank * bs
        nce.normalize(emb, q_idx, p_idx, off)
        for rep in range(2):  # twice: the workspace must self-reset
            loss, acc = nce.forward(ls, off)
        er = emb.clone().requires_grad_(True); lr_ = ls.clone().requires_grad_(True)
        q = F.normalize(er[q_idx.long()], dim=-1); p = F.normalize(er[p_idx.long()], dim=-1)
        allp = others.clone(); allp = torch.cat([allp[:off], p, allp[off + bs:]], 0)
        allp_leaf = allp.detach().requires_grad_(True)
        score = q @ allp_leaf.t() * lr_.exp()
        tgt = off + torch.arange(bs, device=dev)
        rl = F.cross_entropy(score, tgt); ra = (score.argmax(1) == tgt).float().mean()
        rl.backward()
        dls = torch.zeros((), device=dev)
        d_qn, d_allp = nce.backward(ls, off, dls)
        r1, _ = rel_err(d_allp, allp_leaf.grad)
        # full chain for local rows: d_emb
        d_emb = torch.zeros(2 * bs, E, device=dev)
        nce.normalize_bwd(q_idx, p_idx, d_allp[off:off + bs], d_emb, off)
        er2 = emb.clone().requires_grad_(True)
        q2 = F.normalize(er2[q_idx.long()], dim=-1); p2 = F.normalize(er2[p_idx.long()], dim=-1)
        allp2 = torch.cat([others[:off], p2, others[off + bs:]], 0)
        F.cross_entropy(q2 @ allp2.t() * ls.exp(), tgt).backward()
        r2, _ = rel_err(d_emb, er2.grad)
        good = abs(loss.item() - rl.item()) < 1e-4 * max(1, abs(rl.item())) and abs(acc.item() - ra.item()) < 1e-6 and r1 < 1e-4 and r2 < 1e-4 and abs(dls.item() - lr_.grad.item()) < 1e-3 * max(1, abs(lr_.grad.item()))
        print(f"infonce bs={bs} W={W} E={E}: loss {loss.item():.6f} vs {rl.item():.6f} acc {acc.item():.4f} vs {ra.item():.4f} d_allp={r1:.2e} d_emb={r2:.2e} dls {dls.item():.5f} vs {lr_.grad.item():.5f} {'OK' if good else 'FAIL'}", flush=True)
        ok &= good
    return ok


def _make_batch(dims, bs, seed=2023, device="cpu"):
    from uniir_b200.preprocess import synthetic_clip_tokens
    g = torch.Generator().manual_seed(seed)
    n = 2 * bs
    img = torch.randn(n, 3, dims.image_size, dims.image_size, generator=g)
    ids = synthetic_clip_tokens(n, g, dims.context, dims.vocab)
    mt = torch.ones(n, dtype=torch.long); mv = torch.ones(n, dtype=torch.long)
    mt[1::5] = 0; mv[2::7] = 0
    return dict(txt_batched=ids, image_batched=img, txt_mask_batched=mt, image_mask_batched=mv,
                index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})


def model_case(name, bs, mb):
    from oracle.clip_oracle import CLIPScoreFusionOracle
    from uniir_b200.clip_sf import CLIPScoreFusion
    torch.manual_seed(0)
    orc = CLIPScoreFusionOracle(name)
    mine = CLIPScoreFusion(name, device="cuda", micro_batch=mb)
    missing = mine.load_state_dict(orc.state_dict(), strict=True)
    batch = _make_batch(orc.clip_model.dims, bs)
    t0 = time.time()
    out_o = orc(batch)
    out_o["loss"].backward()
    t_cpu = time.time() - t0
    with torch.no_grad():
        emb_o = orc.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"], batch["image_mask_batched"])
    gb = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    emb_m = mine.encode_multimodal_input(gb["txt_batched"], gb["image_batched"], gb["txt_mask_batched"], gb["image_mask_batched"]).cpu()
    cos = torch.nn.functional.cosine_similarity(emb_m, emb_o, dim=-1)
    out_m = mine(gb)
    out_m["loss"].backward()
    torch.cuda.synchronize()
    print(f"model {name} bs={bs} mb={mb}: emb cos min={cos.min().item():.6f} mean={cos.mean().item():.6f} | loss {out_m['loss'].item():.5f} vs {out_o['loss'].item():.5f} acc {out_m['accuracy'].item():.3f} vs {out_o['accuracy'].item():.3f} (cpu {t_cpu:.1f}s)")
    ok = cos.min().item() > 1 - 1e-3 and abs(out_m["loss"].item() - out_o["loss"].item()) < 2e-2 * max(1.0, abs(out_o["loss"].item()))
    worst = []
    for (n_, p_o) in orc.named_parameters():
        p_m = mine.get_parameter(n_)
        if p_o.grad is None or p_m.grad is None:
            print("  no grad for", n_, p_o.grad is None, p_m.grad is None); ok = False; continue
        go, gm = p_o.grad.flatten().double(), p_m.grad.cpu().flatten().double()
        c = (torch.dot(go, gm) / (go.norm() * gm.norm() + 1e-30)).item() if go.numel() > 1 else 1.0 - abs((go - gm).item()) / (abs(go.item()) + 1e-12)
        rn = (gm.norm() / (go.norm() + 1e-30)).item()
        worst.append((c, rn, n_))
    worst.sort()
    for c, rn, n_ in worst[:6]:
        print(f"  grad cos={c:.5f} norm_ratio={rn:.4f} {n_}")
    mc = sum(c for c, _, _ in worst) / len(worst)
    print(f"  grads: {len(worst)} tensors, mean cos={mc:.5f}, min cos={worst[0][0]:.5f}")
    ok &= worst[0][0] > 0.98
    return ok


def case_model_tiny():
    return model_case("tiny-test", 8, 8) & model_case("tiny-test", 12, 8)


def case_model_b32():
    return model_case("ViT-B/32", 4, 8)


def case_train_graph():
    """graphs vs eager: identical loss trajectory over 3 steps on the tiny model + timing on L/14"""
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.engine import FlatAdamW, train_step
    ok = True
    res = {}
    for use in (False, True):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=1)
        m.use_cuda_graphs = use
        opt = FlatAdamW(m.clip_model, lr=1e-3)
        b = _make_batch(m.dims, 16)
        gb = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
        res[use] = [train_step(m, gb, opt)[0].item() for _ in range(5)]
    print("eager :", res[False]); print("graphs:", res[True])
    ok &= all(abs(a - b) < 1e-3 * max(1, abs(a)) for a, b in zip(res[False], res[True])) and res[True][-1] < res[True][0]
    # L/14 micro timing: 256 pairs with mb 64 -> 8 micro-batches
    m = CLIPScoreFusion("ViT-L/14", device="cuda", micro_batch=64)
    opt = FlatAdamW(m.clip_model, lr=1e-5)
    import bench
    b = bench.synth_batch(m.dims, 256, 1, torch.device("cuda"))
    for _ in range(3): train_step(m, b, opt)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(2): train_step(m, b, opt)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 2
    print(f"L/14 256 pairs/step with graphs: {dt*1e3:.1f} ms -> {256/dt:.1f} pairs/s ({256*1.052e12/dt/1e12:.1f} algorithmic TFLOP/s)")
    return ok


def case_gemm_perf2():
    """model-shape GEMMs with their real epilogues (M = 64 items x 257 tokens)"""
    M = 16448
    f = dict(device=dev)
    cfgs = [
        ("qkv fwd  bias", dict(M=M, N=3072, K=1024, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=True, act=0, residual=None, aux=False, ags=False, accumulate=False, split_k=1)),
        ("out fwd  bias+res32", dict(M=M, N=1024, K=1024, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.float32, bias=True, act=0, residual=torch.float32, aux=False, ags=False, accumulate=False, split_k=1)),
        ("fc1 fwd  bias+qgelu+aux", dict(M=M, N=4096, K=1024, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=True, act=1, residual=None, aux=True, ags=False, accumulate=False, split_k=1)),
        ("fc1 fwd  bias+qgelu", dict(M=M, N=4096, K=1024, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=True, act=1, residual=None, aux=False, ags=False, accumulate=False, split_k=1)),
        ("fc2 fwd  bias+res32", dict(M=M, N=1024, K=4096, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.float32, bias=True, act=0, residual=torch.float32, aux=False, ags=False, accumulate=False, split_k=1)),
        ("fc2 dgrad qgelu'", dict(M=M, N=4096, K=1024, a_mn=False, b_mn=True, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=False, act=1, residual=None, aux=False, ags=True, accumulate=False, split_k=1)),
        ("fc1 dgrad", dict(M=M, N=1024, K=4096, a_mn=False, b_mn=True, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=False, act=0, residual=None, aux=False, ags=False, accumulate=False, split_k=1)),
        ("fc wgrad", dict(M=4096, N=1024, K=M, a_mn=True, b_mn=True, dtype=torch.bfloat16, out_dtype=torch.float32, bias=False, act=0, residual=None, aux=False, ags=False, accumulate=True, split_k=1)),
        ("out wgrad split", dict(M=1024, N=1024, K=M, a_mn=True, b_mn=True, dtype=torch.bfloat16, out_dtype=torch.float32, bias=False, act=0, residual=None, aux=False, ags=False, accumulate=True, split_k=4)),
        ("txt qkv fwd", dict(M=4928, N=2304, K=768, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.bfloat16, bias=True, act=0, residual=None, aux=False, ags=False, accumulate=False, split_k=1)),
        ("txt fc2 fwd", dict(M=4928, N=768, K=3072, a_mn=False, b_mn=False, dtype=torch.bfloat16, out_dtype=torch.float32, bias=True, act=0, residual=torch.float32, aux=False, ags=False, accumulate=False, split_k=1)),
    ]
    for name, c in cfgs:
        ms = ops.time_gemm_config(c, iters=20)
        print(f"{name:28s} M={c['M']:6d} N={c['N']:5d} K={c['K']:6d}: {ms*1e3:8.1f} us  {2.0*c['M']*c['N']*c['K']/ms/1e9:7.1f} TF/s", flush=True)
    return True


def case_train_speed():
    """L/14 step speed at a few micro-batch sizes (env MBS='64,128')"""
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.engine import FlatAdamW, train_step
    import bench
    for mb in [int(x) for x in os.environ.get("MBS", "64,128").split(",")]:
        m = CLIPScoreFusion("ViT-L/14", device="cuda", micro_batch=mb)
        opt = FlatAdamW(m.clip_model, lr=1e-5)
        pairs = 512
        b = bench.synth_batch(m.dims, pairs, 1, torch.device("cuda"))
        for _ in range(3): train_step(m, b, opt)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(2): l, a = train_step(m, b, opt)
        torch.cuda.synchronize(); dt = (time.time() - t0) / 2
        print(f"L/14 mb={mb}: {pairs} pairs/step {dt*1e3:.1f} ms -> {pairs/dt:.1f} pairs/s ({pairs*1.052e12/dt/1e12:.1f} alg TFLOP/s) loss {l.item():.4f}", flush=True)
        del m, opt, b
        torch.cuda.empty_cache()
    return True


if __name__ == "__main__":
    name = sys.argv[1]
    ok = globals()["case_" + name]()
    print(f"CASE {name}: {'PASS' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)
