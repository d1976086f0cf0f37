"""GPU parity of the individual C-ABI kernels against plain fp32 CPU references (torch CPU): tcgen05
GEMM in every operand-major mode and epilogue, LayerNorm fwd/bwd, embedding kernels, attention
fwd/bwd, AdamW.  Tolerances are bf16-input tolerances, stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from uniir_b200 import ops
    return ops


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (1000, 768, 1024), (333, 4096, 768), (1024, 776, 1000)])
def test_gemm_all_operand_majors(ops, a_mn, b_mn, M, N, K):
    if a_mn and M % 8:
        M += 8 - M % 8
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).bfloat16()
    ref = (A.t() if a_mn else A).float() @ (B.t() if b_mn else B).float().t()
    out = ops.gemm(A.to(dev), B.to(dev), a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    assert rel(out, ref) < 1e-5  # exact products, fp32 accumulate: only summation order differs


def test_gemm_fp16_and_split_k(ops):
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(512, 512, generator=g).half(), torch.randn(512, 512, generator=g).half()
    assert rel(ops.gemm(A.to(dev), B.to(dev), out_dtype=torch.float32), A.float() @ B.float().t()) < 1e-5
    X, dY = torch.randn(4096, 256, generator=g).bfloat16(), torch.randn(4096, 384, generator=g).bfloat16()
    acc = torch.ones(384, 256, device=dev)
    ops.wgrad(dY.to(dev), X.to(dev), acc)  # narrow output -> split-K atomics
    assert rel(acc - 1, dY.float().t() @ X.float()) < 1e-4
    acc2 = torch.ones(384, 256, device=dev)
    ops.gemm(dY.to(dev), X.to(dev), a_mn=True, b_mn=True, out=acc2, accumulate=True)
    assert rel(acc2 - 1, dY.float().t() @ X.float()) < 1e-5


@pytest.mark.parametrize("dyn", [False, True])
def test_gemm_epilogues(ops, dyn, monkeypatch):
    from uniir_b200 import lib as L
    if dyn:
        monkeypatch.setenv("UNIIR_GEMM_DYN", "1")  # runtime-flag epilogue instead of the static variants
    g = torch.Generator().manual_seed(2)
    M, N, K = 777, 1024, 512
    A = torch.randn(M, K, generator=g).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    pre = A.float() @ B.float().t() + bias
    Ad, Bd, bd = A.to(dev), B.to(dev), bias.to(dev)
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = ops.gemm(Ad, Bd, bias=bd, act=L.ACT_QUICKGELU, aux_out=aux)
    assert rel(out, pre * torch.sigmoid(1.702 * pre)) < 5e-3 and rel(aux, pre) < 5e-3   # bf16 outputs
    assert rel(ops.gemm(Ad, Bd, bias=bd, act=L.ACT_QUICKGELU), pre * torch.sigmoid(1.702 * pre)) < 5e-3
    assert rel(ops.gemm(Ad, Bd, bias=bd, act=L.ACT_GELU), F.gelu(pre)) < 5e-3
    x = res.to(dev)
    ops.gemm(Ad, Bd, bias=bd, residual=x, out=x)                     # in-place fp32 residual stream
    assert rel(x, pre + res) < 1e-5
    src = torch.randn(M, N, generator=g).bfloat16()
    Bt = B.t().contiguous()                                         # dgrad form: B stored [K, N]
    out = ops.gemm(Ad, Bt.to(dev), b_mn=True, act=L.ACT_QUICKGELU, act_grad_src=src.to(dev), out_dtype=torch.float32)
    s = src.float(); sg = torch.sigmoid(1.702 * s)
    assert rel(out, (A.float() @ B.float().t()) * (sg * (1 + 1.702 * s * (1 - sg)))) < 1e-3  # tanh.approx sigmoid
    acc = torch.randn(M, N, generator=g); accd = acc.to(dev)
    ops.gemm(Ad, Bd, out=accd, accumulate=True, alpha=0.5)
    assert rel(accd, acc + 0.5 * (A.float() @ B.float().t())) < 1e-5


def test_gemm_rejects_bad_arguments(ops):
    from uniir_b200.lib import UniirError
    A = torch.zeros(16, 12, device=dev, dtype=torch.bfloat16)   # K=12 not a multiple of 8
    with pytest.raises(UniirError):
        ops.gemm(A, A)


@pytest.mark.parametrize("D,rows", [(1024, 300), (768, 77), (128, 50), (512, 9), (1024, 2571), (64, 1), (2048, 37), (36, 19)])
def test_layernorm_fwd_bwd(ops, D, rows):
    """D <= 1024 runs the warp-per-row kernels, wider rows the CTA-per-row ones; both must match torch."""
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g) * 2 + 0.5
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    y, mean, rstd = ops.layernorm_fwd(x.to(dev), w.to(dev), b.to(dev), 1e-5, out_dtype=torch.float32, save_stats=True)
    assert rel(y, F.layer_norm(x, (D,), w, b, 1e-5)) < 1e-6
    assert rel(ops.layernorm_fwd(x.to(dev), w.to(dev), b.to(dev), 1e-5), F.layer_norm(x, (D,), w, b, 1e-5)) < 5e-3
    dy = torch.randn(rows, D, generator=g).bfloat16()
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    F.layer_norm(xr, (D,), wr, br, 1e-5).backward(dy.float())
    dx = torch.ones(rows, D, device=dev); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy.to(dev), x.to(dev), mean, rstd, w.to(dev), dx, dg, db)
    assert rel(dx - 1, xr.grad) < 1e-5 and rel(dg, wr.grad) < 1e-5 and rel(db, br.grad) < 1e-5
    # fused bias-gradient output: column sums of the updated dx (here dx = 1 + LN grad), plus the bf16 copy;
    # fp32 dy and accumulate=False exercise the other load / store variants
    dx2 = torch.ones(rows, D, device=dev); dxb = torch.zeros(rows, D, device=dev, dtype=torch.bfloat16)
    dg2 = torch.zeros(D, device=dev); db2 = torch.zeros(D, device=dev); dc = torch.full((D,), 2.0, device=dev)
    ops.layernorm_bwd(dy.to(dev), x.to(dev), mean, rstd, w.to(dev), dx2, dg2, db2, dx_bf16=dxb, dcolsum=dc)
    assert rel(dx2, dx) < 1e-6 and rel(dg2, wr.grad) < 1e-5 and rel(db2, br.grad) < 1e-5   # template variants may contract FMAs differently
    assert rel(dxb, dx) < 5e-3
    tol = 1e-5 if D <= 1024 else 3e-3            # wide rows take the column sums from the bf16 copy
    assert ((dc - 2).cpu() - (1 + xr.grad).sum(0)).abs().max() < tol * (1 + xr.grad).abs().sum(0).max()
    dx3 = torch.full((rows, D), 7.0, device=dev); dg3 = torch.zeros(D, device=dev); db3 = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy.float().to(dev), x.to(dev), mean, rstd, w.to(dev), dx3, dg3, db3, accumulate=False)
    assert rel(dx3, xr.grad) < 1e-5 and rel(dg3, wr.grad) < 1e-5


@pytest.mark.parametrize("M,N,ld", [(2571, 3072, 3072), (100, 4096, 4096), (33, 768, 2304), (7, 64, 64), (1000, 36, 36), (513, 264, 264)])
def test_colsum_accum(ops, M, N, ld):
    """Bias gradient = column sums of bf16 dy; N % 8 == 0 takes the 128-bit kernel (also on a strided view)."""
    g = torch.Generator().manual_seed(M + N)
    full = torch.randn(M, ld, generator=g).bfloat16()
    dy = full[:, :N]
    out = torch.full((N,), 3.0, device=dev)
    ops.colsum_accum(full.to(dev)[:, :N], out)
    ref = dy.float().sum(0)
    assert ((out - 3).cpu() - ref).abs().max() < 1e-4 * dy.float().abs().sum(0).max()


def test_layernorm_gathered_rows_and_embed_kernels(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(10, 7, 128, generator=g); w, b = torch.randn(128, generator=g), torch.randn(128, generator=g)
    idx = (torch.arange(10) * 7 + torch.randint(0, 7, (10,), generator=g)).int()
    y = ops.layernorm_fwd(x.to(dev), w.to(dev), b.to(dev), 1e-5, row_index=idx.to(dev), out_dtype=torch.float32)
    assert rel(y, F.layer_norm(x.view(70, 128)[idx.long()], (128,), w, b, 1e-5)) < 1e-6
    img = torch.randn(3, 3, 56, 56, generator=g)
    col = ops.patch_im2col(img.to(dev), 14, 640)
    assert rel(col[:, :588], F.unfold(img, 14, stride=14).transpose(1, 2).reshape(48, 588)) < 5e-3
    assert col[:, 588:].abs().max().item() == 0
    patch = torch.randn(48, 128, generator=g).bfloat16(); cls = torch.randn(128, generator=g); pos = torch.randn(17, 128, generator=g)
    x0 = ops.vision_embed_fwd(patch.to(dev), cls.to(dev), pos.to(dev), 3, 17, 128)
    assert rel(x0, torch.cat([cls.expand(3, 1, 128), patch.float().view(3, 16, 128)], 1) + pos) < 1e-6
    ids = torch.randint(1, 500, (5, 77), generator=g).int(); ids[:, 0] = 510
    ids[torch.arange(5), torch.tensor([3, 10, 76, 5, 1])] = 511
    tok, tpos = torch.randn(512, 64, generator=g), torch.randn(77, 64, generator=g)
    xt, eot = ops.text_embed_fwd(ids.to(dev), tok.to(dev), tpos.to(dev))
    assert rel(xt, tok[ids.long()] + tpos) < 1e-6
    assert (eot.cpu().long() == torch.arange(5) * 77 + ids.argmax(-1)).all()   # EOT pooling row = argmax(token id)
    dxt = torch.randn(5, 77, 64, generator=g); dtok = torch.zeros(512, 64, device=dev); dpos = torch.zeros(77, 64, device=dev)
    ops.text_embed_bwd(dxt.to(dev), ids.to(dev), dtok, dpos)
    assert rel(dtok, torch.zeros(512, 64).index_add_(0, ids.long().flatten(), dxt.view(-1, 64))) < 1e-5
    assert rel(dpos, dxt.sum(0)) < 1e-6


def _attn_ref(qkv, B, S, H, causal):
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = q @ k.transpose(-1, -2) * 0.125
    if causal:
        att = att + torch.full((S, S), float("-inf")).triu_(1)
    return (att.softmax(-1) @ v).transpose(1, 2).reshape(B * S, H * 64), torch.logsumexp(att, -1)


@pytest.mark.parametrize("B,S,H,causal", [(2, 64, 1, False), (3, 257, 4, False), (5, 77, 3, True), (2, 197, 2, False),
                                           (2, 50, 2, False), (1, 130, 2, True), (2, 100, 1, False)])
def test_attention_fwd_bwd(ops, B, S, H, causal):
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 1.5).bfloat16()
    out, lse = ops.attn_fwd(qkv.to(dev), B, S, H, causal)
    qr = qkv.float().requires_grad_(True)
    ref, rlse = _attn_ref(qr, B, S, H, causal)
    assert rel(out, ref) < 6e-3 and (lse.cpu() - rlse).abs().max() < 1e-3
    dout = torch.randn(B * S, H * 64, generator=g).bfloat16()
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv.to(dev), out, dout.to(dev), lse, B, S, H, causal)
    assert rel(dqkv, qr.grad) < 1e-2   # bf16 P / dS operands


@pytest.mark.parametrize("B,S,H,causal", [(2, 64, 1, False), (3, 257, 4, False), (5, 77, 3, True), (2, 197, 2, False),
                                           (2, 50, 2, False), (1, 130, 2, True), (2, 100, 1, False), (1, 132, 1, False),
                                           (1, 272, 2, False), (1, 304, 2, False), (20, 257, 16, False), (40, 77, 12, True), (1, 128, 1, True)])
def test_attention_fwd_tcgen05(ops, B, S, H, causal):
    """The tcgen05 forward (attention_tc.cu) against the fp32 reference and against the mma.sync kernel: one
    128-query tile, several tiles, the 1..4-row CUDA-core tail (257, 130, 132), persistence over more heads than
    SMs (320 and 480 heads: K/V double buffering, every barrier wraps), causal masking."""
    g = torch.Generator().manual_seed(S + B)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 1.5).bfloat16()
    ref, rlse = _attn_ref(qkv.float(), B, S, H, causal)
    prev = ops.attn_set_mode(0)
    try:
        out0, lse0 = ops.attn_fwd(qkv.to(dev), B, S, H, causal)
    finally:
        ops.attn_set_mode(prev)
    prev = ops.attn_set_mode(5)
    try:
        out1, lse1 = ops.attn_fwd(qkv.to(dev), B, S, H, causal)
        torch.cuda.synchronize()
    finally:
        ops.attn_set_mode(prev)
    assert rel(out1, ref) < 6e-3 and (lse1.cpu() - rlse).abs().max() < 1e-3
    assert rel(out1, out0) < 6e-3 and (lse1 - lse0).abs().max() < 1e-3


def test_attention_fwd_tcgen05_cross_attention_with_key_padding(ops):
    """BLIP shapes: text queries (100) over image keys (197), and BERT self-attention with a key-padding mask;
    q / k / v are column slices of packed projections (row stride != 64 H)."""
    g = torch.Generator().manual_seed(5)
    B, Sq, Skv, H = 3, 100, 197, 12
    q = (torch.randn(B * Sq, H * 64, generator=g)).bfloat16().to(dev)
    kv = (torch.randn(B * Skv, 2 * H * 64, generator=g)).bfloat16().to(dev)
    qkv = (torch.randn(B * Sq, 3 * H * 64, generator=g)).bfloat16().to(dev)
    kvl = torch.tensor([100, 17, 1], dtype=torch.int32, device=dev)
    res = {}
    for mode in (0, 5):
        prev = ops.attn_set_mode(mode)
        try:
            lse_x = torch.empty(B, H, Sq, device=dev); lse_s = torch.empty(B, H, Sq, device=dev)
            ox = ops.attn_fwd_ex(q, kv[:, :H * 64], kv[:, H * 64:], B, Sq, Skv, H, lse=lse_x)
            os_ = ops.attn_fwd_ex(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], B, Sq, Sq, H, kv_len=kvl, lse=lse_s)
            torch.cuda.synchronize()
        finally:
            ops.attn_set_mode(prev)
        res[mode] = (ox, lse_x, os_, lse_s)
    assert rel(res[5][0], res[0][0]) < 6e-3 and (res[5][1] - res[0][1]).abs().max() < 1e-3
    assert rel(res[5][2], res[0][2]) < 6e-3 and (res[5][3] - res[0][3]).abs().max() < 1e-3
    # and against fp32 for the cross-attention case
    qf = q.float().cpu().view(B, Sq, H, 64).transpose(1, 2); kf = kv[:, :H * 64].float().cpu().view(B, Skv, H, 64).transpose(1, 2)
    vf = kv[:, H * 64:].float().cpu().view(B, Skv, H, 64).transpose(1, 2)
    ref = ((qf @ kf.transpose(-1, -2) * 0.125).softmax(-1) @ vf).transpose(1, 2).reshape(B * Sq, H * 64)
    assert rel(res[5][0], ref) < 6e-3


@pytest.mark.parametrize("B,S,H,causal", [(1, 64, 1, False), (3, 257, 4, False), (5, 77, 3, True), (2, 197, 2, False),
                                           (2, 50, 2, False), (1, 130, 2, True), (2, 100, 1, False), (1, 132, 1, False),
                                           (1, 256, 2, False), (20, 257, 16, False), (40, 77, 12, True), (1, 128, 1, True)])
def test_attention_bwd_tcgen05(ops, B, S, H, causal):
    """The tcgen05 backward (attention_bwd_tc.cu) against the fp32 autograd reference and the mma.sync kernels:
    one and two 128-key tiles, partial 64-query chunks, the CUDA-core tail of 257 / 130 / 132 (key columns and
    query rows beyond the last full tile, prep kernel), persistence over more heads than SMs, causal masks."""
    g = torch.Generator().manual_seed(S + B)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * S, H * 64, generator=g).bfloat16()
    qr = qkv.float().requires_grad_(True)
    ref, _ = _attn_ref(qr, B, S, H, causal)
    ref.backward(dout.float())
    prev = ops.attn_set_mode(0)
    try:
        out, lse = ops.attn_fwd(qkv.to(dev), B, S, H, causal)
        g0 = ops.attn_bwd(qkv.to(dev), out, dout.to(dev), lse, B, S, H, causal)
    finally:
        ops.attn_set_mode(prev)
    prev = ops.attn_set_mode(2)
    try:
        g2 = ops.attn_bwd(qkv.to(dev), out, dout.to(dev), lse, B, S, H, causal)
        torch.cuda.synchronize()
    finally:
        ops.attn_set_mode(prev)
    D = H * 64
    for i in range(3):   # dq, dk, dv separately
        assert rel(g2[:, i * D:(i + 1) * D], qr.grad[:, i * D:(i + 1) * D]) < 1e-2   # bf16 P / dS operands
        assert rel(g2[:, i * D:(i + 1) * D], g0[:, i * D:(i + 1) * D]) < 1e-2


def test_attention_bwd_tcgen05_cross_attention_with_key_padding(ops):
    """BLIP shapes: text queries (100) over image keys (197) and BERT self-attention with a key-padding mask, on
    column slices of packed projections; tcgen05 backward against the mma.sync kernels."""
    g = torch.Generator().manual_seed(6)
    B, Sq, Skv, H = 3, 100, 197, 12
    D = H * 64
    q = torch.randn(B * Sq, D, generator=g).bfloat16().to(dev)
    kv = torch.randn(B * Skv, 2 * D, generator=g).bfloat16().to(dev)
    qkv = torch.randn(B * Sq, 3 * D, generator=g).bfloat16().to(dev)
    dout = torch.randn(B * Sq, D, generator=g).bfloat16().to(dev)
    kvl = torch.tensor([100, 17, 1], dtype=torch.int32, device=dev)
    lse_x = torch.empty(B, H, Sq, device=dev); lse_s = torch.empty(B, H, Sq, device=dev)
    ox = ops.attn_fwd_ex(q, kv[:, :D], kv[:, D:], B, Sq, Skv, H, lse=lse_x)
    os_ = ops.attn_fwd_ex(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, Sq, Sq, H, kv_len=kvl, lse=lse_s)
    res = {}
    for mode in (0, 2):
        prev = ops.attn_set_mode(mode)
        try:
            dq = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16); dkv = torch.zeros(B * Skv, 2 * D, device=dev, dtype=torch.bfloat16)
            ops.attn_bwd_ex(q, kv[:, :D], kv[:, D:], ox, dout, lse_x, B, Sq, Skv, H, dq, dkv[:, :D], dkv[:, D:])
            dqkv = torch.zeros(B * Sq, 3 * D, device=dev, dtype=torch.bfloat16)
            ops.attn_bwd_ex(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], os_, dout, lse_s, B, Sq, Sq, H,
                            dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], kv_len=kvl)
            torch.cuda.synchronize()
        finally:
            ops.attn_set_mode(prev)
        res[mode] = (dq, dkv, dqkv)
    for a, b in zip(res[2], res[0]):
        assert rel(a, b) < 1e-2


def test_adamw_matches_torch(ops):
    g = torch.Generator().manual_seed(4)
    n = 100003
    p = torch.randn(n, generator=g); pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    pd = p.to(dev); gd = torch.zeros(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    p16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        gg = torch.randn(n, generator=g); pr.grad = gg.clone(); opt.step()
        gd.copy_(gg); ops.adamw_step(pd, gd, m, v, p16, 1e-3, (0.9, 0.98), 1e-6, 0.2, step)
    assert rel(pd, pr.detach()) < 1e-6 and gd.abs().max().item() == 0 and rel(p16, pd) < 5e-3
