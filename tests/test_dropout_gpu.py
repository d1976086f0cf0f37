"""GPU parity of the train-mode dropout path: the counter-based mask kernels against their integer restatement
(tests/cpu_kernels.py), the elementwise / DropPath / row-scale / GEMM-epilogue kernels, attention with dropout on the
probabilities (forward and backward, self / cross / key padding / score bias), and the BLIP_FF training step in train
mode against the oracle fed the same masks (the oracle's dropout SITES are pinned to the reference's modules in
tests/test_oracle_blip.py::test_dropout_sites_match_reference_modules_in_train_mode)."""
import pytest
import torch

import cpu_kernels as ck

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _ctx(ops, seed, item0):
    c = ops.drop_ctx(dev)
    c.copy_(torch.tensor([seed, item0], dtype=torch.int32))
    return c


@pytest.mark.parametrize("seed,site,item0,per_item,n,p", [(1, 0, 0, 64, 1000, 0.1), (2 ** 31 - 5, 61, 3, 1536, 40_001, 0.1),
                                                           (123456, 4099, 4000, 59136, 300_000, 0.5), (7, 3, 0, 8, 64, 0.0)])
def test_mask_kernel_is_the_documented_integer_function(seed, site, item0, per_item, n, p):
    from uniir_b200 import ops
    keep = ops.dropout_mask(_ctx(ops, seed, item0), site, per_item, n, p).cpu()
    ref = ck.keep_mask(seed, site, torch.arange(n, dtype=torch.int64) + item0 * per_item, p)
    assert torch.equal(keep.bool(), ref)
    if n >= 40_000:
        assert abs(1 - keep.float().mean().item() - p) < 0.01        # drop rate
    if p == 0.0:
        assert keep.all()


def test_elementwise_dropout_forward_and_gradient_forms():
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(0)
    rows, H, seed, site, item0, p = 96, 768, 99, 12, 16, 0.1
    x = torch.randn(rows, H, generator=g); res = torch.randn(rows, H, generator=g)
    ctx = _ctx(ops, seed, item0)
    keep = ck.keep_mask(seed, site, torch.arange(rows * H) + item0 * 24 * H, p).view(rows, H)
    want = torch.where(keep, x * ck.drop_scale(p), torch.zeros(()))
    y = ops.dropout(x.to(dev), ctx, site, 24 * H, p, residual=res.to(dev))                  # in place, + residual (forward)
    assert torch.equal(y.cpu(), want + res)
    yb = torch.empty(rows, H, device=dev, dtype=torch.bfloat16)
    ops.dropout(x.to(dev), ctx, site, 24 * H, p, out=yb)                                     # fp32 -> bf16 (backward)
    assert torch.equal(yb.cpu(), want.bfloat16())
    xb = x.bfloat16().to(dev)
    assert torch.equal(ops.dropout(xb, ctx, site, 24 * H, p).cpu(), torch.where(keep, x.bfloat16().float() * ck.drop_scale(p), torch.zeros(())).bfloat16())
    odd = torch.randn(1003, generator=g)                                                     # length not a multiple of 4
    k2 = ck.keep_mask(seed, 5, torch.arange(1003) + item0 * 1003, 0.3)
    assert torch.equal(ops.dropout(odd.to(dev), ctx, 5, 1003, 0.3).cpu(), torch.where(k2, odd * ck.drop_scale(0.3), torch.zeros(())))


def test_droppath_scales_row_scale_and_gemm_epilogue():
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(1)
    seed, item0, mb, S = 4242, 8, 16, 50
    p_sites = torch.linspace(0, 0.3, 4).repeat_interleave(2).contiguous()
    ctx = _ctx(ops, seed, item0)
    sc = ops.droppath_scales(ctx, 4096, p_sites.to(dev), mb).cpu()
    for s_, p in enumerate(p_sites.tolist()):
        keep = ck.keep_mask(seed, 4096 + s_, torch.arange(mb) + item0, p)
        assert torch.equal(sc[s_], keep.float() * ck.drop_scale(p))
    assert (sc[:2] == 1).all() and (sc[6:] == 0).any() and (sc[6:] > 1).any()
    # branch gradient: rows of an item scaled together
    x = torch.randn(mb * S, 128, generator=g).bfloat16()
    got = ops.row_scale(x.to(dev).clone(), sc[7].to(dev), S).cpu()
    assert torch.equal(got, (x.float() * sc[7].repeat_interleave(S)[:, None]).bfloat16())
    # residual-branch GEMM: out = res + scale[item] * (A W^T + b), single-CTA and CTA-pair tile paths, static and dynamic epilogue
    for M_items, tokens, N, K in ((16, 50, 128, 256), (64, 197, 1024, 1024)):
        M = M_items * tokens
        a = torch.randn(M, K, generator=g).bfloat16(); w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
        b = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
        s = (torch.rand(M_items, generator=g) > 0.3).float() / 0.7
        want = res + s.repeat_interleave(tokens)[:, None] * (a.float() @ w.float().t() + b)
        out = ops.gemm(a.to(dev), w.to(dev), bias=b.to(dev), residual=res.to(dev), out_dtype=torch.float32,
                       row_scale=s.to(dev), row_scale_rows=tokens)
        assert rel(out, want) < 2e-3
        dropped = (s == 0).repeat_interleave(tokens)
        assert torch.equal(out.cpu()[dropped], res[dropped])             # a dropped branch leaves the residual untouched


def _attn_ref(q, k, v, B, Sq, Skv, H, scale, kv_len, bias, mask):
    sp = lambda t, S: t.view(B, S, H, 64).transpose(1, 2)
    s = sp(q, Sq) @ sp(k, Skv).transpose(-1, -2) * scale
    if bias is not None:
        s = s + bias
    if kv_len is not None:
        s = s.masked_fill(torch.arange(Skv)[None, None, None, :] >= kv_len.long()[:, None, None, None], float("-inf"))
    return ((s.softmax(-1) * mask) @ sp(v, Skv)).transpose(1, 2).reshape(B * Sq, H * 64)


@pytest.mark.parametrize("B,Sq,Skv,H,padded,biased", [(3, 40, 40, 2, True, False), (2, 77, 197, 12, False, False),
                                                       (2, 24, 50, 1, False, False), (2, 100, 100, 3, False, True),
                                                       (4, 334, 334, 2, False, True)])
def test_attention_with_probability_dropout(B, Sq, Skv, H, padded, biased):
    """O = dropout(softmax(S)) V and its backward with the mask of (seed, site, batch position): the cases of the BLIP
    BERT stack (self-attention with key padding, cross-attention into 197 image tokens) and of the CLIP_FF T5 stack
    (score bias, scale 1, S = 334); item0 != 0 as for a later micro-batch of the batch."""
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(B * Sq + Skv)
    seed, site, item0, p = 31337, 1 + 5 * 3 + 2, 5, 0.1
    D = H * 64
    q = torch.randn(B * Sq, D, generator=g).bfloat16(); kv = torch.randn(B * Skv, 2 * D, generator=g).bfloat16()
    dout = torch.randn(B * Sq, D, generator=g).bfloat16()
    kvl = torch.tensor([Skv, 7, 1, Skv - 3][:B], dtype=torch.int32) if padded else None
    bias = torch.randn(H, Sq, Skv, generator=g) if biased else None
    scale = 1.0 if biased else 0.125
    if biased:
        q = (q.float() * 0.125).bfloat16()
    ctx = _ctx(ops, seed, item0)
    idx = torch.arange(B * H * Sq * Skv, dtype=torch.int64) + item0 * H * Sq * Skv
    mask = ck.keep_mask(seed, site, idx, p).view(B, H, Sq, Skv).float() * ck.drop_scale(p)
    qf, kvf = q.float().requires_grad_(True), kv.float().requires_grad_(True)
    bf = bias.clone().requires_grad_(True) if biased else None
    ref = _attn_ref(qf, kvf[:, :D], kvf[:, D:], B, Sq, Skv, H, scale, kvl, bf, mask)
    ref.backward(dout.float())
    qd, kvd = q.to(dev), kv.to(dev)
    lse = torch.empty(B, H, Sq, device=dev)
    drop = (ctx, site, p)
    kl = None if kvl is None else kvl.to(dev)
    if biased:
        out = ops.attn_fwd_bias(qd, kvd[:, :D], kvd[:, D:], B, Sq, Skv, H, bias.to(dev), scale=scale, lse=lse, drop=drop)
    else:
        out = ops.attn_fwd_ex(qd, kvd[:, :D], kvd[:, D:], B, Sq, Skv, H, kv_len=kl, lse=lse, drop=drop)
    assert rel(out, ref.detach()) < 8e-3
    # the log-sum-exp is that of the undropped softmax
    with torch.no_grad():
        sp = lambda t, S: t.view(B, S, H, 64).transpose(1, 2)
        s = sp(q.float(), Sq) @ sp(kv[:, :D].float(), Skv).transpose(-1, -2) * scale
        if biased:
            s = s + bias
        if kvl is not None:
            s = s.masked_fill(torch.arange(Skv)[None, None, None, :] >= kvl.long()[:, None, None, None], float("-inf"))
    assert (lse.cpu() - torch.logsumexp(s, -1)).abs().max().item() < 2e-3
    dq = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16); dkv = torch.zeros(B * Skv, 2 * D, device=dev, dtype=torch.bfloat16)
    if biased:
        dbias = torch.zeros(H, Sq, Skv, device=dev)
        ops.attn_bwd_bias(qd, kvd[:, :D], kvd[:, D:], out, dout.to(dev), lse, B, Sq, Skv, H, bias.to(dev), dbias, dq,
                          dkv[:, :D], dkv[:, D:], scale=scale, drop=drop)
        assert rel(dbias, bf.grad) < 1.5e-2
    else:
        ops.attn_bwd_ex(qd, kvd[:, :D], kvd[:, D:], out, dout.to(dev), lse, B, Sq, Skv, H, dq, dkv[:, :D], dkv[:, D:],
                        kv_len=kl, drop=drop)
    assert rel(dq, qf.grad) < 1.5e-2 and rel(dkv, kvf.grad) < 1.5e-2
    # another seed gives another mask; p = 0 is the plain kernel
    out2 = ops.attn_fwd_ex(qd, kvd[:, :D], kvd[:, D:], B, Sq, Skv, H, kv_len=kl, drop=(_ctx(ops, seed + 1, item0), site, p)) if not biased else None
    if out2 is not None:
        assert rel(out2, out) > 0.05
        out0 = ops.attn_fwd_ex(qd, kvd[:, :D], kvd[:, D:], B, Sq, Skv, H, kv_len=kl, drop=(ctx, site, 0.0))
        prev = ops.attn_set_mode(0)
        try:
            plain = ops.attn_fwd_ex(qd, kvd[:, :D], kvd[:, D:], B, Sq, Skv, H, kv_len=kl)
        finally:
            ops.attn_set_mode(prev)
        assert torch.equal(out0, plain)


def test_blip_ff_training_step_in_train_mode_matches_oracle_with_the_same_masks():
    """Three steps of BLIPFeatureFusion in train mode (hidden dropout 0.1, attention dropout 0.1, DropPath 0.3 on the tiny
    ViT; momentum encoders in train mode too) against the oracle with the product's masks at the reference's sites.
    Phase A encodes the 16 items in one micro-batch, phase C re-forwards them in two of 8, so the masks must follow the
    batch position; step 0 runs eagerly, step 1 captures CUDA graphs, step 2 replays them with new seeds."""
    import oracle.blip_oracle as bo
    from oracle.blip_oracle import BLIPFeatureFusionOracle
    from test_blip_host_logic import _mask_hook
    from test_oracle_blip import _inputs
    from uniir_b200.blip_ff import BLIPFeatureFusion
    torch.manual_seed(5)
    orc = BLIPFeatureFusionOracle("tiny-test", embed_dim=64, queue_size=16).eval()
    with torch.no_grad():
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_:
                p_.mul_(4.0)
    orc.copy_params()
    rate = 0.3
    m = BLIPFeatureFusion(vit="tiny-test", image_size=64, embed_dim=64, queue_size=16, micro_batch=16, seed=3)
    assert m.training and m.hidden_dropout_prob == 0.1 and m.attention_probs_dropout_prob == 0.1 and m.drop_path_rate == 0.0
    m.set_dropout(0.1, 0.1, rate)
    m.train_micro_batch = 8
    missing, unexpected = m.load_state_dict(orc.state_dict(), strict=False)
    assert not unexpected and not missing
    d = orc.dims
    seeds = torch.Generator().manual_seed(3 * 1000003)       # BLIPFeatureFusion._next_drop_seed, rank 0
    plain_loss = None
    for step in range(3):
        ids, att, img = _inputs(d, 16, 24, 20 + step)
        b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=None,
                 image_mask_batched=None, p_did_list=[5, 6, 5, 7, 8, 9, 10, 11 + step],
                 index_mapping={"query": [[2 * i] for i in range(8)], "pos_cand": [[2 * i + 1] for i in range(8)]})
        sd = {"": int(torch.randint(0, 2 ** 31 - 1, (1,), generator=seeds)), "_m": int(torch.randint(0, 2 ** 31 - 1, (1,), generator=seeds))}
        orc.zero_grad()
        if step == 0:
            import copy
            with torch.no_grad():
                plain_loss = copy.deepcopy(orc).compute_contrastive_loss(b, alpha=0.3)["loss"].item()
        bo.DROPOUT = _mask_hook(sd, rate, d.v_layers)
        try:
            ref = orc.compute_contrastive_loss(b, alpha=0.3); ref["loss"].backward()
        finally:
            bo.DROPOUT = None
        if m._online.grad is not None:
            for gbuf in m._online.grad.values():
                gbuf.zero_()
        out = m.compute_contrastive_loss(b, alpha=0.3)
        m.backward_contrastive()
        torch.cuda.synchronize()
        if step == 0:
            assert abs(ref["loss"].item() - plain_loss) > 1e-3          # the masks change the loss
        assert abs(out["loss"].item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item())), (step, out["loss"].item(), ref["loss"].item())
        worst, allcos = (1.0, ""), []
        for n_, p_o in orc.named_parameters():
            if p_o.grad is None or n_.endswith("self.key.bias") or p_o.grad.numel() == 1:
                continue
            a, c = p_o.grad.flatten().double(), m._online.g(n_).cpu().flatten().double()
            cos = (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item()
            allcos.append((round(cos, 5), n_))
            worst = min(worst, (cos, n_))
        assert worst[0] > 0.995, (step, sorted(allcos)[:8])
        assert torch.nn.functional.cosine_similarity(m.cand_queue.cpu().t(), orc.cand_queue.t(), dim=-1).min() > 1 - 1e-3
        with torch.no_grad():
            for n_, p_o in orc.named_parameters():
                if p_o.grad is not None:
                    p_o -= 1e-2 * p_o.grad
                    m.get_parameter(n_).copy_(p_o)
    # eval mode: no dropout, the deterministic encoder
    m.eval()
    ids, att, img = _inputs(d, 4, 24, 50)
    with torch.no_grad():
        ref = orc.encode_multimodal_input(ids, att, img)
    a = m.encode_multimodal_input({"input_ids": ids, "attention_mask": att}, img).cpu()
    b2 = m.encode_multimodal_input({"input_ids": ids, "attention_mask": att}, img).cpu()
    assert torch.equal(a, b2) and torch.nn.functional.cosine_similarity(a, ref, dim=-1).min().item() > 1 - 1e-3
    m.train()
    c = m.encode_multimodal_input({"input_ids": ids, "attention_mask": att}, img).cpu()
    assert (c - a).abs().max().item() > 1e-3                            # train mode: fresh masks per call, like nn.Dropout


def test_clip_ff_training_step_in_train_mode_matches_oracle_with_the_same_masks():
    """CLIP_FF in train mode: the T5 stack's dropout (rate 0.1 at transformers' six kinds of sites) with the product's
    masks fed to the oracle; two steps (new seed each), 16 items re-forwarded in two micro-batches of 8."""
    import oracle.clip_ff_oracle as co
    from oracle.clip_ff_oracle import CLIPFeatureFusionOracle
    from test_clip_ff_host_logic import t5_mask_hook
    from uniir_b200.clip_ff import CLIPFeatureFusion
    from uniir_b200.engine import FlatAdamW, train_step
    torch.manual_seed(21)
    orc = CLIPFeatureFusionOracle("tiny-test")
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    m = CLIPFeatureFusion("tiny-test", device=dev, micro_batch=8, seed=9)
    assert m.training and m.t5_dropout_rate == 0.1
    m.train_micro_batch = 8
    m.load_state_dict(orc.state_dict(), strict=True)
    d = orc.clip_model.dims
    g = torch.Generator().manual_seed(22)
    bs = 8
    seeds = torch.Generator().manual_seed(9 * 1000003)
    opt = FlatAdamW(m._flat, lr=1e-4)
    for step in range(2):
        ids = torch.zeros(2 * bs, d.context, dtype=torch.long)
        for i in range(2 * bs):
            n = int(torch.randint(1, d.context - 2, (1,), generator=g))
            ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g); ids[i, 1 + n] = d.vocab - 1
        img = torch.randn(2 * bs, 3, d.image_size, d.image_size, generator=g)
        one = torch.ones(2 * bs, dtype=torch.long)
        b = dict(txt_batched=ids, image_batched=img, txt_mask_batched=one, image_mask_batched=one,
                 index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})
        with torch.no_grad():
            plain = orc(b)["loss"].item()
        orc.zero_grad()
        co.DROPOUT = t5_mask_hook(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=seeds)))
        try:
            ref = orc(b)
            ref["loss"].backward()
        finally:
            co.DROPOUT = None
        assert abs(ref["loss"].item() - plain) > 1e-4
        loss, acc = train_step(m, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}, opt, do_step=False)
        torch.cuda.synchronize()
        assert abs(loss.item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item())), (step, loss.item(), ref["loss"].item(), plain)
        worst = (1.0, "")
        for n_, p_o in orc.named_parameters():
            a, c = p_o.grad.flatten().double(), m._flat.g(n_).cpu().flatten().double()
            if a.numel() > 1:
                worst = min(worst, ((torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item(), n_))
        assert worst[0] > 0.99, (step, worst)
        for buf in m._flat.grad.values():
            buf.zero_()
