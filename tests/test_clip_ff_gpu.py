"""GPU parity of the CLIP_FF encode path (uniir_b200/clip_ff.py) against the oracle, whose T5 stack is
pinned against transformers' T5Stack (tests/test_oracle_clip_ff.py).

All five passed on the driver's B200 at the end of round 1 (GPUTEST_r01: 5 xpassed) and again in round 2
(gpurun_out/r2c1_tests.log), so the round-1 xfail marker is gone: a CLIP_FF regression now fails the suite."""
import pytest
import torch

from conftest import NO_DROP

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_new_kernels_against_torch():
    from uniir_b200 import lib as L
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(0)
    # RMS norm (T5LayerNorm), bf16 and fp32 outputs
    x = torch.randn(501, 768, generator=g) * 2; w = torch.randn(768, generator=g)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    y32 = ops.rmsnorm_fwd(x.to(dev), w.to(dev), 1e-6, out=torch.empty(501, 768, device=dev))
    y16 = ops.rmsnorm_fwd(x.to(dev), w.to(dev), 1e-6, out=torch.empty(501, 768, device=dev, dtype=torch.bfloat16))
    assert (y32.cpu() - ref).abs().max() < 1e-5 and ((y16.float().cpu() - ref).norm() / ref.norm()) < 5e-3
    # token mean
    t = torch.randn(5, 334, 768, generator=g)
    assert (ops.token_mean(t.to(dev)).cpu() - t.mean(1)).abs().max() < 1e-5
    # ReLU epilogue
    a = torch.randn(200, 64, generator=g).bfloat16(); b = torch.randn(256, 64, generator=g).bfloat16()
    out = torch.empty(200, 256, device=dev, dtype=torch.bfloat16)
    ops.gemm(a.to(dev), b.to(dev), act=L.ACT_RELU, out=out)
    r = torch.relu(a.float() @ b.float().t())
    assert ((out.float().cpu() - r).norm() / r.norm()) < 6e-3 and out.min().item() >= 0
    # attention with an additive score bias and no 1/sqrt(d) scaling (T5), S = 334 and a short ragged S
    for B, S, H in ((2, 334, 12), (3, 94, 2)):
        qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 0.3).bfloat16()
        bias = torch.randn(H, S, S, generator=g)
        D = H * 64
        o = ops.attn_fwd_bias(qkv.to(dev)[:, :D], qkv.to(dev)[:, D:2 * D], qkv.to(dev)[:, 2 * D:], B, S, S, H, bias.to(dev),
                              scale=1.0).float().cpu()
        q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        rr = ((q @ k.transpose(-1, -2) + bias).softmax(-1) @ v).transpose(1, 2).reshape(B * S, D)
        assert ((o - rr).norm() / rr.norm()).item() < 6e-3


def test_clip_ff_encode_matches_oracle_and_api_contract():
    from oracle.clip_ff_oracle import CLIPFeatureFusionOracle
    from uniir_b200.clip_ff import CLIPFeatureFusion
    torch.manual_seed(11)
    orc = CLIPFeatureFusionOracle("tiny-test").eval()
    with torch.no_grad():      # T5 weights at a scale where the fusion stack matters
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    m = CLIPFeatureFusion("tiny-test", device=dev, micro_batch=4).eval()      # eval: no T5 dropout, like the oracle
    missing, unexpected = m.load_state_dict(orc.state_dict(), strict=True)
    d = orc.clip_model.dims
    g = torch.Generator().manual_seed(5)
    B = 6                                  # micro-batch 4: one full chunk and a ragged tail
    ids = torch.zeros(B, d.context, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(B, 3, d.image_size, d.image_size, generator=g)
    one = torch.ones(B, dtype=torch.long)
    with torch.no_grad():
        ref = orc.encode_multimodal_input(ids, img, one, one)
    out = m.encode_multimodal_input(ids.to(dev), img.to(dev), one.to(dev), 0 * one.to(dev)).cpu()   # masks ignored
    cos = torch.nn.functional.cosine_similarity(out, ref, dim=-1)
    assert cos.min().item() > 1 - 1e-3, cos
    emb, idl = m({"txt_batched": ids, "image_batched": img, "txt_mask_batched": one, "image_mask_batched": one,
                  "did_list": list(range(10, 10 + B))}, encode_mbeir_batch=True)
    assert emb.shape == (B, d.embed_dim) and idl == list(range(10, 10 + B))


def test_blip_step_under_cuda_graphs_equals_eager():
    """BLIPEncoder.run(use_graph=True) (BLIPFeatureFusion.use_cuda_graphs, the default): eager, capture and replay must
    give the eager path's loss and gradients."""
    from oracle.blip_oracle import BLIPFeatureFusionOracle
    from uniir_b200.blip_ff import BLIPFeatureFusion
    from test_oracle_blip import _inputs
    torch.manual_seed(3)
    orc = BLIPFeatureFusionOracle("tiny-test", embed_dim=64, queue_size=16).eval()
    res = {}
    for use in (False, True):
        m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=16, micro_batch=8)
        m.train_micro_batch = 8
        m.use_cuda_graphs = use
        m.load_state_dict(orc.state_dict(), strict=False)
        losses = []
        for step in range(3):      # 2 micro-batches per step: eager + capture in step 0, replays afterwards
            ids, att, img = _inputs(orc.dims, 16, 24, 40 + step)
            b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=None,
                     image_mask_batched=None, p_did_list=[5, 6, 5, 7, 8, 9, 10, 11 + step],
                     index_mapping={"query": [[2 * i] for i in range(8)], "pos_cand": [[2 * i + 1] for i in range(8)]})
            if m._online.grad is not None:
                for gbuf in m._online.grad.values():
                    gbuf.zero_()
            out = m.compute_contrastive_loss(b, alpha=0.3)
            m.backward_contrastive()
            torch.cuda.synchronize()
            losses.append(out["loss"].item())
        res[use] = (losses, {n: m._online.g(n).clone() for n, *_ in m._online._specs})
    for a, b_ in zip(res[False][0], res[True][0]):
        assert abs(a - b_) < 1e-5 * max(1.0, abs(a))
    for n, ga in res[False][1].items():
        assert (ga - res[True][1][n]).norm() <= 2e-3 * (ga.norm() + 1e-12), n


def test_new_backward_kernels_against_torch():
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(1)
    # RMS norm backward (rstd recomputed), accumulate and overwrite
    x = torch.randn(333, 768, generator=g) * 1.5; w = torch.randn(768, generator=g); dy = torch.randn(333, 768, generator=g)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * wr).backward(dy)
    dx = torch.full((333, 768), 0.5, device=dev); dg = torch.zeros(768, device=dev)
    ops.rmsnorm_bwd(dy.to(dev), x.to(dev), w.to(dev), 1e-6, dx, dg, accumulate=True)
    assert (dx.cpu() - 0.5 - xr.grad).abs().max() < 1e-4 and (dg.cpu() - wr.grad).abs().max() < 1e-3 * wr.grad.abs().max()
    ops.rmsnorm_bwd(dy.to(dev), x.to(dev), w.to(dev), 1e-6, dx, dg, accumulate=False)
    assert (dx.cpu() - xr.grad).abs().max() < 1e-4
    # token mean backward
    d = torch.randn(5, 768, generator=g)
    o = ops.token_mean_bwd(d.to(dev), 334, torch.empty(5, 334, 768, device=dev))
    assert (o.cpu() - (d / 334)[:, None, :].expand(5, 334, 768)).abs().max() < 1e-7
    # attention backward with a score bias (T5: scale 1), incl. the bias gradient summed over the batch
    B, S, H = 3, 94, 2
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 0.3).bfloat16()
    bias = torch.randn(H, S, S, generator=g)
    dout = torch.randn(B * S, D, generator=g).bfloat16()
    qd = qkv.to(dev)
    lse = torch.empty(B, H, S, device=dev)
    out = ops.attn_fwd_bias(qd[:, :D], qd[:, D:2 * D], qd[:, 2 * D:], B, S, S, H, bias.to(dev), scale=1.0, lse=lse)
    dqkv = torch.zeros(B * S, 3 * D, device=dev, dtype=torch.bfloat16); dbias = torch.zeros(H, S, S, device=dev)
    ops.attn_bwd_bias(qd[:, :D], qd[:, D:2 * D], qd[:, 2 * D:], out, dout.to(dev), lse, B, S, S, H, bias.to(dev), dbias,
                      dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], scale=1.0)
    xq = qkv.float().requires_grad_(True); br = bias.clone().requires_grad_(True)
    q, k, v = xq.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ((q @ k.transpose(-1, -2) + br).softmax(-1) @ v).transpose(1, 2).reshape(B * S, D).backward(dout.float())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(dqkv, xq.grad) < 1.2e-2 and rel(dbias, br.grad) < 1.2e-2


def test_clip_ff_training_step_matches_oracle():
    from oracle.clip_ff_oracle import CLIPFeatureFusionOracle
    from uniir_b200.clip_ff import CLIPFeatureFusion
    from uniir_b200.engine import FlatAdamW, train_step
    torch.manual_seed(21)
    orc = CLIPFeatureFusionOracle("tiny-test")
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    m = CLIPFeatureFusion("tiny-test", device=dev, micro_batch=8)
    m.train_micro_batch = 8
    m.set_dropout(0.0)           # the deterministic arithmetic; train mode with the T5 dropout: tests/test_dropout_gpu.py
    m.load_state_dict(orc.state_dict(), strict=True)
    d = orc.clip_model.dims
    g = torch.Generator().manual_seed(22)
    bs = 8
    ids = torch.zeros(2 * bs, d.context, dtype=torch.long)
    for i in range(2 * bs):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(2 * bs, 3, d.image_size, d.image_size, generator=g)
    one = torch.ones(2 * bs, dtype=torch.long)
    b = dict(txt_batched=ids, image_batched=img, txt_mask_batched=one, image_mask_batched=one,
             index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})
    ref = orc(b)
    ref["loss"].backward()
    opt = FlatAdamW(m._flat, lr=1e-4)
    loss, acc = train_step(m, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}, opt, do_step=False)
    torch.cuda.synchronize()
    assert abs(loss.item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item()))
    worst = (1.0, "")
    for n_, p_o in orc.named_parameters():
        a, c = p_o.grad.flatten().double(), m._flat.g(n_).cpu().flatten().double()
        if a.numel() > 1:
            worst = min(worst, ((torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item(), n_))
    assert worst[0] > 0.995, worst
