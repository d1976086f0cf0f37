"""Generate the committed golden vectors under tests/golden/.

    python scripts/make_golden.py

The reference holds no golden vectors of its own.  Where /root/reference exists (the build
container) the CLIP_SF vectors are produced by the REFERENCE'S OWN CLASS -- the unmodified
uniir_clip/clip_scorefusion/clip_sf.py imported through tests/ref_clip_shim.py -- and the oracle is asserted to agree
to 1e-6 before anything is written (`source` = "reference" in the file); elsewhere they come from the oracle
(`source` = "oracle").  They travel to the GPU box, where /root/reference does not exist.
  clip_tiny.npz   : tiny-test CLIP_SF weights + synthetic batch -> embeddings, loss, accuracy,
                    gradients of 4 representative parameters
  search_small.npz: fp16 queries/candidates/ids -> FAISS-semantics top-k (incl. N < k padding)
  infonce.npz     : embeddings + index mapping -> loss/accuracy/d_emb for world=1 and a
                    simulated rank 1 of world 2
"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.clip_oracle import CLIPScoreFusionOracle, contrastive_loss_from_embeddings
from oracle.search_oracle import flat_ip_search
from uniir_b200.preprocess import synthetic_clip_tokens

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_clip_shim as shim  # noqa: E402

SOURCE = "reference" if shim.available() else "oracle"


def _reference_sf(orc, in_batch_neg_num=0):
    """The reference CLIPScoreFusion carrying the oracle's weights (None off the build container)."""
    if not shim.available():
        return None
    ref = shim.load_clip_sf().CLIPScoreFusion(model_name="tiny-test", device="cpu", config=shim.Cfg(False, in_batch_neg_num))
    ref.load_state_dict(orc.state_dict(), strict=True)
    return ref


def _reference_loss(emb, im, scale, in_batch_neg_num=0):
    """clip_sf.py:68-147 executed verbatim on given fused embeddings (encode_multimodal_input patched out)."""
    import math
    ref = shim.load_clip_sf().CLIPScoreFusion(model_name="tiny-test", device="cpu", config=shim.Cfg(False, in_batch_neg_num))
    with torch.no_grad():
        ref.clip_model.logit_scale.fill_(math.log(100.0))
    assert ref.get_logit_scale().item() == float(scale)
    ref.encode_multimodal_input = lambda *a: emb
    dummy = dict(txt_batched=None, image_batched=None, txt_mask_batched=None, image_mask_batched=None, index_mapping=im)
    return ref(dummy)


def clip_tiny():
    torch.manual_seed(1234)
    m = CLIPScoreFusionOracle("tiny-test")
    d = m.clip_model.dims
    g = torch.Generator().manual_seed(7)
    bs = 8
    n = 2 * bs
    ids = synthetic_clip_tokens(n, g, d.context, d.vocab)
    img = torch.randn(n, 3, d.image_size, d.image_size, generator=g)
    mt = torch.ones(n, dtype=torch.long); mv = torch.ones(n, dtype=torch.long)
    mt[3] = 0; mv[6] = 0; mt[9] = 0
    batch = dict(txt_batched=ids, image_batched=img, txt_mask_batched=mt, image_mask_batched=mv,
                 index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})
    batch["txt_batched"] = ids.long()
    ref = _reference_sf(m)
    src = ref if ref is not None else m          # the vectors come from the reference class where it exists
    out = src(batch)
    out["loss"].backward()
    with torch.no_grad():
        emb = src.encode_multimodal_input(ids.long(), img, mt, mv)
    if ref is not None:                          # ... and the oracle must agree before anything is written
        o2 = m(batch); o2["loss"].backward()
        assert abs(o2["loss"].item() - out["loss"].item()) < 1e-6 and o2["accuracy"].item() == out["accuracy"].item()
        with torch.no_grad():
            assert torch.allclose(m.encode_multimodal_input(ids.long(), img, mt, mv), emb, atol=1e-6)
        for k, p in m.named_parameters():
            assert torch.allclose(p.grad, ref.get_parameter(k).grad, atol=1e-6, rtol=1e-5), k
    m = src
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    grads = {k: m.get_parameter(k).grad.numpy() for k in
             ["clip_model.visual.proj", "clip_model.text_projection", "clip_model.logit_scale",
              "clip_model.visual.transformer.resblocks.0.mlp.c_fc.weight",
              "clip_model.transformer.resblocks.1.attn.in_proj_weight", "clip_model.visual.conv1.weight"]}
    np.savez_compressed(os.path.join(OUT, "clip_tiny.npz"), ids=ids.numpy(), img=img.numpy().astype(np.float32),
                        mt=mt.numpy(), mv=mv.numpy(), emb=emb.numpy(), loss=out["loss"].item(), acc=out["accuracy"].item(),
                        source=SOURCE,
                        **{"w:" + k: v for k, v in sd.items()}, **{"g:" + k: v for k, v in grads.items()})


def search_small():
    rng = np.random.default_rng(11)
    d = 64
    c = rng.standard_normal((2500, d)).astype(np.float16)
    c[17] = 0  # zero row: normalize_L2 leaves it, score 0
    q = rng.standard_normal((37, d)).astype(np.float16)
    q[5] = 0
    ids = (rng.permutation(2500).astype(np.int64) * 10_000_000 // 7) + 3  # hash_did-like sparse ids
    D, I = flat_ip_search(q, c, ids, 10, dtype=torch.float64)
    D2, I2 = flat_ip_search(q, c[:6], ids[:6], 10, dtype=torch.float64)  # N < k -> -1 padding
    np.savez_compressed(os.path.join(OUT, "search_small.npz"), q=q, c=c, ids=ids, D=D, I=I, D_small=D2, I_small=I2)


def infonce():
    g = torch.Generator().manual_seed(3)
    bs, E = 6, 32
    emb = torch.randn(2 * bs, E, generator=g, requires_grad=True)
    im = {"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]}
    scale = torch.tensor(np.log(100.0), dtype=torch.float32).exp()   # what logit_scale.exp() gives (clip_sf.py:65-66)
    o1 = contrastive_loss_from_embeddings(emb, im, scale)
    (g1,) = torch.autograd.grad(o1["loss"], emb)
    other = torch.nn.functional.normalize(torch.randn(bs, E, generator=g), dim=-1)  # rank 0's p rows
    o2 = contrastive_loss_from_embeddings(emb, im, scale, all_p_fn=lambda p: torch.cat([other, p], 0), rank=1)
    (g2,) = torch.autograd.grad(o2["loss"], emb)
    # hard-negative branch (clip_sf.py:105-131): [q, p, n0, n1] per pair, in_batch_neg_num 2
    hbs, neg, inb = 5, 2, 2
    hemb = torch.randn(hbs * (2 + neg), E, generator=g, requires_grad=True)
    him = {"query": [[4 * i] for i in range(hbs)], "pos_cand": [[4 * i + 1] for i in range(hbs)],
           "neg_cand_list": [[4 * i + 2, 4 * i + 3] for i in range(hbs)]}
    oh = contrastive_loss_from_embeddings(hemb, him, scale, in_batch_neg_num=inb)
    (gh,) = torch.autograd.grad(oh["loss"], hemb)
    if shim.available():        # the reference's own loss code on the same embeddings must give the same numbers
        for e_, im_, o_, g_, inb_ in ((emb, im, o1, g1, 0), (hemb, him, oh, gh, inb)):
            e2 = e_.detach().clone().requires_grad_(True)
            r = _reference_loss(e2, im_, scale, inb_)
            (gr,) = torch.autograd.grad(r["loss"], e2)
            assert abs(r["loss"].item() - o_["loss"].item()) < 1e-6 and r["accuracy"].item() == o_["accuracy"].item()
            assert torch.allclose(gr, g_, atol=1e-7, rtol=1e-5)
    np.savez_compressed(os.path.join(OUT, "infonce.npz"), emb=emb.detach().numpy(), other=other.numpy(),
                        loss1=o1["loss"].item(), acc1=o1["accuracy"].item(), g1=g1.numpy(),
                        loss2=o2["loss"].item(), acc2=o2["accuracy"].item(), g2=g2.numpy(),
                        hn_emb=hemb.detach().numpy(), hn_loss=oh["loss"].item(), hn_acc=oh["accuracy"].item(),
                        hn_g=gh.numpy(), hn_in_batch_neg_num=inb, source=SOURCE)


def blip_tiny():
    from oracle.blip_oracle import BLIP_CONFIGS, BLIPFeatureFusionOracle
    from test_oracle_blip import _inputs
    torch.manual_seed(4321)
    m = BLIPFeatureFusionOracle("tiny-test", queue_size=64).eval()
    d = m.dims
    ids, att, img = _inputs(d, 6, 24, 5)
    with torch.no_grad():
        emb = m.encode_multimodal_input(ids, att, img)
    np.savez_compressed(os.path.join(OUT, "blip_tiny.npz"), ids=ids.numpy(), att=att.numpy(), img=img.numpy(), emb=emb.numpy(),
                        **{"w:" + k: v.numpy() for k, v in m.state_dict().items() if k.startswith(("visual_encoder.", "text_encoder."))})


def blip_sf_tiny():
    """BLIP_SF (blip_sf.py) masked score-fusion embeddings of the oracle, which tests/test_oracle_blip.py pins
    against the real reference class."""
    from oracle.blip_oracle import BLIPScoreFusionOracle
    from test_oracle_blip import _inputs
    torch.manual_seed(8765)
    m = BLIPScoreFusionOracle("tiny-test", embed_dim=64, queue_size=64).eval()
    ids, att, img = _inputs(m.dims, 6, 24, 6)
    mt = torch.tensor([1, 1, 0, 1, 0, 1]); mv = torch.tensor([1, 0, 1, 1, 1, 0])
    with torch.no_grad():
        emb = m.encode_multimodal_input(ids, att, img, mt, mv)
    keep = ("visual_encoder.", "text_encoder.", "vision_proj.", "text_proj.")
    np.savez_compressed(os.path.join(OUT, "blip_sf_tiny.npz"), ids=ids.numpy(), att=att.numpy(), img=img.numpy(), mt=mt.numpy(),
                        mv=mv.numpy(), emb=emb.numpy(),
                        **{"w:" + k: v.numpy() for k, v in m.state_dict().items() if k.startswith(keep)})


if __name__ == "__main__":
    if "--only-blip-sf" in sys.argv:
        blip_sf_tiny()
    else:
        clip_tiny(); search_small(); infonce(); blip_tiny(); blip_sf_tiny()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
