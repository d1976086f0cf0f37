"""Drop-in for the reference ``CLIPScoreFusion`` (src/models/uniir_clip/clip_scorefusion/
clip_sf.py:13-168) running on the B200-native kernels.

Same constructor signature, same ``forward(batch, encode_mbeir_batch=False)`` /
``encode_mbeir_batch`` / ``get_tokenizer`` / ``get_img_preprocess_fn`` surface, same
``clip_model.*`` state-dict keys, same return values and error behaviour.  Differences, all
deliberate:
  * compute is bf16-in / fp32-accumulate on hand-written sm_100a kernels, no autocast;
  * the in-batch loss uses embedding-level gradient caching: the towers run in micro-batches
    (forward without saving -> loss and d(embedding) on the full batch -> re-forward with
    saved activations + backward per micro-batch), which is exactly the big-batch loss and
    lets global batch 4096 train on one GPU;
  * weights are randomly initialised (upstream scales) unless a checkpoint is loaded -- there
    is no network to fetch the pretrained CLIP weights that ``clip.load`` downloads.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from . import dist_utils as du
from . import lib as L
from . import ops
from .clip_towers import CLIPEncoder
from .params import CLIP_CONFIGS, FlatParamModule, clip_param_specs


def _dist_on():
    return dist.is_available() and dist.is_initialized()


class _ContrastiveLossFn(torch.autograd.Function):
    """Bridges the hand-written backward into autograd so the reference engine's
    ``scaler.scale(loss).backward()`` (uniir_clip/engine.py:34) keeps working."""

    @staticmethod
    def forward(ctx, model, batch, *params):
        loss, acc = model._loss_forward(batch)
        ctx.model = model
        ctx.batch = batch
        loss, acc = loss.clone(), acc.clone()
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, _g_acc):
        model = ctx.model
        model._loss_backward(ctx.batch, g_loss)
        grads = tuple(model.clip_model.g(name).clone() for name, *_ in model.clip_model._specs)
        for buf in model.clip_model.grad.values():
            buf.zero_()
        return (None, None) + grads


class CLIPScoreFusion(nn.Module):
    def __init__(self, model_name="ViT-B/32", device="cuda", jit=False, download_root=None, config=None,
                 micro_batch=None, seed=2023):
        super().__init__()
        if model_name not in CLIP_CONFIGS:
            raise RuntimeError(f"Model {model_name} not found; available models = {list(CLIP_CONFIGS)}")
        L.load()  # fail loudly if the CUDA extension is missing: there is no fallback path
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.UniirError("uniir_b200.CLIPScoreFusion runs on CUDA only (no CPU fallback); "
                               "use oracle/ for CPU reference numbers")
        self.model_name = model_name
        self.dims = CLIP_CONFIGS[model_name]
        self.device_ = dev
        self.clip_model = FlatParamModule(clip_param_specs(self.dims), dev, seed=seed)
        self.loss_function = nn.CrossEntropyLoss()  # kept for attribute parity (clip_sf.py:27)
        if config is not None:  # clip_sf.py:28-30
            self.gather_embeddings = config.model.gather_embeddings
            self.in_batch_neg_num = config.data_config.in_batch_neg_num
        self.micro_batch = micro_batch
        self.use_cuda_graphs = True
        self.skip_masked_towers = True   # rows with mask 0 never enter that tower (same result)
        # Gradient caching runs every micro-batch forward twice (embeddings first, activations again before its
        # backward).  As many micro-batches as HBM holds keep their activations from the first pass instead
        # (one extra encoder = one extra set of activation buffers each): "auto" sizes that from free memory,
        # an int forces it, 0 disables.  Same arithmetic, one forward less for those micro-batches.
        self.retain_micro_batches = "auto"
        self.retain_reserve_bytes = 16 << 30   # left free for the loss, NCCL and allocator slack under "auto"
        self._retained = []
        self._upload = None          # (chunk, events) of an image upload in flight (_prep / _wait_upload)
        self._h2d_stream = None
        self._enc = {}
        self._nce = None
        self._train_state = None

    # ---- reference surface ---------------------------------------------------------
    def get_img_preprocess_fn(self):
        from .preprocess import clip_image_preprocess
        return clip_image_preprocess(self.dims.image_size)

    def get_tokenizer(self):
        from .preprocess import clip_tokenizer
        return clip_tokenizer(context_length=77, vocab=self.dims.vocab)

    def get_logit_scale(self):
        return self.clip_model.logit_scale.exp()

    def encode_text(self, text_tensor):
        img = torch.zeros(text_tensor.shape[0], 3, self.dims.image_size, self.dims.image_size,
                          device=self.device_)
        one = torch.ones(text_tensor.shape[0], dtype=torch.long, device=self.device_)
        return self.encode_multimodal_input(text_tensor, img, one, 0 * one)

    def encode_image(self, image_tensor):
        ids = torch.zeros(image_tensor.shape[0], 77, dtype=torch.int32, device=self.device_)
        one = torch.ones(image_tensor.shape[0], dtype=torch.long, device=self.device_)
        return self.encode_multimodal_input(ids, image_tensor, 0 * one, one)

    def fuse_embeddings(self, img_emb, txt_emb):
        return img_emb + txt_emb

    @torch.no_grad()
    def encode_multimodal_input(self, txt_tensor, img_tensor, txt_mask, img_mask):
        """[B,77] int, [B,3,H,W] fp32, [B] int64, [B] int64 -> fused fp32 [B, E] (clip_sf.py:53-63)."""
        emb = torch.empty(txt_tensor.shape[0], self.dims.embed_dim, device=self.device_, dtype=torch.float32)
        self._embed_into(emb, txt_tensor, img_tensor, txt_mask, img_mask)
        return emb

    def forward(self, batch, encode_mbeir_batch=False):
        if encode_mbeir_batch:
            return self.encode_mbeir_batch(batch)
        return self.compute_inbatch_contrastive_loss(batch)

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        assert id_list is not None, "id_list must be provided."
        assert isinstance(id_list[0], int), "id_list must be hashed to int."
        embeddings = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"],
                                                  batch["txt_mask_batched"], batch["image_mask_batched"])
        assert embeddings.size(0) == len(id_list), "embeddings and id_batched must have the same batch size."
        return embeddings, id_list

    def compute_inbatch_contrastive_loss(self, batch):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = [self.clip_model.get_parameter(n) for n, *_ in self.clip_model._specs]
            loss, acc = _ContrastiveLossFn.apply(self, batch, *params)
        else:
            loss, acc = self._loss_forward(batch)
            loss, acc = loss.clone(), acc.clone()
        return {"loss": loss, "accuracy": acc}

    # ---- internals -----------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.clip_model.rebind_after_apply()
        return out

    def _encoder(self, mb, train):
        key = (mb, train)
        if key not in self._enc:
            # one live encoder per mode keeps the activation memory bounded
            for k in [k for k in self._enc if k[1] == train]:
                del self._enc[k]
            self._enc[key] = CLIPEncoder(self.clip_model, self.dims, mb, self.device_, train)
        return self._enc[key]

    def _pick_mb(self, n, train):
        if self.micro_batch:
            return min(self.micro_batch, n)
        cap = 64 if train else 256
        return min(cap, n)

    def _prep(self, txt, img, chunk=None):
        """Inputs on the device.  A pinned host image batch (the end-to-end path: 602 KB per item, 4.9 GB for the
        4096-pair batch) is uploaded on a side stream in micro-batch chunks, each with an event, so that only the
        first chunk's copy is exposed and the rest overlaps the towers (`_wait_upload` before a chunk is read)."""
        txt = txt.to(self.device_, torch.int32, non_blocking=True).contiguous()
        if (chunk and img.device.type == "cpu" and img.is_pinned() and img.dtype == torch.float32 and img.is_contiguous()
                and img.shape[0] > chunk):
            main = torch.cuda.current_stream(self.device_)
            if self._h2d_stream is None:
                self._h2d_stream = torch.cuda.Stream(device=self.device_)
            side = self._h2d_stream
            dev = torch.empty(img.shape, device=self.device_, dtype=torch.float32)
            side.wait_stream(main)           # the allocator may hand out memory the main stream is still using
            dev.record_stream(side)
            events = []
            with torch.cuda.stream(side):
                for s in range(0, img.shape[0], chunk):
                    dev[s:s + chunk].copy_(img[s:s + chunk], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events.append(ev)
            self._upload = (chunk, events)
            return txt, dev
        img = img.to(self.device_, torch.float32, non_blocking=True).contiguous()
        return txt, img

    def _wait_upload(self, s=None, e=None):
        """Make the current stream wait for the uploaded image rows [s, e) (all rows if None)."""
        up = self._upload
        if up is None:
            return
        chunk, events = up
        lo, hi = (0, len(events)) if s is None else (s // chunk, (e + chunk - 1) // chunk)
        cur = torch.cuda.current_stream(self.device_)
        for ev in events[lo:hi]:
            cur.wait_event(ev)

    def _sync_weights(self, enc):
        self.clip_model.refresh_shadow()
        enc.refresh_weights()

    def _mask_plan(self, mt, mv):
        """None when every item has both modalities (the fused two-tower pass runs); otherwise the batch
        rows each tower has to encode, as int32 device tensors.  Costs one small device->host read."""
        if not self.skip_masked_towers:
            return None
        nz = torch.stack([mt != 0, mv != 0])
        if bool(nz.all()):
            return None
        return (nz[0].nonzero().flatten().to(torch.int32).contiguous(),
                nz[1].nonzero().flatten().to(torch.int32).contiguous())

    def _retained_encoders(self, enc, n_mb):
        """Extra training encoders whose activations survive from the embedding pass to the backward pass."""
        want = self.retain_micro_batches
        if self._retained and (self._retained[0].mb != enc.mb or self._retained[0].P is not enc.P):
            self._retained = []
        if want == "auto":
            if not getattr(self, "_retain_sized", None) == enc.mb:
                # cost of one encoder = what the main one took; keep `retain_reserve_bytes` free
                torch.cuda.synchronize()
                per = max(enc.bytes_allocated, 1)
                free, _ = torch.cuda.mem_get_info(self.device_)
                free += torch.cuda.memory_reserved(self.device_) - torch.cuda.memory_allocated(self.device_)
                self._retain_auto = max(0, int((free - self.retain_reserve_bytes) // per))
                self._retain_sized = enc.mb
            want = self._retain_auto
        want = max(0, min(int(want), n_mb))
        while len(self._retained) < want:
            try:
                r = CLIPEncoder(self.clip_model, self.dims, enc.mb, self.device_, True)
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                if self.retain_micro_batches == "auto":
                    self._retain_auto = len(self._retained)
                break
            r.conv_w16 = enc.conv_w16        # one bf16 copy of the patch-embedding weight (refresh_weights)
            self._retained.append(r)
        return self._retained[:want]

    def _embed_into(self, emb, txt, img, mt, mv, enc=None, retain=False):
        """Forward in micro-batches; the tail chunk is padded.  Returns the mask plan.  retain=True (training,
        every item has both modalities): the first micro-batches run on the retained encoders and keep their
        activations for _loss_backward."""
        n = txt.shape[0]
        mb = self._pick_mb(n, False) if enc is None else enc.mb
        txt, img = self._prep(txt, img, chunk=mb)     # no-op for device tensors (training: uploaded by _loss_forward)
        mt = mt.to(self.device_, torch.int64).contiguous()
        mv = mv.to(self.device_, torch.int64).contiguous()
        enc = enc or self._encoder(mb, False)
        self._sync_weights(enc)
        plan = self._mask_plan(mt, mv)
        self._n_retained = 0
        if plan is None:
            kept = self._retained_encoders(enc, n // mb) if (retain and n % mb == 0) else []
            self._n_retained = len(kept)
            for i, s in enumerate(range(0, n, mb)):
                e = min(n, s + mb)  # a ragged tail is zero-padded inside set_inputs; padded rows are dropped
                self._wait_upload(s, e)
                if i < len(kept):
                    kept[i].set_inputs(txt[s:e], img[s:e], mt[s:e], mv[s:e])
                    emb[s:e].copy_(kept[i].fwd_save(self.use_cuda_graphs))
                    continue
                enc.set_inputs(txt[s:e], img[s:e], mt[s:e], mv[s:e])
                emb[s:e].copy_(enc.embed(self.use_cuda_graphs)[:e - s])
            self._upload = None
            return None
        # each tower encodes only its unmasked rows; out = mt * text + mv * image is scattered back
        self._wait_upload()
        self._upload = None
        emb.zero_()
        for tower, rows, src, mask in (("text", plan[0], txt, mt), ("vision", plan[1], img, mv)):
            for s in range(0, rows.numel(), mb):
                r = rows[s:s + mb]
                enc.set_tower_input(tower, src, r)
                ops.score_fuse_scatter(enc.run_tower(tower, False, self.use_cuda_graphs), r, mask, emb)
        return plan

    def _world(self):
        if getattr(self, "gather_embeddings", False):
            if not _dist_on():
                raise RuntimeError("gather_embeddings=True needs an initialised process group "
                                   "(reference clip_sf.py:103 calls torch.distributed.nn.all_gather)")
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _loss_forward(self, batch):
        """Phase A (embeddings, no activations) + loss fwd/bwd on the embeddings."""
        im = batch["index_mapping"]
        dev = self.device_
        q_idx = torch.tensor(im["query"], dtype=torch.int32).flatten().to(dev)
        p_idx = torch.tensor(im["pos_cand"], dtype=torch.int32).flatten().to(dev)
        bs = q_idx.numel()
        n_items = batch["txt_batched"].shape[0]
        E = self.dims.embed_dim
        rank, world = self._world()
        mb = self._pick_mb(n_items, True)
        if n_items % mb != 0 or mb % 8 != 0:
            # training micro-batches must tile the batch exactly (token-contraction GEMMs)
            mb = next((m for m in range(mb - mb % 8, 7, -8) if n_items % m == 0), None)
            if mb is None:
                raise ValueError(f"encoder batch of {n_items} items cannot be tiled by a micro-batch "
                                 "that is a multiple of 8")
        enc = self._encoder(mb, True)
        self._train_mb = mb
        st = dict(q_idx=q_idx, p_idx=p_idx, bs=bs, n_items=n_items, rank=rank, world=world, enc=enc)
        st["txt"], st["img"] = self._prep(batch["txt_batched"], batch["image_batched"], chunk=mb)
        st["mt"] = batch["txt_mask_batched"].to(dev, torch.int64).contiguous()
        st["mv"] = batch["image_mask_batched"].to(dev, torch.int64).contiguous()
        emb = torch.empty(n_items, E, device=dev, dtype=torch.float32)
        st["plan"] = self._embed_into(emb, st["txt"], st["img"], st["mt"], st["mv"], enc=enc, retain=True)
        st["n_retained"] = self._n_retained
        if "neg_cand_list" in im:   # hard-negative branch (clip_sf.py:105-131): no gather, fused fwd+bwd
            n_idx = torch.tensor(im["neg_cand_list"], dtype=torch.int32).to(dev).contiguous()
            P = self.clip_model
            P.ensure_grad()
            st["hard_dls"] = torch.zeros((), device=dev, dtype=torch.float32)
            loss, acc, st["hard_d_emb"] = ops.hardneg_loss(emb, q_idx, p_idx, n_idx, int(getattr(self, "in_batch_neg_num", 0)),
                                                           P.w("logit_scale"), st["hard_dls"])
            st["emb"] = emb
            self._train_state = st
            return loss, acc
        nce = self._nce
        if nce is None or (nce.bs, nce.n, nce.E) != (bs, bs * world, E):
            nce = self._nce = ops.InfoNCE(bs, bs * world, E, dev)
        off = rank * bs
        pn = nce.normalize(emb, q_idx, p_idx, off)
        if world > 1:
            du.all_gather_rows(pn.clone(), out=nce.all_p)
        loss, acc = nce.forward(self.clip_model.w("logit_scale"), off)
        st["emb"] = emb
        self._train_state = st
        return loss, acc

    def _loss_backward(self, batch, g_loss=None):
        """Loss backward on embeddings, then Phase C: re-forward with saved activations and
        backprop each micro-batch.  Gradients accumulate into clip_model.grad (flat fp32)."""
        st = self._train_state
        assert st is not None, "backward without forward"
        P = self.clip_model
        P.ensure_grad()
        nce, enc = self._nce, st["enc"]
        bs, rank, world, E = st["bs"], st["rank"], st["world"], self.dims.embed_dim
        off = rank * bs
        go = None if g_loss is None else g_loss.to(torch.float32).reshape(()).contiguous()
        if "hard_d_emb" in st:
            d_emb = st["hard_d_emb"] if go is None else st["hard_d_emb"] * go
            P.g("logit_scale").add_(st["hard_dls"] if go is None else st["hard_dls"] * go)
        else:
            d_qn, d_all_p = nce.backward(P.w("logit_scale"), off, P.g("logit_scale"), grad_out=go)
            d_pn = du.reduce_scatter_rows(d_all_p) if world > 1 else d_all_p
            d_emb = torch.zeros(st["n_items"], E, device=self.device_, dtype=torch.float32)
            nce.normalize_bwd(st["q_idx"], st["p_idx"], d_pn, d_emb, off)
        mb = enc.mb
        if st["plan"] is None:
            for i, s in enumerate(range(0, st["n_items"], mb)):
                e = s + mb
                if i < st["n_retained"]:     # activations kept since the embedding pass: backward only
                    self._retained[i].bwd_only(d_emb[s:e], self.use_cuda_graphs)
                    continue
                enc.set_inputs(st["txt"][s:e], st["img"][s:e], st["mt"][s:e], st["mv"][s:e], d_emb[s:e])
                enc.fwdbwd(self.use_cuda_graphs)
        else:
            # masked towers skipped: per tower, re-forward + backward over its compacted rows; the rows that
            # pad the last micro-batch carry a zero output gradient, so they add nothing to any parameter
            d_emb = d_emb.contiguous()
            for tower, rows, src, mask, d_in in (("text", st["plan"][0], st["txt"], st["mt"], enc.d_t),
                                                 ("vision", st["plan"][1], st["img"], st["mv"], enc.d_v)):
                for s in range(0, rows.numel(), mb):
                    r = rows[s:s + mb]
                    enc.set_tower_input(tower, src, r)
                    ops.score_fuse_gather_bwd(d_emb, r, mask, out=d_in)
                    enc.run_tower(tower, True, self.use_cuda_graphs)
        self._train_state = None
