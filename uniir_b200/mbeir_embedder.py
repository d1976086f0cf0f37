"""Embedding extraction -- drop-in for reference src/common/mbeir_embedder.py on the B200 path.

Keeps the reference contracts:
  * ``generate_embeds_and_ids_for_dataset_with_gather(model, data_loader, device, use_fp16)`` ->
    ``(np.float16 [N,d], list[int])`` on rank 0 (mbeir_embedder.py:32-120), with each rank embedding
    the contiguous row block of ``ContiguousDistributedSampler`` (dist_utils.py:100-109);
  * files ``{uniir_dir}/{embed_dir_name}/{path_suffix}/{split}/mbeir_{name}_{split}_{embed,ids}.npy``
    (mbeir_embedder.py:346-369);
  * CLI ``--uniir_dir --mbeir_data_dir --config_path`` under ``torch.distributed.run``.
Differences: no DDP wrapper around an inference model, ids gathered as an int64 tensor instead of
``gather_object`` of python lists, uniqueness checked with ``torch.unique`` on the device; and
``generate_local_embeds_and_ids`` + ``save_rank_shards`` / ``load_rank_shard`` hand each rank's rows to the
retrieval stage directly (no rank-0 funnel of the 8.6 GB matrix).
The M-BEIR dataset / collator classes are the reference's own (``data.mbeir_dataset``): the loader is
outside the hot path, so ``main`` imports them from the reference tree on PYTHONPATH.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torch.distributed as dist

from . import dist_utils as du


@torch.no_grad()
def generate_local_embeds_and_ids(model, data_loader, device):
    """This rank's rows only: (fp16 [n,d], int64 [n]) DEVICE tensors, no collective (extraction
    shards by contiguous row block and needs no exchange)."""
    chunks, id_list = [], []
    for batch in data_loader:
        for key, value in batch.items():
            if isinstance(value, torch.Tensor):
                batch[key] = value.to(device, non_blocking=True)
        emb, ids = model(batch, encode_mbeir_batch=True)
        chunks.append(emb.half())
        id_list.extend(ids)
    d = chunks[0].shape[1] if chunks else getattr(getattr(model, "dims", None), "embed_dim", 0)
    emb = torch.cat(chunks, 0) if chunks else torch.empty(0, d, device=device, dtype=torch.float16)
    return emb, torch.tensor(id_list, dtype=torch.int64, device=device)


def _gather_ragged(x, fill):
    """All-gather of per-rank row blocks with different row counts -> list of W tensors (device)."""
    W, r = du.world(), du.rank()
    sizes = torch.zeros(W, dtype=torch.int64, device=x.device)
    sizes[r] = x.shape[0]
    dist.all_reduce(sizes)
    n_max = int(sizes.max())
    pad = torch.full((n_max,) + tuple(x.shape[1:]), fill, device=x.device, dtype=x.dtype)
    pad[:x.shape[0]] = x
    allx = du.all_gather_rows(pad).view(W, n_max, *x.shape[1:])
    return [allx[w, :int(sizes[w])] for w in range(W)]


@torch.no_grad()
def generate_embeds_and_ids_for_dataset_with_gather(model, data_loader, device, use_fp16=True):
    """``use_fp16`` is accepted for signature parity; the B200 towers always run bf16-in/fp32-accumulate
    and the stored embeddings are fp16 exactly like the reference (mbeir_embedder.py:56,110)."""
    emb, ids_t = generate_local_embeds_and_ids(model, data_loader, device)
    id_list = ids_t.tolist()
    if du.is_on() and du.world() > 1:
        all_e = _gather_ragged(emb, 0)
        all_i = _gather_ragged(ids_t, -1)
        if du.rank() != 0:
            return None, id_list
        emb, ids_t = torch.cat(all_e, 0), torch.cat(all_i, 0)
    assert ids_t.numel() == emb.shape[0]
    assert torch.unique(ids_t).numel() == ids_t.numel(), "Hashed IDs should be unique"
    return emb.cpu().numpy(), ids_t.cpu().tolist()


# ---- per-rank shard hand-off: no rank-0 funnel of 8.6 GB, no FAISS file -------------------
def shard_paths(embed_path, ids_path, r, world):
    tag = f".rank{r:02d}-of-{world:02d}.npy"
    return embed_path[:-4] + tag, ids_path[:-4] + tag


@torch.no_grad()
def save_rank_shards(embed_path, ids_path, emb, ids):
    """Every rank writes ITS rows next to where the reference writes the gathered file.  Only the int64 ids are
    exchanged (45 MB for 5.6 M rows) to keep the reference's global uniqueness assertion
    (mbeir_embedder.py:103-108).  Returns this rank's (embed, ids) shard paths."""
    r, W = (du.rank(), du.world()) if du.is_on() else (0, 1)
    all_ids = torch.cat(_gather_ragged(ids, -1)) if W > 1 else ids
    assert torch.unique(all_ids).numel() == all_ids.numel(), "Hashed IDs should be unique"
    e_path, i_path = shard_paths(embed_path, ids_path, r, W)
    os.makedirs(os.path.dirname(e_path), exist_ok=True)
    np.save(e_path, emb.cpu().numpy().astype(np.float16, copy=False))
    np.save(i_path, ids.cpu().numpy().astype(np.int64, copy=False))
    return e_path, i_path


def load_rank_shard(embed_path, ids_path, r, world):
    """Rows of rank r out of `world`: its own shard file when the extraction ran with the same world size,
    otherwise the reference's contiguous row block (dist_utils.py:100-109) cut from whatever shard files (or
    the single gathered file) exist -- read with mmap, so only these rows are touched."""
    e_path, i_path = shard_paths(embed_path, ids_path, r, world)
    if os.path.exists(e_path):
        return np.load(e_path), np.load(i_path)
    base = os.path.basename(embed_path)[:-4]
    folder = os.path.dirname(embed_path)
    import re
    found = sorted(f for f in os.listdir(folder) if re.fullmatch(re.escape(base) + r"\.rank\d+-of-\d+\.npy", f))
    if found:
        w_old = int(found[0].rsplit("-of-", 1)[1][:-4])
        assert len(found) == w_old, f"expected {w_old} shard files for {base}, found {len(found)}"
        embs = [np.load(os.path.join(folder, f), mmap_mode="r") for f in found]
        idss = [np.load(shard_paths(embed_path, ids_path, k, w_old)[1], mmap_mode="r") for k in range(w_old)]
    else:
        embs, idss = [np.load(embed_path, mmap_mode="r")], [np.load(ids_path, mmap_mode="r")]
    n = sum(e.shape[0] for e in embs)
    lo, hi = du.shard_range(n, world, r)
    out_e, out_i, start = [], [], 0
    for e, i in zip(embs, idss):
        a, b = max(lo, start), min(hi, start + e.shape[0])
        if a < b:
            out_e.append(np.asarray(e[a - start:b - start])); out_i.append(np.asarray(i[a - start:b - start]))
        start += e.shape[0]
    d = embs[0].shape[1]
    return (np.concatenate(out_e) if out_e else np.zeros((0, d), np.float16),
            np.concatenate(out_i) if out_i else np.zeros((0,), np.int64))


def embed_paths(uniir_dir, embed_dir_name, path_suffix, split, name):
    base = os.path.join(uniir_dir, embed_dir_name, path_suffix, split)
    return (os.path.join(base, f"mbeir_{name}_{split}_embed.npy"), os.path.join(base, f"mbeir_{name}_{split}_ids.npy"))


def save_embeddings(embed_path, ids_path, emb, ids):
    os.makedirs(os.path.dirname(embed_path), exist_ok=True)
    np.save(embed_path, np.asarray(emb, dtype=np.float16))
    np.save(ids_path, np.asarray(ids, dtype=np.int64))


def build_model_from_config(config):
    """The four arms of reference common/utils.py:64-152 (CLIP_SF, CLIP_FF [encode path], BLIP_SF, BLIP_FF) on the B200
    backend (same config fields, same checkpoint layout ``torch.load(path)["model"]``)."""
    name = config.model.name
    mc = config.model
    if name in ("CLIPScoreFusion", "CLIPFeatureFusion"):       # common/utils.py:67-110
        if name == "CLIPScoreFusion":
            from .clip_sf import CLIPScoreFusion as cls
        else:
            from .clip_ff import CLIPFeatureFusion as cls      # encode path only (embedding extraction / retrieval)
        model = cls(model_name=mc.clip_vision_model_name,
                    download_root=os.path.join(config.uniir_dir, mc.pretrained_clip_model_dir))
    elif name in ("BLIPFeatureFusion", "BLIPScoreFusion"):     # common/utils.py:97-152
        if name == "BLIPFeatureFusion":
            from .blip_ff import BLIPFeatureFusion as cls
        else:
            from .blip_sf import BLIPScoreFusion as cls
        model = cls(med_config=os.path.join("../models/uniir_blip", "backbone/configs/med_config.json"),
                    image_size=mc.image_size, vit=mc.vit, vit_grad_ckpt=mc.vit_grad_ckpt,
                    vit_ckpt_layer=mc.vit_ckpt_layer, embed_dim=mc.embed_dim, queue_size=mc.queue_size, config=mc)
    else:
        raise NotImplementedError(f"model {name!r}: the reference knows CLIPScoreFusion, CLIPFeatureFusion, "
                                  "BLIPScoreFusion and BLIPFeatureFusion (common/utils.py:64-152)")
    ckpt = os.path.join(config.uniir_dir, mc.ckpt_config.ckpt_dir, mc.ckpt_config.ckpt_name)
    assert os.path.exists(ckpt), f"Checkpoint file {ckpt} does not exist."
    print(f"loading {name} checkpoint from {ckpt}")
    model.load_state_dict(torch.load(ckpt, map_location="cpu")["model"])
    return model


def main(config):
    """mbeir_embedder.py:464-495: build model, check the duck-typed API, embed every configured split."""
    model = build_model_from_config(config)
    model.eval()
    for attr in ("encode_mbeir_batch", "get_img_preprocess_fn", "get_tokenizer"):
        assert callable(getattr(model, attr, None)), f"The provided model does not have a callable '{attr}' method."
    try:
        from data.mbeir_data_utils import build_mbeir_dataset_from_config  # reference src/ on PYTHONPATH
    except ImportError as e:
        raise ImportError("the M-BEIR dataset/collator classes are the reference's own (src/data); add the "
                          "reference's src/ to PYTHONPATH to run the embedder CLI") from e
    raise NotImplementedError("dataset wiring is delegated to the reference's generate_embeds_for_config; call "
                              "generate_embeds_and_ids_for_dataset_with_gather with its DataLoader (INTEGRATION.md)")


def parse_arguments():
    p = argparse.ArgumentParser(description="Generate Embeddings for MBEIR")
    p.add_argument("--uniir_dir", type=str, default="/data/UniIR")
    p.add_argument("--mbeir_data_dir", type=str, default="/data/UniIR/mbeir_data")
    p.add_argument("--config_path", default="config.yaml", help="Path to the config file.")
    return p.parse_args()


if __name__ == "__main__":
    from . import config as cfg
    args = parse_arguments()
    config = cfg.load(args.config_path)
    config.uniir_dir = args.uniir_dir
    config.mbeir_data_dir = args.mbeir_data_dir
    du.init_distributed_mode(config.dist_config)
    main(config)
