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
``generate_embeds_for_config`` / ``main`` / ``cli`` are the reference CLI body (mbeir_embedder.py:195-527) over this
package's own M-BEIR dataset / collator classes (``mbeir_dataset.py``): the reference tree is not needed at run time.
"""
from __future__ import annotations

import argparse
import gc
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


def _concat_rank_shards(embed_path, ids_path, world):
    """Rank 0: stitch the per-rank shard files into the reference's single .npy pair on the HOST (streamed through
    a memory map; nothing is funnelled through a GPU or a collective), keeping the reference's global uniqueness
    check (mbeir_embedder.py:103-108)."""
    parts = [shard_paths(embed_path, ids_path, r, world) for r in range(world)]
    embs = [np.load(e, mmap_mode="r") for e, _ in parts]
    ids = np.concatenate([np.load(i) for _, i in parts]) if parts else np.zeros((0,), np.int64)
    n, d = sum(e.shape[0] for e in embs), embs[0].shape[1]
    assert len(ids) == n
    assert len(np.unique(ids)) == n, "Hashed IDs should be unique"
    os.makedirs(os.path.dirname(embed_path), exist_ok=True)
    out = np.lib.format.open_memmap(embed_path, mode="w+", dtype=np.float16, shape=(n, d))
    at = 0
    for e in embs:
        out[at:at + e.shape[0]] = e
        at += e.shape[0]
    out.flush()
    del out
    np.save(ids_path, ids.astype(np.int64, copy=False))
    return n


def _splits_to_embed(config):
    """(split, split_dir, dataset names, candidate-pool names) per enabled block of embed_config
    (mbeir_embedder.py:212-240)."""
    embed_config, data_config = config.embed_config, config.data_config
    splits = []
    for split_name in ("train", "val", "test"):
        ds_cfg = getattr(embed_config, f"{split_name}_datasets_config", None)
        if ds_cfg and ds_cfg.enable_embed:
            names, pools = ds_cfg.datasets_name, ds_cfg.correspond_cand_pools_name
            assert len(names) == len(pools), "Mismatch between datasets and candidate pools."
            splits.append((split_name, getattr(data_config, f"{split_name}_dir_name"), names, pools))
    pool_cfg = embed_config.cand_pools_config
    if pool_cfg and pool_cfg.enable_embed:
        pools = pool_cfg.cand_pools_name_to_embed
        splits.append(("cand_pool", data_config.cand_pool_dir_name, [None] * len(pools), pools))
    return splits


@torch.no_grad()
def generate_embeds_for_config(model, img_preprocess_fn, tokenizer, config):
    """Embeds every configured query split and candidate pool (mbeir_embedder.py:195-461).  Each rank embeds its
    contiguous row block (ContiguousDistributedSampler) and writes it as a shard file next to the target; rank 0
    then stitches the shards into the reference's ``mbeir_{name}_{split}_{embed,ids}.npy`` pair on the host.  The
    shard files stay: ``mbeir_retriever.search_index`` under ``torch.distributed.run`` loads them straight into the
    row-sharded index (no rank-0 funnel).  The union pool is the concatenation of the single pools'
    files (mbeir_embedder.py:388-461)."""
    from torch.utils.data import DataLoader
    from .mbeir_dataset import (MBEIRCandidatePoolCollator, MBEIRCandidatePoolDataset, MBEIRMainCollator,
                                MBEIRMainDataset, Mode)
    uniir_dir, mbeir_data_dir = config.uniir_dir, config.mbeir_data_dir
    embed_config, data_config = config.embed_config, config.data_config
    embed_dir_name, expt = embed_config.embed_dir_name, config.experiment.path_suffix
    cand_pool_dir = data_config.cand_pool_dir_name
    image_size = tuple(map(int, str(data_config.image_size).split(",")))
    main, W, r = du.rank() == 0, du.world(), du.rank()
    device = config.dist_config.gpu_id
    splits = _splits_to_embed(config)
    if main:
        print("-" * 30)
        for split_name, split_dir, names, pools in splits:
            what = f"Candidate pools to embed: {pools}" if split_name == "cand_pool" else f"Datasets to embed: {names}"
            print(f"Split: {split_name}, Split dir: {split_dir}, {what}")
            print("-" * 30)
    for split_name, split_dir, names, pools in splits:
        for dataset_name, cand_pool_name in zip(names, pools):
            cand_pool_name = cand_pool_name.lower()
            if split_name == "cand_pool":
                cand_path = os.path.join(cand_pool_dir, f"mbeir_{cand_pool_name}_{split_name}.jsonl")
                if main:
                    print(f"\nEmbedder Log: Generating embeddings for {cand_path}...")
                dataset = MBEIRCandidatePoolDataset(mbeir_data_dir=mbeir_data_dir, cand_pool_data_path=cand_path,
                                                    img_preprocess_fn=img_preprocess_fn, print_config=main)
                collator = MBEIRCandidatePoolCollator(tokenizer=tokenizer, image_size=image_size)
                mid_name = cand_pool_name
            else:
                dataset_name = dataset_name.lower()
                query_path = os.path.join(split_dir, f"mbeir_{dataset_name}_{split_name}.jsonl")
                cand_path = os.path.join(cand_pool_dir, f"mbeir_{cand_pool_name}_cand_pool.jsonl")
                if main:
                    print(f"\nEmbedder Log: Generating embeddings for {query_path} with {cand_path}...")
                dataset = MBEIRMainDataset(mbeir_data_dir=mbeir_data_dir, query_data_path=query_path,
                                           cand_pool_path=cand_path, query_instruct_path=data_config.query_instruct_path,
                                           img_preprocess_fn=img_preprocess_fn, mode=Mode.EVAL,
                                           enable_query_instruct=data_config.enable_query_instruct,
                                           shuffle_cand=data_config.shuffle_cand, print_config=main)
                collator = MBEIRMainCollator(tokenizer=tokenizer, image_size=image_size, mode=Mode.EVAL)
                mid_name = dataset_name
            sampler = du.ContiguousDistributedSampler(dataset, num_replicas=W, rank=r)   # dataset is in sorted order
            loader = DataLoader(dataset, batch_size=config.dataloader_config.batch_size,
                                num_workers=config.dataloader_config.num_workers, pin_memory=True, sampler=sampler,
                                shuffle=False, collate_fn=collator, drop_last=False)
            if du.is_on():
                dist.barrier()
            embed_path, ids_path = embed_paths(uniir_dir, embed_dir_name, expt, split_name, mid_name)
            emb, ids = generate_local_embeds_and_ids(model, loader, device)
            if W > 1:
                save_rank_shards(embed_path, ids_path, emb, ids)
                dist.barrier()                                   # every shard is on disk
                if main:
                    n = _concat_rank_shards(embed_path, ids_path, W)
            else:
                assert torch.unique(ids).numel() == ids.numel(), "Hashed IDs should be unique"
                save_embeddings(embed_path, ids_path, emb.cpu().numpy(), ids.cpu().numpy())
                n = ids.numel()
            if main:
                print(f"Embedder Log: Embedding list length: {n}")
                print(f"Embedder Log: Saved embeddings to {embed_path}.")
                print(f"Embedder Log: Saved ids to {ids_path}.")
            if du.is_on():
                dist.barrier()
            del emb, ids, loader, dataset, collator, sampler
            gc.collect()
            torch.cuda.empty_cache()
        if split_name == "cand_pool" and embed_config.cand_pools_config.embed_union_pool:
            if main:   # the union (global) pool is the concatenation of the single pools' saved files
                print("\nEmbedder Log: Generating embeddings for union pool...")
                all_e, all_i = [], []
                for cand_pool_name in pools:
                    e_path, i_path = embed_paths(uniir_dir, embed_dir_name, expt, split_name, cand_pool_name.lower())
                    all_e.append(np.load(e_path))
                    all_i.append(np.load(i_path))
                    print(f"Embedder Log: Concatenating embeddings from {e_path} and ids from {i_path}.")
                all_e, all_i = np.concatenate(all_e, axis=0), np.concatenate(all_i, axis=0)
                assert len(all_e) == len(all_i), "Mismatch between embeddings and IDs length."
                e_path, i_path = embed_paths(uniir_dir, embed_dir_name, expt, split_name, "union")
                save_embeddings(e_path, i_path, all_e, all_i)
                print(f"Embedder Log: Saved embeddings to {e_path}.")
                print(f"Embedder Log: Saved ids to {i_path}.")
                del all_e, all_i
                gc.collect()
            if du.is_on():
                dist.barrier()


def main(config):
    """mbeir_embedder.py:464-495: seed, build the model, check the duck-typed API, embed every configured split.
    No DDP wrapper: an inference-only model has nothing to all-reduce."""
    import random
    seed = config.seed + du.rank()
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    model = build_model_from_config(config)
    model.eval()
    if not callable(getattr(model, "encode_mbeir_batch", None)):
        raise AttributeError("The provided model does not have a callable 'encode' method.")
    if not callable(getattr(model, "get_img_preprocess_fn", None)):
        raise AttributeError("The provided model does not have an 'img_preprocess_fn' attribute.")
    if not callable(getattr(model, "get_tokenizer", None)):
        raise AttributeError("The provided model does not have a 'tokenizer' attribute.")
    print(f"Model is set up on GPU {config.dist_config.gpu_id}.")
    generate_embeds_for_config(model=model, img_preprocess_fn=model.get_img_preprocess_fn(),
                               tokenizer=model.get_tokenizer(), config=config)


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(description="Generate Embeddings for MBEIR")
    p.add_argument("--uniir_dir", type=str, default="/data/UniIR")
    p.add_argument("--mbeir_data_dir", type=str, default="/data/UniIR/mbeir_data")
    p.add_argument("--config_path", default="config.yaml", help="Path to the config file.")
    return p.parse_args(argv)


def cli(argv=None):
    """`python -m [torch.distributed.run --nproc_per_node=N -m] uniir_b200.mbeir_embedder --config_path ...`
    (mbeir_embedder.py:506-527)."""
    from . import config as cfg
    args = parse_arguments(argv)
    config = cfg.load(args.config_path)
    config.uniir_dir = args.uniir_dir
    config.mbeir_data_dir = args.mbeir_data_dir
    args.dist_url = config.dist_config.dist_url
    du.init_distributed_mode(args)
    config.dist_config.gpu_id = args.gpu
    config.dist_config.distributed_mode = args.distributed
    if du.rank() == 0:
        import yaml
        print(yaml.safe_dump(config.to_dict(), sort_keys=False))
    try:
        main(config)
    finally:
        if config.dist_config.distributed_mode:
            dist.destroy_process_group()


if __name__ == "__main__":
    cli()
