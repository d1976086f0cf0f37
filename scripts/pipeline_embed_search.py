"""BASELINE configs[3] -> configs[4] as ONE pipeline, one rank per GPU (torch.distributed.run):

  1. embedding extraction: `--items` synthetic M-BEIR-shaped candidates (224x224 image + 77-token text, CLIP_SF
     ViT-L/14, inference batch 2048) row-sharded contiguously over the ranks (ContiguousDistributedSampler partition);
     every step uploads its batch from pinned host memory and stores fp16 embeddings;
  2. every rank writes ITS rows as a shard file (mbeir_embedder.save_rank_shards: fp16 rows + int64 ids; only the ids
     are exchanged, for the reference's uniqueness assertion) -- no rank-0 funnel;
  3. every rank loads its own shard file into the row-sharded index (mbeir_embedder.load_rank_shard ->
     retrieval.ShardedFlatIPIndex) and the `--queries` queries are searched top-50: one all-gather of the queries,
     the fused kernel against the local shard, one all-to-all of the partial lists, merge.

Prints one JSON line on rank 0 (device-timed, max over ranks).  The image batch is one fixed pinned host batch
re-uploaded every step; the token ids change every step, so the rows are distinct.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pipeline_embed_search.py \
        --items 5600000 --queries 190000
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from uniir_b200 import dist_utils as du
from uniir_b200 import mbeir_embedder as me
from uniir_b200.clip_sf import CLIPScoreFusion
from uniir_b200.preprocess import synthetic_clip_tokens
from uniir_b200.retrieval import ShardedFlatIPIndex

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=5_600_000)
ap.add_argument("--queries", type=int, default=190_000)
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--out", default=None, help="directory for the shard files (default: a temp dir, removed afterwards)")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
peaks = bench.load_peaks()
out_dir = args.out or (tempfile.mkdtemp(prefix="uniir_pipe_") if rank == 0 else None)
if world > 1:
    box = [out_dir]
    dist.broadcast_object_list(box, src=0)
    out_dir = box[0]
embed_path, ids_path = me.embed_paths(out_dir, "embed", "CLIP_SF/Large/Instruct/InBatch/", "cand_pool", "union")

model = CLIPScoreFusion(bench.MODEL, device=dev, micro_batch=256)
model.eval()
lo, hi = du.shard_range(args.items, world, rank)
B = args.batch
g = torch.Generator().manual_seed(2023 + rank)
img = torch.empty(B, 3, 224, 224, pin_memory=True)
for s in range(0, B, 256):
    img[s:s + 256].normal_(generator=g)
ones = torch.ones(B, dtype=torch.long).pin_memory()
emb_local = torch.empty(hi - lo, model.dims.embed_dim, device=dev, dtype=torch.float16)
ids_local = torch.arange(lo, hi, dtype=torch.int64, device=dev) + 90_000_000      # hash_did-shaped, unique


def embed_rows(a, b):
    n = b - a
    ids = synthetic_clip_tokens(B, g, 77, model.dims.vocab).pin_memory()
    batch = dict(txt_batched=ids, image_batched=img, txt_mask_batched=ones, image_mask_batched=ones,
                 did_list=list(range(lo + a, lo + a + B)))
    emb, _ = model(batch, encode_mbeir_batch=True)       # pinned host batch in: uploaded in chunks by the model
    emb_local[a:b] = emb[:n].half()


# warm-up (graph capture, allocator), untimed
for _ in range(3):
    embed_rows(0, min(B, hi - lo))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_wall = time.time()
e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
e0.record()
for a in range(0, hi - lo, B):
    embed_rows(a, min(hi - lo, a + B))
e1.record()
torch.cuda.synchronize()
t_embed = e0.elapsed_time(e1) / 1e3
# ---- shard files (host I/O, wall clock)
t0 = time.time()
me.save_rank_shards(embed_path, ids_path, emb_local, ids_local)
if world > 1:
    dist.barrier()
t_write = time.time() - t0
del emb_local
torch.cuda.empty_cache()
# ---- index from this rank's own shard file + sharded search
t0 = time.time()
e_np, i_np = me.load_rank_shard(embed_path, ids_path, rank, world)
index = ShardedFlatIPIndex(model.dims.embed_dim, dev)
index.add_local_shard(e_np, i_np)
torch.cuda.synchronize()
t_load = time.time() - t0
qlo, qhi = du.shard_range(args.queries, world, rank)
gd = torch.Generator(device=dev).manual_seed(7 + rank)
q = torch.randn(qhi - qlo, model.dims.embed_dim, generator=gd, device=dev).half()
for _ in range(2):
    index.search_sharded(q, 50)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e2.record()
D, I = index.search_sharded(q, 50)
e3.record()
torch.cuda.synchronize()
t_search = e2.elapsed_time(e3) / 1e3
t = torch.tensor([t_embed, t_write, t_load, t_search, time.time() - t_wall], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
t_embed, t_write, t_load, t_search, t_total = t.tolist()
ok = bool((I >= 90_000_000).all()) and bool((D[:, :-1] >= D[:, 1:]).all())
if rank == 0:
    alg = 175.33e9 * args.items / t_embed / 1e12 / world
    fl = 2.0 * args.queries * args.items * 768 / t_search / 1e12 / world
    print(json.dumps({
        "pipeline": "embed (configs[3]) -> per-rank shard files -> sharded top-50 search (configs[4])",
        "n_gpus": world, "items": args.items, "queries": args.queries,
        "embed_items_per_sec": args.items / t_embed, "embed_s": t_embed,
        "embed_tflops_per_gpu_algorithmic": alg, "embed_frac_of_sustained_bf16": alg / peaks["bf16_sustained"],
        "shard_write_s": t_write, "shard_bytes_per_rank": (hi - lo) * 768 * 2, "shard_load_s": t_load,
        "search_queries_per_sec": args.queries / t_search, "search_s": t_search,
        "search_tflops_per_gpu": fl, "search_frac_of_burst_bf16": fl / peaks["bf16_burst"],
        "total_wall_s_after_warmup": t_total, "results_sane": ok,
        "note": "device-timed phases, max over ranks; H2D of every batch inside the embed phase; ids exchanged once "
                "(uniqueness check), embeddings never leave their rank"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0 and args.out is None:
    shutil.rmtree(out_dir, ignore_errors=True)
