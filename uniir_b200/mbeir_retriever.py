"""Index + retrieval -- drop-in for the FAISS stage of reference src/common/mbeir_retriever.py.

Same function names, arguments and return values as the reference:
  * ``create_index(config)``                                      mbeir_retriever.py:34-129
  * ``search_index(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10)
        -> (np.float32 [Q,k] descending, np.int64 [Q,k])``         :188-222
  * ``search_index_with_batch(query_embeddings_batch, index, k)``  :225-232
  * ``compute_recall_at_k``                                        :149-166
Instead of a FAISS ``IDMap,Flat`` file holding 4-byte normalised vectors, the ``.index`` file keeps the
stored fp16 embeddings + int64 ids (half the bytes, no normalisation pass: 1/||row|| is recomputed on
the device in microseconds), and loaded indices are cached so the 32 (dataset, pool) iterations of
``run_retrieval`` do not re-read a 17 GB file each.  Search runs the fused tcgen05
GEMM + top-k kernel on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import gc
import os

import numpy as np
import torch

from .retrieval import FlatIPIndex

_INDEX_CACHE = {}


def write_index(path, emb16, ids):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:  # keep the reference's ".index" file name (np.savez would append .npz)
        np.savez(f, embeddings=np.asarray(emb16, dtype=np.float16), ids=np.asarray(ids, dtype=np.int64))


def read_index(path, device="cuda"):
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _INDEX_CACHE:
        _INDEX_CACHE.clear()  # one resident pool at a time (UNION is 8.6 GB in fp16)
        with np.load(path) as z:
            emb, ids = z["embeddings"], z["ids"]
        idx = FlatIPIndex(emb.shape[1], device)
        idx.add_with_ids(emb, ids)
        _INDEX_CACHE[key] = idx
    return _INDEX_CACHE[key]


def create_index(config):
    """Builds the flat inner-product index for every configured candidate pool."""
    uniir_dir = config.uniir_dir
    index_config = config.index_config
    expt_dir_name = config.experiment.path_suffix
    pools = index_config.cand_pools_config
    assert pools.enable_idx, "Indexing is not enabled for candidate pool"
    split_name = "cand_pool"
    for cand_pool_name in pools.cand_pools_name_to_idx:
        cand_pool_name = cand_pool_name.lower()
        base = os.path.join(uniir_dir, index_config.embed_dir_name, expt_dir_name, split_name)
        embed_path = os.path.join(base, f"mbeir_{cand_pool_name}_{split_name}_embed.npy")
        ids_path = os.path.join(base, f"mbeir_{cand_pool_name}_{split_name}_ids.npy")
        print(f"Building index for {embed_path} and {ids_path}")
        emb = np.load(embed_path)
        ids = np.load(ids_path)
        assert len(ids) == len(set(ids.tolist())), "IDs should be unique"
        faiss_config = index_config.faiss_config
        assert faiss_config.dim == emb.shape[1], "The dimension of the index does not match the dimension of the embeddings!"
        assert faiss_config.metric == "METRIC_INNER_PRODUCT" and faiss_config.idx_type == "Flat", \
            "only the reference's IDMap,Flat + METRIC_INNER_PRODUCT index is implemented"
        index_path = os.path.join(uniir_dir, index_config.index_dir_name, expt_dir_name, split_name,
                                  f"mbeir_{cand_pool_name}_{split_name}.index")
        write_index(index_path, emb, ids)
        print(f"Successfully indexed {len(ids)} documents")
        print(f"Index saved to: {index_path}")
        del emb, ids
        gc.collect()


def search_index_with_batch(query_embeddings_batch, index_gpu, num_cand_to_retrieve=10):
    assert isinstance(query_embeddings_batch, np.ndarray) and query_embeddings_batch.dtype == np.float32
    return index_gpu.search(query_embeddings_batch, num_cand_to_retrieve)


def search_index(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
    """Whole-query-set exact cosine top-k.  ``batch_size`` is accepted for parity; the fused kernel
    processes all queries in one launch."""
    query_embeddings = np.load(query_embed_path)
    print(f"Loaded query embeddings from {query_embed_path} with shape: {query_embeddings.shape}")
    index = read_index(cand_index_path)
    print(f"Number of documents in the index: {index.ntotal}")
    distances, indices = index.search(query_embeddings, num_cand_to_retrieve)
    return distances, indices


def compute_recall_at_k(relevant_docs, retrieved_indices, k):
    if not relevant_docs:
        return 0.0
    return 1.0 if set(relevant_docs).intersection(set(retrieved_indices[:k])) else 0.0


# ----------------------------------------------------------------------------------------------
# Hard-negative mining (reference mbeir_retriever.py:606-708).  The search is the fused
# kernel (whole training query set in one launch, k = retrieval_config.k); the selection rule and the
# jsonl in/out formats are the reference's.
# ----------------------------------------------------------------------------------------------
DATASET_CAN_NUM_UPPER_BOUND = 10000000   # reference src/data/preprocessing/utils.py:47-48
DATASET_QUERY_NUM_UPPER_BOUND = 500000


def hash_qid(qid):
    dataset_id, data_within_id = map(int, qid.split(":"))
    return dataset_id * DATASET_QUERY_NUM_UPPER_BOUND + data_within_id


def unhash_qid(hashed_qid):
    return f"{int(hashed_qid) // DATASET_QUERY_NUM_UPPER_BOUND}:{int(hashed_qid) % DATASET_QUERY_NUM_UPPER_BOUND}"


def hash_did(did):
    dataset_id, data_within_id = map(int, did.split(":"))
    return dataset_id * DATASET_CAN_NUM_UPPER_BOUND + data_within_id


def unhash_did(hashed_did):
    return f"{int(hashed_did) // DATASET_CAN_NUM_UPPER_BOUND}:{int(hashed_did) % DATASET_CAN_NUM_UPPER_BOUND}"


def select_hard_negatives(retrieved_dids, pos_cand_list, neg_cand_list, num_hard_negs):
    """mbeir_retriever.py:667-684: retrieved documents (best first) that are neither positives nor already
    negatives; a short list is repeated cyclically up to num_hard_negs; then truncated."""
    known = set(pos_cand_list) | set(neg_cand_list)
    hard = [d for d in retrieved_dids if d not in known]
    if 0 < len(hard) < num_hard_negs:
        hard = hard * (num_hard_negs // len(hard)) + hard[:num_hard_negs % len(hard)]
    elif len(hard) == 0:
        print("Warning: hard_negatives list is empty.")
    return hard[:num_hard_negs]


def mine_hard_negatives(query_data_list, query_ids, retrieved_indices, num_hard_negs):
    """Extends every query's neg_cand_list in place (mbeir_retriever.py:658-685).  query_ids: hashed int64 [Q] in
    the order of query_data_list; retrieved_indices: int64 [Q,k] hashed dids, best first (-1 = no result)."""
    assert len(query_ids) == len(retrieved_indices) == len(query_data_list)
    for i, hashed in enumerate(query_ids):
        query_data = query_data_list[i]
        assert unhash_qid(hashed) == query_data["qid"]
        dids = [unhash_did(idx) for idx in retrieved_indices[i] if idx >= 0]
        query_data["neg_cand_list"].extend(
            select_hard_negatives(dids, query_data["pos_cand_list"], query_data["neg_cand_list"], num_hard_negs))
    return query_data_list


def run_hard_negative_mining(config, search_fn=None):
    """Drop-in for the reference function of the same name: reads the training jsonl + query embeddings,
    searches the first candidate pool with k = retrieval_config.k, writes
    {mbeir_data_dir}/train/{hard_negs_dir_name}/mbeir_{dataset}_hard_negs_train.jsonl."""
    import json
    search_fn = search_fn or search_index
    rc = config.retrieval_config
    expt = config.experiment.path_suffix
    tcfg = rc.train_datasets_config
    assert tcfg.enable_retrieve, "Hard negative mining is not enabled for training data"
    dataset_name = tcfg.datasets_name[0].lower()            # only the first dataset / pool, like the reference
    cand_pool_name = tcfg.correspond_cand_pools_name[0].lower()
    split = "train"
    query_data_path = os.path.join(config.mbeir_data_dir, "train", f"mbeir_{dataset_name}_{split}.jsonl")
    with open(query_data_path) as f:
        query_data_list = [json.loads(line) for line in f if line.strip()]
    embed_dir = os.path.join(config.uniir_dir, rc.embed_dir_name, expt, split)
    query_ids = np.load(os.path.join(embed_dir, f"mbeir_{dataset_name}_{split}_ids.npy"))
    embed_path = os.path.join(embed_dir, f"mbeir_{dataset_name}_{split}_embed.npy")
    index_path = os.path.join(config.uniir_dir, rc.index_dir_name, expt, "cand_pool",
                              f"mbeir_{cand_pool_name}_cand_pool.index")
    print("-" * 30)
    print(f"Hard Negative mining, Datasets: {dataset_name}, Candidate Pools: {cand_pool_name}")
    print("-" * 30)
    _, retrieved = search_fn(embed_path, index_path, batch_size=query_ids.shape[0], num_cand_to_retrieve=rc.k)
    mine_hard_negatives(query_data_list, query_ids, retrieved, rc.num_hard_negs)
    out_path = os.path.join(config.mbeir_data_dir, "train", rc.hard_negs_dir_name,
                            f"mbeir_{dataset_name}_hard_negs_{split}.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        for entry in query_data_list:
            f.write(json.dumps(entry) + "\n")
    print(f"MBEIR Train Data with Hard Negatives saved to {out_path}")
    print(f"Total number of entries in {out_path}: {len(query_data_list)}")
    return out_path


def parse_arguments():
    p = argparse.ArgumentParser(description="FAISS-free index / retrieval pipeline on B200")
    p.add_argument("--uniir_dir", type=str, default="/data/UniIR")
    p.add_argument("--mbeir_data_dir", type=str, default="/data/UniIR/mbeir_data")
    p.add_argument("--config_path", default="config.yaml")
    p.add_argument("--query_embedder_config_path", default="")
    p.add_argument("--enable_hard_negative_mining", action="store_true")
    p.add_argument("--enable_create_index", action="store_true")
    p.add_argument("--enable_retrieval", action="store_true")
    return p.parse_args()


def main():
    from . import config as cfg
    args = parse_arguments()
    config = cfg.load(args.config_path)
    config.uniir_dir = args.uniir_dir
    config.mbeir_data_dir = args.mbeir_data_dir
    if args.enable_create_index:
        create_index(config)
    if args.enable_hard_negative_mining:
        run_hard_negative_mining(config)
    if args.enable_retrieval:
        raise NotImplementedError("run_retrieval keeps the reference's Python metric / run-file / TSV code (SURVEY a13); "
                                  "import search_index from this module into it (INTEGRATION.md)")


if __name__ == "__main__":
    main()
