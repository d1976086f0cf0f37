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
    if args.enable_retrieval or args.enable_hard_negative_mining:
        raise NotImplementedError("run_retrieval / run_hard_negative_mining keep the reference's Python metric and "
                                  "TSV code; import search_index from this module into it (INTEGRATION.md)")


if __name__ == "__main__":
    main()
