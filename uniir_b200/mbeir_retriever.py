"""Index + retrieval -- drop-in for the FAISS stage of reference src/common/mbeir_retriever.py.

Same function names, arguments and return values as the reference:
  * ``create_index(config)``                                      mbeir_retriever.py:34-129
  * ``search_index(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10)
        -> (np.float32 [Q,k] descending, np.int64 [Q,k])``         :188-222
  * ``search_index_with_batch(query_embeddings_batch, index, k)``  :225-232
  * ``compute_recall_at_k``                                        :149-166
  * ``load_qrel`` / ``run_retrieval`` (run files, Recall@k per task, TSV) :169-186, :312-603
  * ``run_hard_negative_mining``                                   :606-708
  * CLI ``--enable_create_index / --enable_retrieval / --enable_hard_negative_mining``   :711-765
Instead of a FAISS ``IDMap,Flat`` file holding 4-byte normalised vectors, the ``.index`` file keeps the
stored fp16 embeddings + int64 ids (half the bytes, no normalisation pass: 1/||row|| is recomputed on
the device in microseconds) in a flat layout any row block of which can be memory-mapped, and loaded
indices are cached so the 32 (dataset, pool) iterations of ``run_retrieval`` do not re-read a 17 GB file
each.  Search runs the fused tcgen05 GEMM + top-k kernel; there is no CPU fallback.

Where the candidates live (reference: FAISS shards the index over all visible GPUs inside one process,
:98-100, 204-206):
  * plain ``python -m uniir_b200.mbeir_retriever``   -> ``MultiGPUFlatIPIndex``: one process, every visible GPU;
  * ``python -m torch.distributed.run --nproc_per_node=N -m uniir_b200.mbeir_retriever``  -> one rank per GPU,
    ``ShardedFlatIPIndex``: each rank maps ITS row block of the index file (or of the embedder's per-rank shard
    files), one all-gather of the queries, one all-to-all of the partial lists; rank 0 writes the results.
"""
from __future__ import annotations

import argparse
import csv
import gc
import json
import os
import struct
import warnings
from collections import defaultdict
from datetime import datetime

import numpy as np
import torch

from . import dist_utils as du
from .mbeir_dataset import get_mbeir_task_name, hash_did, hash_qid, unhash_did, unhash_qid  # noqa: F401
from .retrieval import FlatIPIndex, MultiGPUFlatIPIndex, ShardedFlatIPIndex

_INDEX_CACHE = {}
_MAGIC = b"UNIIRIDX"          # .index layout: magic, u64 version, u64 N, u64 d, pad to 64 B | int64 ids [N] | fp16 [N, d]
_HEADER = 64


def write_index(path, emb16, ids):
    """Streams the rows into the flat ``.index`` layout (no second in-memory copy of an 8.6 GB pool)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
    n, d = emb16.shape
    assert ids.shape == (n,)
    with open(path, "wb") as f:
        f.write((_MAGIC + struct.pack("<QQQ", 2, n, d)).ljust(_HEADER, b"\0"))
        f.write(ids.tobytes())
        step = max(1, (64 << 20) // max(1, 2 * d))
        for lo in range(0, n, step):
            f.write(np.ascontiguousarray(emb16[lo:lo + step], dtype=np.float16).tobytes())


_LINK = b"UNIIRLNK"           # .index stub: magic + JSON {"embed": path, "ids": path} -- the index IS the embedder's output


def write_index_link(path, embed_path, ids_path):
    """O(1) "index build" (SURVEY.md 8f-2): the flat inner-product index needs nothing the embedder has not already
    written, so the ``.index`` file can simply point at the embedding / id files (and, through them, at the per-rank
    shard files a distributed embedder run left beside them)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_LINK + json.dumps({"embed": os.path.abspath(embed_path), "ids": os.path.abspath(ids_path)}).encode())


def _link_target(path):
    with open(path, "rb") as f:
        head = f.read(8)
        if head != _LINK:
            return None
        t = json.loads(f.read().decode())
    return t["embed"], t["ids"]


def index_shape(path):
    link = _link_target(path)
    if link is not None:
        return tuple(np.load(link[0], mmap_mode="r").shape)
    with open(path, "rb") as f:
        head = f.read(_HEADER)
    if head[:8] != _MAGIC:
        return None
    _, n, d = struct.unpack("<QQQ", head[8:32])
    return int(n), int(d)


def read_index_rows(path, lo=0, hi=None):
    """(fp16 [hi-lo, d], int64 [hi-lo]) of the rows [lo, hi) of an index file; memory-mapped, so a rank touches
    only its own block.  Also reads the round-1 ``np.savez`` layout."""
    link = _link_target(path)
    if link is not None:
        from .mbeir_embedder import load_rank_shard, shard_paths
        emb, ids = np.load(link[0], mmap_mode="r"), np.load(link[1], mmap_mode="r")
        hi = emb.shape[0] if hi is None else min(hi, emb.shape[0])
        W, r = du.world(), du.rank()
        if W > 1 and (lo, hi) == du.shard_range(emb.shape[0], W, r) and os.path.exists(shard_paths(link[0], link[1], r, W)[0]):
            return load_rank_shard(link[0], link[1], r, W)      # this rank's own shard file, written by the embedder
        return emb[lo:hi], ids[lo:hi]
    shape = index_shape(path)
    if shape is None:                                   # round-1 files: a zip of two arrays
        with np.load(path) as z:
            emb, ids = z["embeddings"], z["ids"]
        hi = len(ids) if hi is None else hi
        return emb[lo:hi], ids[lo:hi]
    n, d = shape
    hi = n if hi is None else min(hi, n)
    ids = np.memmap(path, dtype=np.int64, mode="r", offset=_HEADER, shape=(n,))
    emb = np.memmap(path, dtype=np.float16, mode="r", offset=_HEADER + 8 * n, shape=(n, d))
    return emb[lo:hi], ids[lo:hi]


def read_index(path, device=None):
    """The searchable index of a ``.index`` file, cached by (path, mtime): a ``ShardedFlatIPIndex`` holding this
    rank's row block under an initialised process group, else a ``MultiGPUFlatIPIndex`` over every visible GPU
    (``device`` pins it to one)."""
    key = (os.path.abspath(path), os.path.getmtime(path), du.world(), str(device))
    if key not in _INDEX_CACHE:
        _INDEX_CACHE.clear()  # one resident pool at a time (UNION is 8.6 GB in fp16)
        if du.is_on() and du.world() > 1:
            n, d = index_shape(path) or read_index_rows(path)[0].shape
            lo, hi = du.shard_range(n, du.world(), du.rank())
            emb, ids = read_index_rows(path, lo, hi)
            idx = ShardedFlatIPIndex(d, device or torch.device("cuda", torch.cuda.current_device()))
            idx.ntotal = n
            if hi > lo:
                idx.add_local_shard(np.ascontiguousarray(emb), np.ascontiguousarray(ids))
        else:
            emb, ids = read_index_rows(path)
            idx = FlatIPIndex(emb.shape[1], device) if device is not None else MultiGPUFlatIPIndex(emb.shape[1])
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
        emb = np.load(embed_path, mmap_mode="r")
        ids = np.load(ids_path)
        assert len(ids) == len(set(ids.tolist())), "IDs should be unique"
        faiss_config = index_config.faiss_config
        assert faiss_config.dim == emb.shape[1], "The dimension of the index does not match the dimension of the embeddings!"
        assert faiss_config.metric == "METRIC_INNER_PRODUCT" and faiss_config.idx_type == "Flat", \
            "only the reference's IDMap,Flat + METRIC_INNER_PRODUCT index is implemented"
        index_path = os.path.join(uniir_dir, index_config.index_dir_name, expt_dir_name, split_name,
                                  f"mbeir_{cand_pool_name}_{split_name}.index")
        if index_config.get("link_embeddings", False):       # extension: no 8.6 GB copy, the index points at the .npy pair
            write_index_link(index_path, embed_path, ids_path)
        else:
            write_index(index_path, emb, ids)
        print(f"Successfully indexed {len(ids)} documents")
        print(f"Index saved to: {index_path}")
        del emb, ids
        gc.collect()


def _as_fp16_queries(q):
    """The kernel multiplies fp16 operands exactly in fp32.  Every reference call site hands over embeddings that
    were stored as fp16 (mbeir_embedder.py:56,110; interactive_retriever.py:155-170), so the cast back is exact;
    genuinely fp32 queries would be rounded (2^-11 relative) -- say so instead of doing it silently."""
    q = np.asarray(q)
    if q.dtype == np.float16:
        return q
    q16 = q.astype(np.float16)
    if not np.array_equal(q16.astype(q.dtype), q):
        warnings.warn("query embeddings are not exactly representable in fp16 and are rounded to it (relative error "
                      "2^-11); store / pass fp16 embeddings, as the reference pipeline does, for exact scores")
    return q16


def search_index_with_batch(query_embeddings_batch, index_gpu, num_cand_to_retrieve=10):
    assert isinstance(query_embeddings_batch, np.ndarray) and query_embeddings_batch.dtype == np.float32
    return index_gpu.search(_as_fp16_queries(query_embeddings_batch), num_cand_to_retrieve)


def search_index(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
    """Whole-query-set exact cosine top-k -> (np.float32 [Q,k] descending, np.int64 [Q,k]) on EVERY rank.
    ``batch_size`` is accepted for parity; the fused kernel processes all queries in one launch."""
    query_embeddings = np.load(query_embed_path)
    print(f"Loaded query embeddings from {query_embed_path} with shape: {query_embeddings.shape}")
    index = read_index(cand_index_path)
    print(f"Number of documents in the index: {index.ntotal}")
    k = num_cand_to_retrieve
    if not isinstance(index, ShardedFlatIPIndex):
        return index.search(_as_fp16_queries(query_embeddings), k)
    # one rank per GPU: my contiguous slice of the queries in, my slice of the results out, then everyone gets all
    Q, W, r = query_embeddings.shape[0], du.world(), du.rank()
    lo, hi = du.shard_range(Q, W, r)
    q_local = torch.from_numpy(_as_fp16_queries(query_embeddings[lo:hi])).to(index.device)
    D, I = index.search_sharded(q_local, k)
    per = du.shard_range(Q, W, 0)[1]
    pad_d = torch.full((per, k), -3.4028234663852886e38, device=index.device)
    pad_i = torch.full((per, k), -1, device=index.device, dtype=torch.int64)
    pad_d[:hi - lo], pad_i[:hi - lo] = D, I
    all_d = du.all_gather_rows(pad_d).view(W, per, k)
    all_i = du.all_gather_rows(pad_i).view(W, per, k)
    sizes = [du.shard_range(Q, W, w)[1] - du.shard_range(Q, W, w)[0] for w in range(W)]
    D = torch.cat([all_d[w, :sizes[w]] for w in range(W)]).cpu().numpy()
    I = torch.cat([all_i[w, :sizes[w]] for w in range(W)]).cpu().numpy()
    return D, I


def compute_recall_at_k(relevant_docs, retrieved_indices, k):
    if not relevant_docs:
        return 0.0
    return 1.0 if set(relevant_docs).intersection(set(retrieved_indices[:k])) else 0.0


def load_qrel(filename):
    """qrels: ``qid 0 did relevance task_id`` per line -> ({qid: [did...]} for relevance > 0, {qid: task_id})
    (mbeir_retriever.py:169-186)."""
    qrel, qid_to_taskid = {}, {}
    with open(filename) as f:
        for line in f:
            query_id, _, doc_id, relevance_score, task_id = line.strip().split()
            if int(relevance_score) > 0:
                qrel.setdefault(query_id, []).append(doc_id)
                qid_to_taskid.setdefault(query_id, task_id)
    print(f"Retriever: Loaded {len(qrel)} queries from {filename}")
    print(f"Retriever: Average number of relevant documents per query: "
          f"{sum(len(v) for v in qrel.values()) / max(1, len(qrel)):.2f}")
    return qrel, qid_to_taskid


def write_run_file(path, run_id, hashed_query_ids, distances, indices, qid_to_taskid):
    """TREC-style run file, one line per (query, rank): ``qid Q0 did rank score run_id task_id``
    (mbeir_retriever.py:433-443).  Padding entries (id -1: fewer than k candidates) are skipped."""
    with open(path, "w") as f:
        for q, (dist_row, idx_row) in enumerate(zip(distances, indices)):
            qid = unhash_qid(hashed_query_ids[q])
            task_id = qid_to_taskid[qid]
            for rank, (hashed_doc_id, score) in enumerate(zip(idx_row, dist_row), start=1):
                if hashed_doc_id < 0:
                    continue
                f.write(f"{qid} Q0 {unhash_did(hashed_doc_id)} {rank} {score} {run_id} {task_id}\n")


def get_raw_retrieved_candidates(queries_path, candidates_path, retrieved_indices, hashed_query_ids):
    """{qid: {"query": entry, "candidates": [entries best first]}} for downstream use (mbeir_retriever.py:235-309
    without the complement retriever, which belongs to the UniRAG wrappers -- out of scope)."""
    def load(path, key):
        out = {}
        with open(path) as f:
            for line in f:
                e = json.loads(line.strip())
                assert e[key] not in out, f"{key}s must be unique"
                out[e[key]] = e
        return out
    queries, cands = load(queries_path, "qid"), load(candidates_path, "did")
    out = {}
    for q, idx_row in enumerate(retrieved_indices):
        qid = unhash_qid(hashed_query_ids[q])
        out[qid] = {"query": queries[qid], "candidates": [cands[unhash_did(i)] for i in idx_row if i >= 0]}
    return out


_DATASET_ORDER = {n: i + 1 for i, n in enumerate(
    ["visualnews_task0", "mscoco_task0", "fashion200k_task0", "webqa_task1", "edis_task2", "webqa_task2",
     "visualnews_task3", "mscoco_task3", "fashion200k_task3", "nights_task4", "oven_task6", "infoseek_task6",
     "fashioniq_task7", "cirr_task7", "oven_task8", "infoseek_task8"])}
_RECALL_METRICS = ["Recall@1", "Recall@5", "Recall@10", "Recall@20", "Recall@50"]


def write_results_tsv(path, eval_results):
    """The reference's result sheet (mbeir_retriever.py:506-603): rows sorted by task, dataset, split, pool; one line
    per (task, dataset, split, metric, single pool) with the union-pool value beside it when there is one."""
    rows = sorted(eval_results, key=lambda x: (x["TaskID"], _DATASET_ORDER.get(x["Dataset"].lower(), 99),
                                               {"val": 1, "test": 2}.get(x["Split"].lower(), 99),
                                               {"union": 99}.get(x["CandPool"].lower(), 0)))
    grouped = defaultdict(dict)
    for r in rows:
        key = (r["TaskID"], r["Task"], r["Dataset"], r["Split"])
        grouped[key].setdefault(r["CandPool"], {}).update({m: r.get(m) for m in _RECALL_METRICS})
    out = [["TaskID", "Task", "Dataset", "Split", "Metric", "CandPool", "Value", "UnionPool", "UnionValue"]]
    for (task_id, task, dataset, split), pools in grouped.items():
        union = pools.get("union", {})
        for metric in _RECALL_METRICS:
            for pool, metrics in pools.items():
                if pool == "union" or metrics.get(metric) is None:
                    continue
                out.append([task_id, task, dataset, split, metric, pool, metrics[metric]] +
                           (["union", union.get(metric, "N/A")] if union else ["", ""]))
    with open(path, "w", newline="") as f:
        csv.writer(f, delimiter="\t").writerows(out)
    return out


def run_retrieval(config, query_embedder_config=None):
    """Retrieval + evaluation over every configured (dataset, candidate pool) pair (mbeir_retriever.py:312-603):
    search with k = the largest Recall@k asked for, write the run file (and, with ``raw_retrieval``, the retrieved
    candidates), compute Recall@k per task, write the TSV.  Under ``torch.distributed.run`` every rank searches its
    shard; rank 0 alone writes files.  Returns the list of result dicts."""
    uniir_dir, mbeir_data_dir = config.uniir_dir, config.mbeir_data_dir
    rc = config.retrieval_config
    expt = config.experiment.path_suffix
    main = du.rank() == 0
    results_dir = os.path.join(uniir_dir, rc.results_dir_name, expt)
    run_dir, cands_dir, tsv_dir = (os.path.join(results_dir, n) for n in ("run_files", "retrieved_candidates", "final_tsv"))
    if main:
        for d_ in (run_dir, cands_dir, tsv_dir):
            os.makedirs(d_, exist_ok=True)
    splits = []
    for split_name in ("train", "val", "test"):
        ds = getattr(rc, f"{split_name}_datasets_config", None)
        if ds and ds.enable_retrieve:
            names, pools = ds.datasets_name, ds.correspond_cand_pools_name
            qrels, metrics = getattr(ds, "correspond_qrels_name", None), getattr(ds, "correspond_metrics_name", None)
            assert len(names) == len(pools) == len(qrels) == len(metrics), \
                "Mismatch between datasets and candidate pools and qrels."
            splits.append((split_name, os.path.join(uniir_dir, rc.embed_dir_name, expt, split_name), names, pools, qrels, metrics))
    if main:
        print("-" * 30)
        for split_name, _, names, pools, _, metrics in splits:
            print(f"Split: {split_name}, Retrieval Datasets: {names}, Candidate Pools: {pools}, Metric: {metrics})")
            print("-" * 30)
    eval_results = []
    index_dir = os.path.join(uniir_dir, rc.index_dir_name, expt, "cand_pool")
    qrel_dir = os.path.join(mbeir_data_dir, rc.qrel_dir_name)
    for split, embed_dir, names, pools, qrels, metrics in splits:
        for dataset_name, cand_pool_name, qrel_name, metric_names in zip(names, pools, qrels, metrics):
            print("\n" + "-" * 30)
            print(f"Retriever: Retrieving for query:{dataset_name} | split:{split} | from cand_pool:{cand_pool_name}")
            dataset_name, cand_pool_name, qrel_name = dataset_name.lower(), cand_pool_name.lower(), qrel_name.lower()
            qrel, qid_to_taskid = load_qrel(os.path.join(qrel_dir, split, f"mbeir_{qrel_name}_{split}_qrels.txt"))
            hashed_query_ids = np.load(os.path.join(embed_dir, f"mbeir_{dataset_name}_{split}_ids.npy"))
            embed_query_path = os.path.join(embed_dir, f"mbeir_{dataset_name}_{split}_embed.npy")
            cand_index_path = os.path.join(index_dir, f"mbeir_{cand_pool_name}_cand_pool.index")
            recall_metrics = [m.strip() for m in metric_names.split(",") if "recall" in m.lower()]
            k = max(int(m.split("@")[1]) for m in recall_metrics)
            print(f"Retriever: Searching with k={k}")
            distances, indices = search_index(embed_query_path, cand_index_path, batch_size=hashed_query_ids.shape[0],
                                              num_cand_to_retrieve=k)
            pool_kind = "union_pool" if cand_pool_name == "union" else "single_pool"
            run_id = f"mbeir_{dataset_name}_{pool_kind}_{split}_k{k}"
            if main:
                run_path = os.path.join(run_dir, f"{run_id}_run.txt")
                write_run_file(run_path, run_id, hashed_query_ids, distances, indices, qid_to_taskid)
                print(f"Retriever: Run file saved to {run_path}")
                if rc.get("raw_retrieval", False):
                    if rc.get("retrieve_image_text_pairs", False):
                        raise NotImplementedError("retrieve_image_text_pairs needs the reference's InteractiveRetriever "
                                                  "(UniRAG wrappers, out of scope); raw_retrieval alone is supported")
                    retrieved = get_raw_retrieved_candidates(
                        os.path.join(mbeir_data_dir, rc.query_dir_name, f"{split}/mbeir_{dataset_name}_{split}.jsonl"),
                        os.path.join(mbeir_data_dir, rc.candidate_dir_name, f"mbeir_{cand_pool_name}_{split}_cand_pool.jsonl"),
                        indices, hashed_query_ids)
                    out_path = os.path.join(cands_dir, f"{run_id}_retrieved.jsonl")
                    with open(out_path, "w") as f:
                        for v in retrieved.values():
                            f.write(json.dumps(v) + "\n")
                    print(f"Retriever: Retrieved file saved to {out_path}")
            # Recall@k per task (hit@k: 1 if any relevant document is in the top k, :149-166)
            by_task = defaultdict(lambda: defaultdict(list))
            for q, idx_row in enumerate(indices):
                dids = [unhash_did(i) for i in idx_row if i >= 0]
                qid = unhash_qid(hashed_query_ids[q])
                for m in recall_metrics:
                    by_task[qid_to_taskid[qid]][m].append(compute_recall_at_k(qrel[qid], dids, int(m.split("@")[1])))
            for task_id, recalls in by_task.items():
                result = {"TaskID": int(task_id), "Task": get_mbeir_task_name(int(task_id)), "Dataset": dataset_name,
                          "Split": split, "CandPool": cand_pool_name}
                for m in recall_metrics:
                    result[m] = round(sum(recalls[m]) / len(recalls[m]), 4)
                    print(f"Retriever: Mean {m}: {result[m]}")
                eval_results.append(result)
    if rc.write_to_tsv and main:
        tsv_path = os.path.join(tsv_dir, f"eval_results_{datetime.now().strftime('%m-%d-%H')}.tsv")
        write_results_tsv(tsv_path, eval_results)
        print(f"Retriever: Results saved to {tsv_path}")
    return eval_results


# ----------------------------------------------------------------------------------------------
# Hard-negative mining (reference mbeir_retriever.py:606-708).  The search is the fused
# kernel (whole training query set in one launch, k = retrieval_config.k); the selection rule and the
# jsonl in/out formats are the reference's.
# ----------------------------------------------------------------------------------------------
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
    if du.rank() != 0:
        return out_path
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        for entry in query_data_list:
            f.write(json.dumps(entry) + "\n")
    print(f"MBEIR Train Data with Hard Negatives saved to {out_path}")
    print(f"Total number of entries in {out_path}: {len(query_data_list)}")
    return out_path


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(description="FAISS-free index / retrieval pipeline on B200")
    p.add_argument("--uniir_dir", type=str, default="/data/UniIR")
    p.add_argument("--mbeir_data_dir", type=str, default="/data/UniIR/mbeir_data")
    p.add_argument("--config_path", default="config.yaml")
    p.add_argument("--query_embedder_config_path", default="")
    p.add_argument("--enable_hard_negative_mining", action="store_true")
    p.add_argument("--enable_create_index", action="store_true")
    p.add_argument("--enable_retrieval", action="store_true")
    return p.parse_args(argv)


def main(argv=None):
    """mbeir_retriever.py:731-765.  (The reference reads `query_embedder_config` even when the flag that defines it
    is absent -- an UnboundLocalError at :757; here it is simply None.)"""
    import types
    from . import config as cfg
    args = parse_arguments(argv)
    config = cfg.load(args.config_path)
    config.uniir_dir = args.uniir_dir
    config.mbeir_data_dir = args.mbeir_data_dir
    query_embedder_config = None
    if args.query_embedder_config_path:
        query_embedder_config = cfg.load(args.query_embedder_config_path)
        query_embedder_config.uniir_dir, query_embedder_config.mbeir_data_dir = args.uniir_dir, args.mbeir_data_dir
    dist_args = types.SimpleNamespace(dist_url="env://")
    du.init_distributed_mode(dist_args)          # one rank per GPU under torch.distributed.run; else a single process
    try:
        if args.enable_hard_negative_mining:
            run_hard_negative_mining(config)
        if args.enable_create_index:
            if du.rank() == 0:
                create_index(config)
            if du.is_on():
                torch.distributed.barrier()
        if args.enable_retrieval:
            run_retrieval(config, query_embedder_config)
    finally:
        if dist_args.distributed:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
