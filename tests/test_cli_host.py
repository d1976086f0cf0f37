"""Host-side pieces of the embed -> index -> retrieval CLI that need no GPU: the OmegaConf stand-in
on the reference's YAML idioms, file naming, recall@k."""
import os

import numpy as np
import pytest


def test_config_loader_interpolation_and_attribute_access(tmp_path):
    from uniir_b200 import config as cfg
    p = tmp_path / "c.yaml"
    p.write_text('''
experiment:
    instruct_status: "Instruct"
    exp_name: "InBatch"
    description: "${model.name} ${model.size} ${experiment.instruct_status} ${experiment.exp_name}"
    path_suffix: "${model.short_name}/${model.size}/${experiment.instruct_status}/${experiment.exp_name}/"
model:
    name: "CLIPScoreFusion"
    short_name: "CLIP_SF"
    size: "Large"
    gather_embeddings: True
index_config:
    faiss_config: {idx_type: Flat, dim: 768, metric: METRIC_INNER_PRODUCT}
    cand_pools_config:
        enable_idx: True
        cand_pools_name_to_idx: ["visualnews_task0", "UNION"]
seed: 2023
alias: "${seed}"
''')
    c = cfg.load(str(p))
    assert c.experiment.path_suffix == "CLIP_SF/Large/Instruct/InBatch/"
    assert c.experiment.description == "CLIPScoreFusion Large Instruct InBatch"
    assert c.model.gather_embeddings is True and c.index_config.faiss_config.dim == 768
    assert c.alias == 2023 and isinstance(c.alias, int)           # whole-value interpolation keeps the type
    assert c.index_config.cand_pools_config.cand_pools_name_to_idx[1] == "UNION"
    c.uniir_dir = "/data/UniIR"                                    # the CLIs assign onto the config
    assert c.uniir_dir == "/data/UniIR"
    with pytest.raises(AttributeError):
        c.nope


def test_reference_yaml_files_load_unchanged():
    from uniir_b200 import config as cfg
    root = "/root/reference/src/models/uniir_clip/clip_scorefusion/configs_scripts/large"
    if not os.path.isdir(root):
        pytest.skip("reference tree not present (GPU box)")
    for rel in ("eval/inbatch/index.yaml", "eval/inbatch/embed.yaml", "eval/inbatch/retrieval.yaml", "train/inbatch/inbatch.yaml"):
        c = cfg.load(os.path.join(root, rel))
        assert c.model.name == "CLIPScoreFusion"
        assert "${" not in c.experiment.path_suffix


def test_embed_paths_and_recall():
    from uniir_b200.mbeir_embedder import embed_paths
    from uniir_b200.mbeir_retriever import compute_recall_at_k
    e, i = embed_paths("/u", "embed", "CLIP_SF/Large/Instruct/InBatch/", "cand_pool", "union")
    assert e == "/u/embed/CLIP_SF/Large/Instruct/InBatch/cand_pool/mbeir_union_cand_pool_embed.npy"
    assert i.endswith("mbeir_union_cand_pool_ids.npy")
    assert compute_recall_at_k([3, 4], [9, 4, 1], 2) == 1.0 and compute_recall_at_k([3], [9, 4, 3], 2) == 0.0
    assert compute_recall_at_k([], [1], 5) == 0.0


def test_index_file_roundtrip_without_gpu(tmp_path):
    from uniir_b200.mbeir_retriever import write_index
    emb = np.random.default_rng(0).standard_normal((5, 8)).astype(np.float16)
    ids = np.arange(5, dtype=np.int64) * 10_000_000
    p = str(tmp_path / "x" / "mbeir_union_cand_pool.index")
    write_index(p, emb, ids)
    from uniir_b200.mbeir_retriever import index_shape, read_index_rows
    e, i = read_index_rows(p)
    assert index_shape(p) == (5, 8) and (e == emb).all() and (i == ids).all() and e.dtype == np.float16


def _ref_select(retrieved, pos, neg, num_hard_negs):
    """Literal restatement of reference mbeir_retriever.py:667-684 (the oracle for select_hard_negatives)."""
    hard_negatives = [doc_id for doc_id in retrieved if doc_id not in pos and doc_id not in neg]
    if 0 < len(hard_negatives) < num_hard_negs:
        multiplier = num_hard_negs // len(hard_negatives)
        remainder = num_hard_negs % len(hard_negatives)
        hard_negatives = hard_negatives * multiplier + hard_negatives[:remainder]
    return hard_negatives[:num_hard_negs]


def test_hard_negative_selection_matches_reference_rule():
    from uniir_b200.mbeir_retriever import select_hard_negatives
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = int(rng.integers(0, 12))
        retrieved = [f"1:{int(x)}" for x in rng.integers(0, 15, size=k)]
        pos = [f"1:{int(x)}" for x in rng.integers(0, 15, size=int(rng.integers(0, 4)))]
        neg = [f"1:{int(x)}" for x in rng.integers(0, 15, size=int(rng.integers(0, 4)))]
        n = int(rng.integers(1, 9))
        assert select_hard_negatives(retrieved, pos, neg, n) == _ref_select(retrieved, pos, neg, n)
    assert select_hard_negatives([], ["1:1"], [], 3) == []                 # nothing retrieved: stays empty
    assert select_hard_negatives(["1:2"], [], [], 3) == ["1:2"] * 3         # short list repeats cyclically


def test_id_hashing_matches_reference():
    from uniir_b200 import mbeir_retriever as mr
    assert mr.hash_qid("9:123") == 9 * 500000 + 123 and mr.unhash_qid(mr.hash_qid("9:123")) == "9:123"
    assert mr.hash_did("3:9999999") == 3 * 10000000 + 9999999 and mr.unhash_did(np.int64(39999999)) == "3:9999999"
    path = "/root/reference/src/data/preprocessing/utils.py"
    if os.path.exists(path):   # pin against the reference's own constants (not on the GPU box)
        src = open(path).read()
        assert "DATASET_CAN_NUM_UPPER_BOUND = 10000000" in src and "DATASET_QUERY_NUM_UPPER_BOUND = 500000" in src


def test_run_hard_negative_mining_end_to_end_with_injected_search(tmp_path):
    """File formats + flow of run_hard_negative_mining (reference :606-708) on a toy corpus; the search itself
    (GPU kernel, tested in test_search_gpu.py) is injected as exact numpy cosine top-k."""
    import json
    from types import SimpleNamespace as NS
    from uniir_b200 import mbeir_retriever as mr
    rng = np.random.default_rng(1)
    n_c, n_q, d, k = 40, 6, 16, 8
    cand = rng.standard_normal((n_c, d)).astype(np.float16)
    dids = np.array([mr.hash_did(f"2:{i + 1}") for i in range(n_c)], dtype=np.int64)
    q = rng.standard_normal((n_q, d)).astype(np.float16)
    qids = np.array([mr.hash_qid(f"2:{i + 1}") for i in range(n_q)], dtype=np.int64)
    uniir, data = tmp_path / "uniir", tmp_path / "mbeir"
    suffix = "CLIP_SF/Large/Instruct/InBatch/"
    (uniir / "embed" / suffix / "train").mkdir(parents=True)
    np.save(uniir / "embed" / suffix / "train" / "mbeir_agg_train_embed.npy", q)
    np.save(uniir / "embed" / suffix / "train" / "mbeir_agg_train_ids.npy", qids)
    mr.write_index(str(uniir / "index" / suffix / "cand_pool" / "mbeir_union_cand_pool.index"), cand, dids)
    (data / "train").mkdir(parents=True)
    entries = [dict(qid=f"2:{i + 1}", pos_cand_list=[f"2:{i + 1}"], neg_cand_list=[f"2:{i + 2}"]) for i in range(n_q)]
    with open(data / "train" / "mbeir_agg_train.jsonl", "w") as f:
        for e in entries:
            f.write(json.dumps(e) + "\n")

    def np_search(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
        qq = np.load(query_embed_path).astype(np.float32)
        cc, ids = mr.read_index_rows(cand_index_path)
        cc, ids = np.asarray(cc, dtype=np.float32), np.asarray(ids)
        qq /= np.linalg.norm(qq, axis=1, keepdims=True); cc /= np.linalg.norm(cc, axis=1, keepdims=True)
        sc = qq @ cc.T
        order = np.argsort(-sc, axis=1)[:, :num_cand_to_retrieve]
        return np.take_along_axis(sc, order, 1), ids[order]

    cfg = NS(uniir_dir=str(uniir), mbeir_data_dir=str(data), experiment=NS(path_suffix=suffix),
             retrieval_config=NS(embed_dir_name="embed", index_dir_name="index", hard_negs_dir_name="hard_negs", k=k,
                                 num_hard_negs=3, train_datasets_config=NS(enable_retrieve=True, datasets_name=["AGG"],
                                                                           correspond_cand_pools_name=["UNION"])))
    out = mr.run_hard_negative_mining(cfg, search_fn=np_search)
    assert out.endswith("train/hard_negs/mbeir_agg_hard_negs_train.jsonl")
    got = [json.loads(l) for l in open(out)]
    _, I = np_search(str(uniir / "embed" / suffix / "train" / "mbeir_agg_train_embed.npy"),
                     str(uniir / "index" / suffix / "cand_pool" / "mbeir_union_cand_pool.index"), num_cand_to_retrieve=k)
    for i, e in enumerate(got):
        want = _ref_select([mr.unhash_did(x) for x in I[i]], entries[i]["pos_cand_list"], entries[i]["neg_cand_list"], 3)
        assert e["qid"] == entries[i]["qid"] and e["neg_cand_list"] == entries[i]["neg_cand_list"] + want
        assert len(want) == 3 and not set(want) & set(entries[i]["pos_cand_list"])


def test_fp32_queries_are_cast_exactly_or_warned_about():
    """search_index_with_batch takes fp32 and the kernel computes on fp16 -- exact for embeddings
    that were stored as fp16 (every reference call site), a visible warning otherwise."""
    import warnings
    from uniir_b200.mbeir_retriever import _as_fp16_queries
    q16 = np.random.default_rng(0).standard_normal((4, 8)).astype(np.float16)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        out = _as_fp16_queries(q16.astype(np.float32))          # fp16-representable: silent and exact
    assert out.dtype == np.float16 and np.array_equal(out, q16)
    with pytest.warns(UserWarning, match="not exactly representable"):
        _as_fp16_queries(q16.astype(np.float32) + np.float32(1e-4))


def test_word_hash_tokenizer_is_selected_by_env(monkeypatch):
    import torch
    from uniir_b200 import preprocess
    monkeypatch.setenv("UNIIR_CLIP_BPE", "word-hash")
    tok = preprocess.clip_tokenizer(context_length=77, vocab=512)
    t = tok(["A dog.", "a dog. " + "word " * 100, ""])
    assert t.shape == (3, 77) and t.dtype == torch.int32
    assert t[0, 0] == 510 and t[0].max() == 511 and int((t[0] != 0).sum()) == 4       # SOT a dog. EOT
    assert torch.equal(t[0, :3], t[1, :3]) and t[1, 76] == 511 and (t[1] != 0).all()   # truncated, EOT kept last
    assert t[2, 0] == 510 and t[2, 1] == 511 and t[2, 2:].sum() == 0
    monkeypatch.delenv("UNIIR_CLIP_BPE")
    with pytest.raises(FileNotFoundError):
        preprocess.clip_tokenizer()(["needs the BPE file"])
