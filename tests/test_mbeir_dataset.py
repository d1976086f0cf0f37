"""uniir_b200/mbeir_dataset.py against the reference's own src/data/mbeir_dataset.py on the same synthetic M-BEIR tree:
identical items and identical collated batches (tensors, masks, index mapping, hashed ids) in TRAIN mode (with hard
negatives and random candidate / prompt choice under the same seed), EVAL mode and for candidate pools; and
uniir_b200.mbeir_retriever.run_retrieval against the reference's run_retrieval on the same search results: identical
run files and TSV.  Reference-side comparisons are skipped where /root/reference does not exist; the contract checks
run everywhere."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_mbeir as sm  # noqa: E402
from uniir_b200 import mbeir_dataset as md  # noqa: E402

REF_SRC = "/root/reference/src"
has_ref = os.path.isdir(os.path.join(REF_SRC, "data"))


def _ref_dataset_module():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import data.mbeir_dataset as ref
    return ref


def _prep(img):
    a = np.asarray(img.convert("RGB").resize((16, 16)), dtype=np.float32) / 255.0
    return torch.from_numpy(a).permute(2, 0, 1)


def _same(a, b, path=""):
    if torch.is_tensor(a):
        assert torch.is_tensor(b) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), path
    elif isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("mbeir"))
    data = sm.make_tree(root, n_queries=23, n_cands=30, seed=1)
    # hard negatives for the TRAIN-mode test
    import json
    for name, ds, *_ in sm.DATASETS:
        path = os.path.join(root, "query", "train", f"mbeir_{name}_train.jsonl")
        rows = [json.loads(l) for l in open(path)]
        for i, r in enumerate(rows):
            r["neg_cand_list"] = [f"{ds}:{1 + (i * 3 + j) % 30}" for j in range(1 + i % 3)]
        with open(path, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return root, data


def test_helpers():
    assert md.hash_qid("9:123") == 9 * 500000 + 123 and md.unhash_qid(md.hash_qid("9:123")) == "9:123"
    assert md.hash_did("2:77") == 20000077 and md.unhash_did(np.int64(20000077)) == "2:77"
    assert md.format_string('  "a dog\r" ') == "A dog." and md.format_string("Why?") == "Why?" and md.format_string(None) == ""
    assert md.get_mbeir_task_id("image,text", "image") == 7 and md.get_mbeir_task_name(2) == "text -> image,text"
    if has_ref:
        sys.path.insert(0, REF_SRC)
        from data.preprocessing import utils as ru
        for s in ["  \"a dog\r\" ", "x", "", None, "ends!", "  multiple   words here "]:
            assert md.format_string(s) == ru.format_string(s)
        for q in ["0:1", "9:499999", "5:12"]:
            assert md.hash_qid(q) == ru.hash_qid(q) and md.hash_did(q) == ru.hash_did(q)
            assert md.unhash_did(ru.hash_did(q)) == ru.unhash_did(ru.hash_did(q))
        assert md.MBEIR_TASK == ru.MBEIR_TASK


def test_batch_contract(tree):
    root, _ = tree
    name = sm.DATASETS[2][0]     # image,text -> image
    ds = md.MBEIRMainDataset(root, f"query/train/mbeir_{name}_train.jsonl", f"cand_pool/local/mbeir_{name}_cand_pool.jsonl",
                             "instructions/query_instructions.tsv", _prep, mode=md.Mode.TRAIN, hard_neg_num=2,
                             returns={"hashed_p_did": True}, print_config=False)
    col = md.MBEIRMainCollator(sm.word_hash_tokenizer(), (16, 16), mode=md.Mode.TRAIN)
    random.seed(0)
    b = col([ds[i] for i in range(4)])
    assert b["txt_batched"].shape == (16, 77) and b["image_batched"].shape == (16, 3, 16, 16)
    assert b["index_mapping"]["query"] == [[0], [4], [8], [12]] and b["index_mapping"]["neg_cand_list"][1] == [6, 7]
    assert b["p_did_list"].dtype == torch.int64 and b["p_did_list"].shape == (4,)
    # candidates of this pool are image-only: text mask 0, zero-padded text; queries have both
    assert b["txt_mask_batched"].tolist()[:4] == [1, 0, 0, 0] and b["image_mask_batched"].tolist()[:4] == [1, 1, 1, 1]
    pool = md.MBEIRCandidatePoolDataset(root, f"cand_pool/local/mbeir_{sm.DATASETS[0][0]}_cand_pool.jsonl", _prep, print_config=False)
    cb = md.MBEIRCandidatePoolCollator(sm.word_hash_tokenizer(), 16)([pool[i] for i in range(3)])
    assert cb["did_list"] == [90000001, 90000002, 90000003] and all(isinstance(i, int) for i in cb["did_list"])
    assert cb["txt_mask_batched"].tolist() == [0, 0, 0] and "index_mapping" not in cb


@pytest.mark.skipif(not has_ref, reason="/root/reference not present")
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_main_dataset_and_collator_match_reference(tree, mode):
    root, _ = tree
    ref = _ref_dataset_module()
    tok = sm.word_hash_tokenizer()
    for name, *_ in sm.DATASETS:
        split = "train" if mode == "train" else "test"
        kw = dict(mbeir_data_dir=root, query_data_path=f"query/{split}/mbeir_{name}_{split}.jsonl",
                  cand_pool_path=f"cand_pool/local/mbeir_{name}_cand_pool.jsonl",
                  query_instruct_path="instructions/query_instructions.tsv", img_preprocess_fn=_prep,
                  enable_query_instruct=True, shuffle_cand=True, print_config=False)
        if mode == "train":
            kw.update(hard_neg_num=2, returns={"hashed_p_did": True})
        else:
            kw.update(returns={"task_id": True})
        mine = md.MBEIRMainDataset(mode=md.Mode.TRAIN if mode == "train" else md.Mode.EVAL, **kw)
        theirs = ref.MBEIRMainDataset(mode=ref.Mode.TRAIN if mode == "train" else ref.Mode.EVAL, **kw)
        assert len(mine) == len(theirs)
        idx = list(range(len(mine)))
        random.seed(11); a = [mine[i] for i in idx]
        random.seed(11); b = [theirs[i] for i in idx]
        _same(a, b, name)
        cm = md.MBEIRMainCollator(tok, (16, 16), mode=md.Mode.TRAIN if mode == "train" else md.Mode.EVAL)
        cr = ref.MBEIRMainCollator(tok, (16, 16), mode=ref.Mode.TRAIN if mode == "train" else ref.Mode.EVAL)
        random.seed(12); ba = cm([mine[i] for i in idx[:9]])
        random.seed(12); bb = cr([theirs[i] for i in idx[:9]])
        _same(ba, bb, name + "/batch")


@pytest.mark.skipif(not has_ref, reason="/root/reference not present")
def test_candidate_pool_dataset_matches_reference(tree):
    root, _ = tree
    ref = _ref_dataset_module()
    tok = sm.word_hash_tokenizer()
    for name, *_ in sm.DATASETS:
        p = f"cand_pool/local/mbeir_{name}_cand_pool.jsonl"
        mine = md.MBEIRCandidatePoolDataset(root, p, _prep, print_config=False)
        theirs = ref.MBEIRCandidatePoolDataset(root, p, _prep, print_config=False)
        a, b = [mine[i] for i in range(len(mine))], [theirs[i] for i in range(len(theirs))]
        _same(a, b, name)
        _same(md.MBEIRCandidatePoolCollator(tok, (16, 16))(a[:7]), ref.MBEIRCandidatePoolCollator(tok, (16, 16))(b[:7]), name)


# ------------------------------------------------------------------------------------------------------------------
# run_retrieval host logic (run files, Recall@k per task, TSV) against the reference's own function
# ------------------------------------------------------------------------------------------------------------------
def _fake_search_results(tree_data, rng, k):
    """Deterministic (D, I) per (dataset, pool): hashed candidate ids drawn from the pool(s), some queries hitting a
    positive, scores descending."""
    res = {}
    names = [n for n, *_ in sm.DATASETS]
    union = [md.hash_did(c["did"]) for n in names for c in tree_data[n][1]]
    for n in names:
        queries, cands = tree_data[n]
        for pool_name, pool in ((n, [md.hash_did(c["did"]) for c in cands]), ("union", union)):
            I = np.stack([rng.permutation(pool)[:k] for _ in queries]).astype(np.int64)
            for qi, q in enumerate(queries):
                if qi % 3 == 0:
                    I[qi, qi % k] = md.hash_did(q["pos_cand_list"][0])
            D = np.sort(rng.random((len(queries), k)).astype(np.float32))[:, ::-1].copy()
            res[(n, pool_name)] = (D, I)
    return res


def _setup_retrieval(tmp_path, tree):
    root, data = tree
    uniir = str(tmp_path / "uniir")
    cfgs = sm.write_configs(str(tmp_path / "cfg"))
    from uniir_b200 import config as cfg
    c = cfg.load(cfgs["retrieval"])
    c.uniir_dir, c.mbeir_data_dir = uniir, root
    embed_dir = os.path.join(uniir, "embed", c.experiment.path_suffix, "test")
    os.makedirs(embed_dir, exist_ok=True)
    for n, *_ in sm.DATASETS:
        q = data[n][0]
        np.save(os.path.join(embed_dir, f"mbeir_{n}_test_ids.npy"), np.array([md.hash_qid(x["qid"]) for x in q], dtype=np.int64))
        np.save(os.path.join(embed_dir, f"mbeir_{n}_test_embed.npy"), np.zeros((len(q), 4), np.float16))
    return c, uniir


def test_run_retrieval_writes_run_files_and_tsv(tmp_path, tree, monkeypatch):
    from uniir_b200 import mbeir_retriever as mr
    c, uniir = _setup_retrieval(tmp_path, tree)
    res = _fake_search_results(tree[1], np.random.default_rng(5), 10)

    def fake_search(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
        n = os.path.basename(query_embed_path)[len("mbeir_"):-len("_test_embed.npy")]
        pool = os.path.basename(cand_index_path)[len("mbeir_"):-len("_cand_pool.index")]
        D, I = res[(n, pool)]
        assert batch_size == D.shape[0] and num_cand_to_retrieve == 10
        return D, I
    monkeypatch.setattr(mr, "search_index", fake_search)
    out = mr.run_retrieval(c)
    assert len(out) == 6 and {r["CandPool"] for r in out} == {"union"} | {n for n, *_ in sm.DATASETS}
    base = os.path.join(uniir, "retrieval_results", c.experiment.path_suffix)
    runs = sorted(os.listdir(os.path.join(base, "run_files")))
    assert "mbeir_mscoco_task0_single_pool_test_k10_run.txt" in runs and "mbeir_cirr_task7_union_pool_test_k10_run.txt" in runs
    line = open(os.path.join(base, "run_files", runs[0])).readline().split()
    assert len(line) == 7 and line[1] == "Q0" and line[3] == "1" and ":" in line[0] and ":" in line[2]
    for r in out:      # every third query was given a positive inside the top 10
        assert r["Recall@10"] >= 0.33 and r["Recall@1"] <= r["Recall@5"] <= r["Recall@10"]
    tsv = os.listdir(os.path.join(base, "final_tsv"))
    assert len(tsv) == 1
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(base, "final_tsv", tsv[0]))]
    assert rows[0] == ["TaskID", "Task", "Dataset", "Split", "Metric", "CandPool", "Value", "UnionPool", "UnionValue"]
    assert len(rows) == 1 + 3 * 3 and all(r[7] == "union" for r in rows[1:])


@pytest.mark.skipif(not has_ref, reason="/root/reference not present")
def test_run_retrieval_matches_reference_function(tmp_path, tree, monkeypatch):
    """The reference's own run_retrieval (imported with `faiss` / `omegaconf` stubbed: neither is installable here)
    and ours, fed the same search results, must write byte-identical run files and TSV rows."""
    from uniir_b200 import mbeir_retriever as mr
    for n in ("faiss", "omegaconf"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object
    for pth in (REF_SRC, os.path.join(REF_SRC, "common")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import mbeir_retriever as ref
    res = _fake_search_results(tree[1], np.random.default_rng(7), 10)

    def fake_search(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
        n = os.path.basename(query_embed_path)[len("mbeir_"):-len("_test_embed.npy")]
        pool = os.path.basename(cand_index_path)[len("mbeir_"):-len("_cand_pool.index")]
        return res[(n, pool)]
    outs = {}
    for tag, mod in (("mine", mr), ("ref", ref)):
        c, uniir = _setup_retrieval(tmp_path / tag, tree)
        monkeypatch.setattr(mod, "search_index", fake_search)
        mod.run_retrieval(c)
        base = os.path.join(uniir, "retrieval_results", c.experiment.path_suffix)
        files = {}
        for sub in ("run_files", "final_tsv"):
            for f in sorted(os.listdir(os.path.join(base, sub))):
                key = f if sub == "run_files" else "tsv"
                files[key] = open(os.path.join(base, sub, f)).read().replace("\r\n", "\n")
        outs[tag] = files
    assert set(outs["mine"]) == set(outs["ref"])
    for k in outs["ref"]:
        assert outs["mine"][k] == outs["ref"][k], k


def test_index_file_formats_roundtrip(tmp_path):
    from uniir_b200 import mbeir_retriever as mr
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((37, 8)).astype(np.float16)
    ids = np.arange(37, dtype=np.int64) * 10_000_000 + 3
    p = str(tmp_path / "x" / "mbeir_union_cand_pool.index")
    mr.write_index(p, emb, ids)
    assert mr.index_shape(p) == (37, 8)
    e, i = mr.read_index_rows(p)
    assert np.array_equal(e, emb) and np.array_equal(i, ids)
    e, i = mr.read_index_rows(p, 10, 21)
    assert np.array_equal(e, emb[10:21]) and np.array_equal(i, ids[10:21])
    # link-style index: points at the embedder's .npy pair
    np.save(str(tmp_path / "e.npy"), emb); np.save(str(tmp_path / "i.npy"), ids)
    q = str(tmp_path / "y" / "mbeir_union_cand_pool.index")
    mr.write_index_link(q, str(tmp_path / "e.npy"), str(tmp_path / "i.npy"))
    assert mr.index_shape(q) == (37, 8)
    e, i = mr.read_index_rows(q, 5, 9)
    assert np.array_equal(e, emb[5:9]) and np.array_equal(i, ids[5:9])
    # round-1 layout (np.savez) still loads
    r = str(tmp_path / "old.index")
    with open(r, "wb") as f:
        np.savez(f, embeddings=emb, ids=ids)
    e, i = mr.read_index_rows(r, 0, 4)
    assert np.array_equal(e, emb[:4]) and np.array_equal(i, ids[:4])


@pytest.mark.skipif(not has_ref, reason="/root/reference not present")
def test_run_hard_negative_mining_matches_reference_function(tmp_path, tree, monkeypatch):
    """The reference's own run_hard_negative_mining (mbeir_retriever.py:606-708, `faiss` / `omegaconf` stubbed) and
    ours, fed the same search results: the mined training file must be byte-identical -- including the repetition rule
    when fewer than num_hard_negs candidates survive the positive / negative filter."""
    import json
    import shutil
    from types import SimpleNamespace as NS
    from uniir_b200 import mbeir_retriever as mr
    for n in ("faiss", "omegaconf"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object
    for pth in (REF_SRC, os.path.join(REF_SRC, "common")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import mbeir_retriever as ref
    root, data = tree
    name, pool = "webqa_task2", "webqa_task2"
    queries, cands = data[name]
    rng = np.random.default_rng(11)
    k, num_hard_negs = 6, 4
    dids = [md.hash_did(c["did"]) for c in cands]
    I = np.stack([rng.permutation(dids)[:k] for _ in queries]).astype(np.int64)
    for qi, q in enumerate(queries):
        pos = [md.hash_did(d) for d in q["pos_cand_list"]]
        if qi % 4 == 0:                         # only two candidates survive the filter: the repetition rule
            I[qi, :] = (pos * k)[:k]
            others = [d for d in dids if d not in pos]
            I[qi, 1], I[qi, 4] = others[qi % len(others)], others[(qi + 7) % len(others)]
        elif qi % 4 == 1:                       # nothing survives: the reference warns and appends nothing
            I[qi, :] = (pos * k)[:k]
    D = np.sort(rng.random(I.shape).astype(np.float32))[:, ::-1].copy()
    suffix = "CLIP_SF/Large/Instruct/InBatch/"
    outs = {}
    for tag, mod in (("mine", mr), ("ref", ref)):
        mbeir, uniir = tmp_path / tag / "mbeir", tmp_path / tag / "uniir"
        (mbeir / "train").mkdir(parents=True)
        (mbeir / "cand_pool").mkdir(parents=True)
        shutil.copy(os.path.join(root, "query", "train", f"mbeir_{name}_train.jsonl"), mbeir / "train" / f"mbeir_{name}_train.jsonl")
        shutil.copy(os.path.join(root, "cand_pool", "local", f"mbeir_{pool}_cand_pool.jsonl"), mbeir / "cand_pool" / f"mbeir_{pool}_cand_pool.jsonl")
        e = uniir / "embed" / suffix / "train"
        e.mkdir(parents=True)
        np.save(e / f"mbeir_{name}_train_ids.npy", np.array([md.hash_qid(q["qid"]) for q in queries], dtype=np.int64))
        np.save(e / f"mbeir_{name}_train_embed.npy", np.zeros((len(queries), 4), np.float16))
        cfg = NS(uniir_dir=str(uniir), mbeir_data_dir=str(mbeir), experiment=NS(path_suffix=suffix),
                 retrieval_config=NS(embed_dir_name="embed", index_dir_name="index", hard_negs_dir_name="hard_negs", k=k,
                                     num_hard_negs=num_hard_negs,
                                     train_datasets_config=NS(enable_retrieve=True, datasets_name=[name.upper()],
                                                              correspond_cand_pools_name=[pool.upper()])))

        def fake_search(query_embed_path, cand_index_path, batch_size=10, num_cand_to_retrieve=10):
            assert batch_size == len(queries) and num_cand_to_retrieve == k
            assert cand_index_path.endswith(f"index/{suffix}cand_pool/mbeir_{pool}_cand_pool.index")
            return D, I
        monkeypatch.setattr(mod, "search_index", fake_search)
        mod.run_hard_negative_mining(cfg)
        outs[tag] = open(mbeir / "train" / "hard_negs" / f"mbeir_{name}_hard_negs_train.jsonl").read()
    assert outs["mine"] == outs["ref"]
    rows = [json.loads(l) for l in outs["mine"].splitlines()]
    before = [json.loads(l) for l in open(os.path.join(root, "query", "train", f"mbeir_{name}_train.jsonl"))]
    assert len(rows) == len(queries) == len(before)
    for qi, (new, old) in enumerate(zip(rows, before)):
        mined = new["neg_cand_list"][len(old["neg_cand_list"]):]
        assert new["neg_cand_list"][:len(old["neg_cand_list"])] == old["neg_cand_list"]
        assert not set(mined) & (set(old["pos_cand_list"]) | set(old["neg_cand_list"]))
        if qi % 4 == 1:
            assert mined == []                                   # nothing survived the filter
        else:
            assert len(mined) == num_hard_negs                   # repeated up to the length when only 1-2 survived
            if qi % 4 == 0:
                assert len(set(mined)) <= 2
