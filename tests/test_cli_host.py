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
    with np.load(p) as z:
        assert (z["embeddings"] == emb).all() and (z["ids"] == ids).all() and z["embeddings"].dtype == np.float16
