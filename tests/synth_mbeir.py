"""A small synthetic M-BEIR tree (SURVEY.md Appendix E record shapes) + the YAML configs of the reference's eval
pipeline, for the dataset / CLI tests: three datasets with different modality patterns, ~200 query rows, JPEG images,
instructions TSV, qrels, and a word-hash tokenizer standing in for clip.tokenize (no BPE vocabulary offline)."""
import json
import os
import zlib

import numpy as np
import torch

# name, dataset id (preprocessing/utils.py:8-19), query modality, candidate modality, task id
DATASETS = [("mscoco_task0", 9, "text", "image", 0),
            ("webqa_task2", 2, "text", "image,text", 2),
            ("cirr_task7", 8, "image,text", "image", 7)]
WORDS = ("red blue green small large dog cat horse car tree house river mountain street person holding riding "
         "standing near under photo picture caption shows with a the of and two three old new").split()


def _sentence(rng, lo=3, hi=12):
    return " ".join(rng.choice(WORDS) for _ in range(int(rng.integers(lo, hi))))


def _image(path, rng, size=72):
    from PIL import Image
    os.makedirs(os.path.dirname(path), exist_ok=True)
    base = rng.integers(0, 255, (size // 8, size // 8 + 3, 3), dtype=np.uint8)       # non-square, blocky
    Image.fromarray(base).resize((size + 24, size), Image.NEAREST).save(path, "JPEG", quality=90)


def make_tree(root, n_queries=67, n_cands=80, seed=0, n_prompts=2):
    """Writes the tree under `root` (= mbeir_data_dir).  Returns {dataset: (queries, candidates)}.
    n_prompts=1: one instruction per task, so the query text does not depend on the per-rank RNG stream."""
    rng = np.random.default_rng(seed)
    out = {}
    os.makedirs(os.path.join(root, "instructions"), exist_ok=True)
    with open(os.path.join(root, "instructions", "query_instructions.tsv"), "w") as f:
        f.write("query_modality\tcand_modality\ttask\tdataset_id\tprompt_1\tprompt_2\tprompt_3\tprompt_4\n")
        for name, ds, qm, cm, task in DATASETS:
            second = f"retrieve the {cm}" if n_prompts > 1 else ""
            f.write(f"{qm}\t{cm}\t{task}\t{ds}\tFind me a matching {cm} for this {qm}\t{second}\t\t\n")
    for name, ds, qm, cm, task in DATASETS:
        cands, queries = [], []
        for j in range(1, n_cands + 1):
            c = {"did": f"{ds}:{j}", "modality": cm, "txt": None, "img_path": None, "src_content": None}
            if "text" in cm:
                c["txt"] = _sentence(rng)
            if "image" in cm:
                c["img_path"] = f"mbeir_images/{name}/c{j}.jpg"
                _image(os.path.join(root, c["img_path"]), rng)
            cands.append(c)
        for i in range(1, n_queries + 1):
            pos = sorted({int(x) for x in rng.integers(1, n_cands + 1, int(rng.integers(1, 3)))})
            q = {"qid": f"{ds}:{i}", "query_txt": None, "query_img_path": None, "query_modality": qm,
                 "query_src_content": None, "pos_cand_list": [f"{ds}:{p}" for p in pos], "neg_cand_list": [],
                 "task_id": task}
            if "text" in qm:
                q["query_txt"] = _sentence(rng) + ("" if i % 7 else "  \r")
            if "image" in qm:
                q["query_img_path"] = f"mbeir_images/{name}/q{i}.jpg"
                _image(os.path.join(root, q["query_img_path"]), rng)
            queries.append(q)
        for split in ("test", "train"):
            os.makedirs(os.path.join(root, "query", split), exist_ok=True)
            with open(os.path.join(root, "query", split, f"mbeir_{name}_{split}.jsonl"), "w") as f:
                for q in queries:
                    f.write(json.dumps(q) + "\n")
        os.makedirs(os.path.join(root, "cand_pool", "local"), exist_ok=True)
        with open(os.path.join(root, "cand_pool", "local", f"mbeir_{name}_cand_pool.jsonl"), "w") as f:
            for c in cands:
                f.write(json.dumps(c) + "\n")
        os.makedirs(os.path.join(root, "qrels", "test"), exist_ok=True)
        with open(os.path.join(root, "qrels", "test", f"mbeir_{name}_test_qrels.txt"), "w") as f:
            for q in queries:
                for d in q["pos_cand_list"]:
                    f.write(f"{q['qid']} 0 {d} 1 {task}\n")
        out[name] = (queries, cands)
    return out


def word_hash_tokenizer(context_length=77, vocab=512):
    """The product's own stand-in tokenizer (uniir_b200.preprocess.word_hash_tokenizer, UNIIR_CLIP_BPE=word-hash)."""
    from uniir_b200.preprocess import word_hash_tokenizer as wh
    return wh(context_length, vocab)


COMMON_YAML = '''
experiment:
    instruct_status: "Instruct"
    exp_name: "InBatch"
    description: "${model.name} ${model.size} ${experiment.instruct_status} ${experiment.exp_name}"
    path_suffix: "${model.short_name}/${model.size}/${experiment.instruct_status}/${experiment.exp_name}/"
model:
    name: "CLIPScoreFusion"
    short_name: "CLIP_SF"
    size: "Tiny"
    clip_vision_model_name: "tiny-test"
    pretrained_clip_model_dir: "checkpoint/CLIP/"
    ckpt_config:
        ckpt_dir: "checkpoint/${experiment.path_suffix}"
        ckpt_name: "clip_sf_tiny.pth"
dist_config:
    dist_url: "env://"
seed: 2023
'''


def _names(key, indent):
    pad = " " * indent
    return "\n".join(f'{pad}- "{n}"' for n in key)


def write_configs(cfg_dir, image_size=56, batch_size=32, link_index=False):
    """embed.yaml / index.yaml / retrieval.yaml with the reference's keys (…/eval/inbatch/*.yaml)."""
    os.makedirs(cfg_dir, exist_ok=True)
    names = [n for n, *_ in DATASETS]
    embed = COMMON_YAML + f'''
embed_config:
    embed_dir_name: "embed"
    use_fp16: True
    train_datasets_config:
        enable_embed: False
        datasets_name:
        correspond_cand_pools_name:
    val_datasets_config:
        enable_embed: False
        datasets_name:
        correspond_cand_pools_name:
    test_datasets_config:
        enable_embed: True
        datasets_name:
{_names(names, 12)}
        correspond_cand_pools_name:
{_names(names, 12)}
    cand_pools_config:
        enable_embed: True
        embed_union_pool: True
        cand_pools_name_to_embed:
{_names(names, 12)}
dataloader_config:
    num_workers: 0
    batch_size: {batch_size}
data_config:
    image_size: {image_size}, {image_size}
    shuffle_cand: True
    enable_query_instruct: True
    train_dir_name: "query/train"
    val_dir_name: "query/val"
    test_dir_name: "query/test"
    cand_pool_dir_name: "cand_pool/local"
    query_instruct_path: "instructions/query_instructions.tsv"
'''
    index = COMMON_YAML + f'''
index_config:
    faiss_config:
        idx_type: Flat
        dim: 64
        metric: METRIC_INNER_PRODUCT
    embed_dir_name: "embed"
    index_dir_name: "index"
    link_embeddings: {str(bool(link_index))}
    cand_pools_config:
        enable_idx: True
        cand_pools_name_to_idx:
{_names(names + ["UNION"], 12)}
'''
    retrieval = COMMON_YAML + f'''
retrieval_config:
    embed_dir_name: "embed"
    index_dir_name: "index"
    results_dir_name: "retrieval_results"
    qrel_dir_name: "qrels"
    write_to_tsv: True
    raw_retrieval: False
    retrieve_image_text_pairs: False
    query_dir_name:
    candidate_dir_name:
    train_datasets_config:
        enable_retrieve: False
        datasets_name:
        correspond_cand_pools_name:
    val_datasets_config:
        enable_retrieve: False
        datasets_name:
        correspond_cand_pools_name:
        correspond_qrels_name:
    test_datasets_config:
        enable_retrieve: True
        datasets_name:
{_names(names + names, 12)}
        correspond_cand_pools_name:
{_names(names + ["UNION"] * len(names), 12)}
        correspond_qrels_name:
{_names(names + names, 12)}
        correspond_metrics_name:
{_names(["Recall@1, Recall@5, Recall@10"] * (2 * len(names)), 12)}
'''
    paths = {}
    for nm, body in (("embed", embed), ("index", index), ("retrieval", retrieval)):
        paths[nm] = os.path.join(cfg_dir, f"{nm}.yaml")
        with open(paths[nm], "w") as f:
            f.write(body)
    return paths
