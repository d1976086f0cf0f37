"""M-BEIR datasets and collators: the producer of the hot path's ``batch`` dict.

Host-side mirror of reference src/data/mbeir_dataset.py (datasets :118-438, collators :441-610) and of the id /
string helpers of src/data/preprocessing/utils.py:47-117, written so that the embed -> index -> retrieval CLI and the
training loop of this package run without the reference tree on PYTHONPATH.  The contract is the batch layout:

  txt_batched          tokenizer(list[str])      (CLIP: int [B,77]; BLIP: BatchEncoding with input_ids / attention_mask)
  image_batched        f32 [B,3,H,W]             missing images are an all-zero (black) image with mask 0
  txt_mask_batched     i64 [B]                   0 where the text is None / ""
  image_mask_batched   i64 [B]
  index_mapping        {"query": [[i]]*bs, "pos_cand": [[j]]*bs, "neg_cand_list": [[...]]*bs}   (train)
  qid_list / did_list  list[int]                 hashed ids (eval / candidate pool)
  p_did_list           i64 [bs]                  hashed positive-candidate ids (BLIP train)

tests/test_mbeir_dataset.py runs these classes next to the reference's own on the same synthetic M-BEIR tree and
requires identical batches.
"""
from __future__ import annotations

import json
import os
import random
from enum import Enum

import torch
from torch.utils.data import Dataset

DATASET_CAN_NUM_UPPER_BOUND = 10000000    # preprocessing/utils.py:47
DATASET_QUERY_NUM_UPPER_BOUND = 500000    # preprocessing/utils.py:48

MBEIR_TASK = {                            # preprocessing/utils.py:22-32
    "text -> image": 0, "text -> text": 1, "text -> image,text": 2, "image -> text": 3, "image -> image": 4,
    "image -> text,image": 5, "image,text -> text": 6, "image,text -> image": 7, "image,text -> image,text": 8,
}


def hash_qid(qid):
    ds, within = map(int, qid.split(":"))
    return ds * DATASET_QUERY_NUM_UPPER_BOUND + within


def unhash_qid(hashed):
    hashed = int(hashed)
    return f"{hashed // DATASET_QUERY_NUM_UPPER_BOUND}:{hashed % DATASET_QUERY_NUM_UPPER_BOUND}"


def hash_did(did):
    ds, within = map(int, did.split(":"))
    return ds * DATASET_CAN_NUM_UPPER_BOUND + within


def unhash_did(hashed):
    hashed = int(hashed)
    return f"{hashed // DATASET_CAN_NUM_UPPER_BOUND}:{hashed % DATASET_CAN_NUM_UPPER_BOUND}"


def get_mbeir_task_id(source_modality, target_modality):
    return MBEIR_TASK.get(f"{source_modality} -> {target_modality}")


def get_mbeir_task_name(task_id):
    for name, tid in MBEIR_TASK.items():
        if tid == task_id:
            return name
    return None


def format_string(s):
    """Strip, drop carriage returns and surrounding double quotes, capitalise, end with punctuation
    (preprocessing/utils.py:110-116)."""
    s = (s or "").replace("\r", "").strip().strip('"')
    if s:
        s = s[0].upper() + s[1:]
        if s[-1] not in ".?!":
            s += "."
    return s


class Mode(Enum):
    TRAIN = "train"
    EVAL = "eval"


def _read_jsonl(mbeir_data_dir, rel_path):
    path = os.path.join(mbeir_data_dir, rel_path)
    assert os.path.exists(path), f"Data Path {path} does not exist"
    assert path.endswith(".jsonl"), f"Data Path {path} is not a jsonl file"
    with open(path) as f:
        return [json.loads(line) for line in f]


def _read_instructions(mbeir_data_dir, rel_path):
    """query_instructions.tsv: columns query_modality, cand_modality, ?, dataset_id, prompt_1 ... ; key
    "dataset_id, query_modality, cand_modality" -> non-empty prompts (mbeir_dataset.py:76-92)."""
    path = os.path.join(mbeir_data_dir, rel_path)
    assert os.path.exists(path), f"Instructions Path {path} does not exist"
    assert path.endswith(".tsv"), f"Instructions Path {path} is not a tsv file"
    table = {}
    with open(path) as f:
        next(f)
        for line in f.readlines():
            cols = line.strip().split("\t")
            table[f"{cols[3]}, {cols[0]}, {cols[1]}"] = [p for p in cols[4:] if p]
    return table


class _Base(Dataset):
    def __init__(self, mbeir_data_dir, img_preprocess_fn):
        self.mbeir_data_dir = mbeir_data_dir
        self.img_preprocess_fn = img_preprocess_fn or (lambda x: x)

    def _image(self, rel_path):
        if not rel_path:
            return None
        from PIL import Image
        path = os.path.join(self.mbeir_data_dir, rel_path)
        assert os.path.exists(path), f"Image Path {path} does not exist"
        return self.img_preprocess_fn(Image.open(path).convert("RGB"))

    def _item(self, txt, img_path):
        return {"txt": txt, "img": self._image(img_path)}

    def _prompt(self, dataset_id, query_modality, cand_modality):
        key = f"{dataset_id}, {query_modality}, {cand_modality}"
        prompts = self.query_instructions.get(key, [])
        assert prompts, f"Cannot find prompts for {key}"
        prompt = format_string(random.choice(prompts))
        assert prompt, f"Prompt is empty for {key}"
        return prompt


class MBEIRMainDataset(_Base):
    """Queries with their positive (and, in TRAIN mode, negative) candidates (mbeir_dataset.py:118-291)."""

    def __init__(self, mbeir_data_dir, query_data_path, cand_pool_path, query_instruct_path, img_preprocess_fn,
                 mode=Mode.TRAIN, enable_query_instruct=True, shuffle_cand=True, hard_neg_num=0, returns=None,
                 print_config=True):
        super().__init__(mbeir_data_dir, img_preprocess_fn)
        self.query_data = _read_jsonl(mbeir_data_dir, query_data_path)
        self.cand_pool = {}
        for entry in _read_jsonl(mbeir_data_dir, cand_pool_path):
            assert entry.get("did"), f"Cannot find did for {entry}"
            self.cand_pool[entry["did"]] = entry
        self.query_instructions = _read_instructions(mbeir_data_dir, query_instruct_path)
        self.mode, self.shuffle_cand, self.hard_neg_num = mode, shuffle_cand, hard_neg_num
        self.enable_query_instruct = enable_query_instruct
        self.returns = {"hashed_qid": True, "task_id": False, "hashed_p_did": False, **(dict(returns) if returns else {})}
        if print_config:
            print(f"\n---Mbeir Dataset Config---\nMode: {mode}\nQuery Data Path: {query_data_path}\n"
                  f"Candidate Pool Path: {cand_pool_path}\nEnable Query Instructions: {enable_query_instruct}\n"
                  f"Shuffle Candidates: {shuffle_cand}\nHard Negative Number: {hard_neg_num}\nReturns: {self.returns}\n"
                  "--------------------------\n")

    def __len__(self):
        return len(self.query_data)

    def __getitem__(self, index):
        entry = self.query_data[index]
        qid = entry.get("qid")
        dataset_id = qid.split(":")[0] if qid else None
        query_modality = entry.get("query_modality")
        pos_list = entry.get("pos_cand_list", [])
        assert len(pos_list) > 0, f"Cannot find positive candidates for {entry}"
        if self.mode == Mode.EVAL:   # OVEN / INFOSEEK: only positives of the query's own dataset (:203-208)
            pos_list = [d for d in pos_list if d.split(":")[0] == dataset_id]
        pos_did = random.choice(pos_list) if self.shuffle_cand else pos_list[0]
        pos = self.cand_pool.get(pos_did)
        assert pos, f"Cannot find positive candidate {pos_did} for {entry}"
        pos_modality = pos.get("modality")
        pos_txt = format_string(pos.get("txt") or "")
        prompt = self._prompt(dataset_id, query_modality, pos_modality)       # drawn even when unused (RNG parity)
        query_txt = entry.get("query_txt") or ""
        txt = format_string(f"{prompt} {query_txt}") if self.enable_query_instruct else format_string(query_txt)

        negs = []
        if self.mode == Mode.TRAIN and self.hard_neg_num > 0:
            neg_ids = entry.get("neg_cand_list", [])
            assert len(neg_ids) > 0, f"Cannot find negative candidates for {entry}"
            if self.shuffle_cand:
                random.shuffle(neg_ids)
            for i in range(self.hard_neg_num):                                 # wraps around a short list (:238-241)
                neg = self.cand_pool.get(neg_ids[i % len(neg_ids)])
                neg["txt"] = format_string(neg.get("txt") or "")
                negs.append(neg)

        instance = {"query": self._item(txt, entry.get("query_img_path"))}
        if self.mode == Mode.EVAL:
            if self.returns.get("hashed_qid"):
                instance["qid"] = hash_qid(qid)
            if self.returns.get("task_id"):
                instance["task_id"] = get_mbeir_task_id(query_modality, pos_modality)
        else:
            if self.returns.get("hashed_p_did"):
                instance["p_did"] = hash_did(pos_did)
            instance["pos_cand"] = self._item(pos_txt, pos.get("img_path"))
            if negs:
                instance["neg_cand_list"] = [self._item(n["txt"], n.get("img_path")) for n in negs]
        return instance


class MBEIRCandidatePoolDataset(_Base):
    """One candidate per item (mbeir_dataset.py:368-420)."""

    def __init__(self, mbeir_data_dir, cand_pool_data_path, img_preprocess_fn, returns=None, print_config=True):
        super().__init__(mbeir_data_dir, img_preprocess_fn)
        self.cand_pool = _read_jsonl(mbeir_data_dir, cand_pool_data_path)
        self.returns = {"src_content": False, "hashed_did": True, **(dict(returns) if returns else {})}
        if print_config:
            print(f"\n---Mbeir Candidate Pool Dataset Config---\nCandidate Pool Path: {cand_pool_data_path}\n"
                  f"Returns: {self.returns}\n--------------------------\n")

    def __len__(self):
        return len(self.cand_pool)

    def __getitem__(self, index):
        entry = self.cand_pool[index]
        instance = {"txt": format_string(entry.get("txt") or ""), "img": self._image(entry.get("img_path")),
                    "modality": entry.get("modality")}
        if self.returns.get("hashed_did"):
            instance["did"] = hash_did(entry.get("did"))
        if self.returns.get("src_content"):
            instance["src_content"] = entry.get("src_content")
        return instance


class _CollatorBase:
    def __init__(self, tokenizer, image_size):
        if not callable(tokenizer):
            raise TypeError("tokenizer must be a callable taking a list of strings")
        self.tokenizer = tokenizer
        self.H, self.W = (image_size, image_size) if isinstance(image_size, int) else image_size
        self.padded_image = torch.zeros((3, self.H, self.W))

    def _rows(self):
        return {"txt": [], "tm": [], "img": [], "im": []}

    def _push(self, rows, item):
        txt, img = item["txt"], item["img"]
        has_txt, has_img = txt not in (None, ""), img is not None
        rows["txt"].append(txt if has_txt else "")
        rows["tm"].append(int(has_txt))
        rows["img"].append(img if has_img else self.padded_image)
        rows["im"].append(int(has_img))

    def _finish(self, rows):
        out = {"txt_batched": self.tokenizer(rows["txt"]), "image_batched": torch.stack(rows["img"], dim=0),
               "txt_mask_batched": torch.tensor(rows["tm"], dtype=torch.long),
               "image_mask_batched": torch.tensor(rows["im"], dtype=torch.long)}
        t = out["txt_batched"]
        n = t["input_ids"].size(0) if hasattr(t, "input_ids") else len(t)
        assert n == out["image_batched"].size(0) == out["txt_mask_batched"].size(0) == out["image_mask_batched"].size(0)
        return out


class MBEIRMainCollator(_CollatorBase):
    """Queries and candidates of a batch interleaved in one tensor, with the index mapping
    (mbeir_dataset.py:464-546)."""

    def __init__(self, tokenizer, image_size, mode=Mode.TRAIN):
        super().__init__(tokenizer, image_size)
        self.mode = mode

    def __call__(self, batch):
        keys = ["query"]
        qids, task_ids, p_dids = [], [], []
        if self.mode == Mode.EVAL:
            for inst in batch:
                q, t = inst.pop("qid", None), inst.pop("task_id", None)
                if q is not None:
                    qids.append(q)
                if t is not None:
                    task_ids.append(t)
        else:
            for inst in batch:
                p = inst.pop("p_did", None)
                if p is not None:
                    p_dids.append(p)
            keys.append("pos_cand")
            if "neg_cand_list" in batch[0]:
                keys.append("neg_cand_list")
        index_mapping = {k: [[] for _ in batch] for k in keys}
        rows, counter = self._rows(), 0
        for i, inst in enumerate(batch):
            for k in keys:
                for item in (inst[k] if k == "neg_cand_list" else [inst[k]]):
                    index_mapping[k][i].append(counter)
                    counter += 1
                    self._push(rows, item)
        out = self._finish(rows)
        out["index_mapping"] = index_mapping
        if qids:
            out["qid_list"] = qids
        if task_ids:
            out["task_id_list"] = task_ids
        if p_dids:
            out["p_did_list"] = torch.tensor(p_dids)
        return out


class MBEIRCandidatePoolCollator(_CollatorBase):
    """mbeir_dataset.py:578-610."""

    def __call__(self, batch):
        rows, dids = self._rows(), []
        for inst in batch:
            self._push(rows, inst)
            if inst.get("did") is not None:
                dids.append(inst["did"])
        out = self._finish(rows)
        if dids:
            out["did_list"] = dids
        return out
