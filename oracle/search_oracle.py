"""CPU restatement of the reference's FAISS ``IDMap,Flat`` inner-product search.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/common/mbeir_retriever.py:
  * :69      embeddings loaded and cast fp16 -> fp32
  * :76,:194 ``faiss.normalize_L2`` (rows with zero norm are left untouched)
  * :85-103  ``index_factory(d, "IDMap,Flat", METRIC_INNER_PRODUCT)`` + ``add_with_ids``
  * :231     ``index.search(q, k)`` -> (distances fp32 [Q,k] descending, ids int64 [Q,k]);
             missing results (N < k) are id -1 with distance -3.4028235e38 (faiss convention).
faiss itself (conda ``faiss-gpu``, unpinned, common/faiss_env.yml:10) is not installable
here; the reference holds no golden vectors for it -> parity unpinned by the reference.
"""
from __future__ import annotations

import numpy as np
import torch

FAISS_NEG = np.float32(-3.4028234663852886e38)


def normalize_l2(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
    out = x.copy()
    nz = n[:, 0] > 0
    out[nz] = x[nz] / n[nz]
    return out


def build_index(c16):
    """create_index (mbeir_retriever.py:69,76): fp16 -> fp32, normalize_L2.  Returns the fp32
    normalised matrix the Flat index stores."""
    return torch.from_numpy(normalize_l2(np.asarray(c16, dtype=np.float32)))


def flat_ip_search(q16, c16, ids, k, chunk=65536, dtype=torch.float32, threads=None, index=None):
    """Brute-force cosine top-k with FAISS semantics.  q16 [Q,d], c16 [N,d] (fp16 stored
    embeddings), ids int64 [N].  `index` = build_index(c16) to keep the index build (candidate
    normalisation) out of a timed search, as in the reference where it happens in create_index.
    Returns (scores fp32 [Q,k], ids int64 [Q,k])."""
    if threads:
        torch.set_num_threads(threads)
    qn = torch.from_numpy(normalize_l2(np.asarray(q16, dtype=np.float32))).to(dtype)
    Q = qn.shape[0]
    N = len(c16) if index is None else index.shape[0]
    best_s = torch.full((Q, k), float(FAISS_NEG), dtype=dtype)
    best_i = torch.full((Q, k), -1, dtype=torch.int64)
    for s in range(0, N, chunk):
        if index is None:
            cn = torch.from_numpy(normalize_l2(np.asarray(c16[s:s + chunk], dtype=np.float32))).to(dtype)
        else:
            cn = index[s:s + chunk].to(dtype)
        sc = qn @ cn.t()
        kk = min(k, sc.shape[1])
        ts, ti = torch.topk(sc, kk, dim=1)
        cat_s = torch.cat([best_s, ts], 1)
        cat_i = torch.cat([best_i, ti + s], 1)
        o = torch.argsort(cat_s, dim=1, descending=True, stable=True)[:, :k]
        best_s = torch.gather(cat_s, 1, o)
        best_i = torch.gather(cat_i, 1, o)
    ids_t = torch.from_numpy(np.asarray(ids, dtype=np.int64))
    out_i = torch.where(best_i >= 0, ids_t[best_i.clamp(min=0)], torch.full_like(best_i, -1))
    return best_s.float().numpy(), out_i.numpy()
