"""Pin the FAISS-flat-IP search oracle (reference mbeir_retriever.py:69-103,188-232)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle.search_oracle import FAISS_NEG, build_index, flat_ip_search, normalize_l2


def test_normalize_l2_leaves_zero_rows():
    x = np.array([[3.0, 4.0], [0.0, 0.0]], dtype=np.float32)
    n = normalize_l2(x)
    assert np.allclose(n[0], [0.6, 0.8]) and np.all(n[1] == 0)


def test_matches_float64_brute_force_and_id_map():
    rng = np.random.default_rng(0)
    c = rng.standard_normal((700, 32)).astype(np.float16)
    q = rng.standard_normal((9, 32)).astype(np.float16)
    ids = rng.permutation(10_000)[:700].astype(np.int64)
    D, I = flat_ip_search(q, c, ids, 7, chunk=128)
    qn = q.astype(np.float64); qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    cn = c.astype(np.float64); cn /= np.linalg.norm(cn, axis=1, keepdims=True)
    s = qn @ cn.T
    order = np.argsort(-s, axis=1)[:, :7]
    assert (I == ids[order]).all()
    assert np.allclose(D, np.take_along_axis(s, order, 1), atol=1e-6)
    assert (np.diff(D, axis=1) <= 0).all()
    D2, I2 = flat_ip_search(q, None, ids, 7, chunk=100, index=build_index(c))
    assert (I2 == I).all() and np.allclose(D2, D, atol=1e-6)


def test_fewer_candidates_than_k_pads_like_faiss():
    rng = np.random.default_rng(1)
    c = rng.standard_normal((3, 16)).astype(np.float16)
    q = rng.standard_normal((2, 16)).astype(np.float16)
    D, I = flat_ip_search(q, c, np.array([5, 6, 7]), 5)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == FAISS_NEG).all() and (I[:, :3] >= 5).all()


def test_golden_search_vectors():
    g = np.load(os.path.join(GOLDEN, "search_small.npz"))
    D, I = flat_ip_search(g["q"], g["c"], g["ids"], 10)
    nz = np.ones(len(I), bool); nz[5] = False  # all-zero query: all scores tie at 0
    assert (I[nz] == g["I"][nz]).mean() > 0.995  # fp32 vs the fp64 golden: only exact ties may swap
    assert np.abs(D - g["D"]).max() < 2e-6
    Ds, Is = flat_ip_search(g["q"], g["c"][:6], g["ids"][:6], 10)
    assert (Is[nz] == g["I_small"][nz]).all()
