"""GPU parity of the fused GEMM + top-k retrieval path against the FAISS-semantics oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _index(c, ids):
    from uniir_b200.retrieval import FlatIPIndex
    idx = FlatIPIndex(c.shape[1], "cuda")
    idx.add_with_ids(c, ids)
    return idx


def test_golden_vectors_and_padding():
    g = np.load(os.path.join(GOLDEN, "search_small.npz"))
    D, I = _index(g["c"], g["ids"]).search(g["q"], 10)
    nz = np.ones(len(I), bool); nz[5] = False     # query 5 is all-zero: every score ties at 0, order unspecified
    assert (I[nz] == g["I"][nz]).mean() > 0.995 and np.abs(D - g["D"]).max() < 2e-6
    assert (D[5] == 0).all()                      # zero rows score 0 like normalize_L2 (no NaN)
    assert (np.diff(D, axis=1) <= 0).all() and I.dtype == np.int64 and D.dtype == np.float32
    Ds, Is = _index(g["c"][:6], g["ids"][:6]).search(g["q"], 10)                # N < k -> FAISS -1 padding
    assert (Is[nz] == g["I_small"][nz]).all() and (Ds[:, 6:] < -3e38).all() and (Is[:, 6:] == -1).all()


@pytest.mark.parametrize("Q,N,d,k", [(1, 1000, 64, 10), (100, 1000, 64, 10), (300, 5000, 128, 10), (130, 40, 128, 50),
                                      (1000, 100_000, 768, 50), (257, 30_001, 768, 64), (129, 513, 8, 1)])
def test_against_oracle(Q, N, d, k):
    from oracle.search_oracle import flat_ip_search
    rng = np.random.default_rng(Q * 7 + N)
    q = rng.standard_normal((Q, d)).astype(np.float16); c = rng.standard_normal((N, d)).astype(np.float16)
    ids = rng.permutation(N * 3)[:N].astype(np.int64)
    D, I = _index(c, ids).search(q, k)
    Dr, Ir = flat_ip_search(q, c, ids, k)
    kk = min(k, N)
    assert np.abs(D[:, :kk] - Dr[:, :kk]).max() < 2e-6        # fp32 tolerance on cosine scores
    diff = I[:, :kk] != Ir[:, :kk]
    # indices must be identical wherever neighbouring scores are separated by more than fp32 noise
    gap = np.minimum(np.abs(np.diff(Dr[:, :kk], axis=1, prepend=2)), np.abs(np.diff(Dr[:, :kk], axis=1, append=-2)))
    assert not (diff & (gap > 1e-5)).any()
    assert diff.mean() < 2e-3


def test_planted_top50_at_scale_is_recovered_in_order():
    """plant, per query, 50 candidates at cosine 0.90, 0.89 .. 0.41 -> exact ids in exact order."""
    torch.manual_seed(0)
    Q, N, d, k = 2048, 1_000_000, 768, 50
    c = torch.randn(N, d, device="cuda").half()
    q = torch.randn(Q, d, device="cuda")
    qn = torch.nn.functional.normalize(q, dim=-1)
    rows = torch.randperm(N, device="cuda")[:Q * k].view(Q, k)
    cos = (0.90 - 0.01 * torch.arange(k, device="cuda")).view(1, k, 1)
    noise = torch.randn(Q, k, d, device="cuda")
    noise = torch.nn.functional.normalize(noise - (noise * qn[:, None]).sum(-1, keepdim=True) * qn[:, None], dim=-1)
    c[rows.view(-1)] = ((cos * qn[:, None] + (1 - cos ** 2).sqrt() * noise) * 3.0).view(-1, d).half()
    ids = torch.arange(N, device="cuda", dtype=torch.int64) * 3 + 1
    idx = _index(c, ids)
    D, I = idx.search_device(q.half(), k)
    assert (I == ids[rows]).all()
    assert (D[:, :-1] >= D[:, 1:]).all() and (D[:, 0] - 0.90).abs().max() < 5e-3


def test_search_properties_large():
    """size-independent properties at a large shape: sorted, unique, self-retrieval of stored rows."""
    torch.manual_seed(1)
    N, d, k = 500_000, 768, 50
    c = torch.randn(N, d, device="cuda").half()
    idx = _index(c, torch.arange(N, dtype=torch.int64))
    sel = torch.randint(0, N, (4096,), device="cuda")
    D, I = idx.search_device(c[sel], k)
    assert (I[:, 0] == sel).all() and (D[:, 0] - 1).abs().max() < 1e-5      # a stored row retrieves itself at cosine 1
    assert (D[:, :-1] >= D[:, 1:]).all()
    assert all(len(set(r.tolist())) == k for r in I[:64])


def test_merge_kernel_streams_more_lists_than_one_buffer():
    from uniir_b200 import ops
    torch.manual_seed(2)
    Q, L, k = 33, 40, 50     # L*k = 2000 > the 512-entry sort buffer
    s = torch.randn(Q, L, k, device="cuda").sort(dim=2, descending=True).values.contiguous()
    i = torch.randint(0, 10**9, (Q, L, k), device="cuda", dtype=torch.int64)
    i[0, 0, :] = -1  # padding entries are skipped
    D, I = ops.topk_merge(s, i, k)
    flat_s = s.view(Q, -1).clone(); flat_s[0, :k] = -float("inf")
    ref = flat_s.sort(dim=1, descending=True).values[:, :k]
    assert torch.equal(D, ref)


def test_create_index_and_search_index_cli_functions(tmp_path):
    """embed .npy pair -> create_index -> search_index, same signatures/returns as the reference."""
    from oracle.search_oracle import flat_ip_search
    from uniir_b200 import config as cfg
    from uniir_b200.mbeir_embedder import embed_paths, save_embeddings
    from uniir_b200.mbeir_retriever import create_index, search_index
    rng = np.random.default_rng(5)
    c = rng.standard_normal((4000, 64)).astype(np.float16); q = rng.standard_normal((50, 64)).astype(np.float16)
    ids = np.arange(4000, dtype=np.int64) + 9 * 10_000_000
    conf = cfg.create({"uniir_dir": str(tmp_path), "experiment": {"path_suffix": "CLIP_SF/Large/Instruct/InBatch/"},
                       "index_config": {"embed_dir_name": "embed", "index_dir_name": "index",
                                        "faiss_config": {"idx_type": "Flat", "dim": 64, "metric": "METRIC_INNER_PRODUCT"},
                                        "cand_pools_config": {"enable_idx": True, "cand_pools_name_to_idx": ["UNION"]}}})
    e, i = embed_paths(str(tmp_path), "embed", conf.experiment.path_suffix, "cand_pool", "union")
    save_embeddings(e, i, c, ids)
    qe, qi = embed_paths(str(tmp_path), "embed", conf.experiment.path_suffix, "test", "q")
    save_embeddings(qe, qi, q, np.arange(50))
    create_index(conf)
    ipath = os.path.join(str(tmp_path), "index", conf.experiment.path_suffix, "cand_pool", "mbeir_union_cand_pool.index")
    D, I = search_index(qe, ipath, batch_size=10, num_cand_to_retrieve=10)
    Dr, Ir = flat_ip_search(q, c, ids, 10)
    assert D.shape == (50, 10) and D.dtype == np.float32 and I.dtype == np.int64
    assert (I == Ir).mean() > 0.995 and np.abs(D - Dr).max() < 2e-6
