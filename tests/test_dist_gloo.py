"""world_size-2 gloo tests (CPU) of the hot path's exchange logic: the p-embedding all-gather and its
reduce-scatter backward (clip_sf.py:103), and the sharded-retrieval query all-gather + partial-list
all-to-all + merge (mbeir_retriever.py:204-206 IndexShards semantics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    return [q.get() for _ in range(world)]


def _entry(fn, rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world)))
    finally:
        dist.destroy_process_group()


def _collectives(rank, world):
    from uniir_b200 import dist_utils as du
    x = torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * rank
    g = du.all_gather_rows(x)
    full = torch.arange(world * 3 * 2, dtype=torch.float32).view(world * 3, 2) * (rank + 1)
    rs = du.reduce_scatter_rows(full)
    a2a = du.all_to_all_rows(torch.full((world, 2), float(rank)) + torch.arange(world).view(-1, 1) * 10)
    return g.tolist(), rs.tolist(), a2a.tolist()


def test_backend_aware_collectives():
    res = dict(_run(_collectives))
    for r in (0, 1):
        g, rs, a2a = res[r]
        assert g == [[0, 1], [2, 3], [4, 5], [100, 101], [102, 103], [104, 105]]
        base = torch.arange(12, dtype=torch.float32).view(6, 2) * 3  # sum over ranks of (rank+1)
        assert rs == base[r * 3:(r + 1) * 3].tolist()
        # block i of my output = what rank i sent to me = (i + 10*me)
        assert a2a == [[i + 10.0 * r] * 2 for i in range(2)]


def _loss_exchange(rank, world):
    """Gathered loss + reduce-scatter backward == gradient of the global-mean loss."""
    import torch.nn.functional as F
    from uniir_b200 import dist_utils as du
    from oracle.clip_oracle import contrastive_loss_from_embeddings
    bs, E = 4, 8
    g = torch.Generator().manual_seed(5)
    emb_all = torch.randn(world, 2 * bs, E, generator=g)
    im = {"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]}
    emb = emb_all[rank].clone().requires_grad_(True)
    p_local = F.normalize(emb[1::2], dim=-1)
    all_p = du.all_gather_rows(p_local.detach()).requires_grad_(True)
    q = F.normalize(emb[0::2], dim=-1)
    loss = F.cross_entropy(q @ all_p.t() * 20.0, rank * bs + torch.arange(bs))
    loss.backward()
    d_p = du.reduce_scatter_rows(all_p.grad)            # what torch.distributed.nn.all_gather's backward does
    p_local.backward(d_p)
    # single-process ground truth: W * global-mean loss has exactly these per-rank gradients
    embs = [emb_all[r].clone().requires_grad_(True) for r in range(world)]
    ps = torch.cat([F.normalize(e[1::2], dim=-1) for e in embs])
    total = sum(F.cross_entropy(F.normalize(embs[r][0::2], dim=-1) @ ps.t() * 20.0, r * bs + torch.arange(bs))
                for r in range(world))
    total.backward()
    ref = contrastive_loss_from_embeddings(emb_all[rank], im, torch.tensor(20.0),
                                           all_p_fn=lambda p: torch.cat([F.normalize(emb_all[r][1::2], dim=-1) if r != rank else p
                                                                         for r in range(world)]), rank=rank)
    return (loss.item(), ref["loss"].item(), (emb.grad - embs[rank].grad).abs().max().item())


def test_gathered_loss_and_reduce_scatter_gradient():
    for rank, (loss, ref, gerr) in _run(_loss_exchange):
        assert abs(loss - ref) < 1e-6
        assert gerr < 1e-6


class _CpuLocalIndex:
    """CPU stand-in for FlatIPIndex (oracle search) so the exchange/merge indexing can run under gloo."""

    def __init__(self, emb, ids):
        self.emb, self.ids = emb, ids

    def search_device(self, q, k):
        from oracle.search_oracle import flat_ip_search
        D, I = flat_ip_search(q.numpy(), self.emb, self.ids, k)
        return torch.from_numpy(D), torch.from_numpy(I)


def _sharded_search(rank, world):
    from oracle.search_oracle import flat_ip_search
    from uniir_b200.dist_utils import shard_range
    from uniir_b200.retrieval import ShardedFlatIPIndex
    rng = np.random.default_rng(3)
    N, Q, d, k = 301, 21, 16, 5
    c = rng.standard_normal((N, d)).astype(np.float16)
    q = rng.standard_normal((Q, d)).astype(np.float16)
    ids = (np.arange(N, dtype=np.int64) * 13 + 7)
    lo, hi = shard_range(N, world, rank)
    qlo, qhi = shard_range(Q, world, rank)
    idx = ShardedFlatIPIndex(d, "cpu", local=_CpuLocalIndex(c[lo:hi], ids[lo:hi]))

    def merge(ps, pi, kk):  # torch stand-in for uniir_topk_merge
        Qn = ps.shape[0]
        s, i = ps.reshape(Qn, -1), pi.reshape(Qn, -1)
        o = torch.argsort(s, dim=1, descending=True, stable=True)[:, :kk]
        return s.gather(1, o), i.gather(1, o)
    idx._merge = merge
    D, I = idx.search_sharded(torch.from_numpy(q[qlo:qhi]), k)
    Dr, Ir = flat_ip_search(q[qlo:qhi], c, ids, k)
    return bool((I.numpy() == Ir).all()), float(np.abs(D.numpy() - Dr).max())


def test_sharded_search_exchange_matches_unsharded_oracle():
    for rank, (same, err) in _run(_sharded_search):
        assert same and err < 1e-6


class _FakeEmbedModel:
    """CPU stand-in with the model's embed API, so the embedder's gather logic can run under gloo."""

    def __call__(self, batch, encode_mbeir_batch=False):
        ids = batch["did_list"]
        return torch.tensor([[float(i), float(i) * 0.5, 1.0, -2.0] for i in ids]), ids


def _ragged_gather(rank, world):
    from uniir_b200.mbeir_embedder import _gather_ragged, generate_embeds_and_ids_for_dataset_with_gather
    n = 5 if rank == 0 else 3                                 # the reference's dist.gather breaks on unequal sizes
    x = torch.arange(n * 2, dtype=torch.float32).view(n, 2) + 100 * rank
    parts = _gather_ragged(x, 0)
    ok = [p.shape[0] for p in parts] == [5, 3] and torch.equal(parts[rank], x)
    base = 10 if rank == 0 else 50
    loader = [{"did_list": [base + i for i in range(s, min(n, s + 2))]} for s in range(0, n, 2)]
    emb, ids = generate_embeds_and_ids_for_dataset_with_gather(_FakeEmbedModel(), loader, "cpu")
    if rank == 0:
        want = [10, 11, 12, 13, 14, 50, 51, 52]
        return ok and ids == want and emb.dtype == np.float16 and emb.shape == (8, 4) and emb[:, 0].tolist() == [float(i) for i in want]
    return ok and emb is None and ids == [50, 51, 52]


def test_embedder_gather_handles_ragged_rank_shards():
    for rank, ok in _run(_ragged_gather):
        assert ok, rank


def _shard_handoff(rank, world):
    """embedder -> retriever without the rank-0 funnel: each rank saves its rows, a link-style index points at them,
    each rank loads exactly its own shard file again."""
    from uniir_b200 import mbeir_embedder as me, mbeir_retriever as mr
    from uniir_b200.dist_utils import shard_range
    root = os.environ["UNIIR_TEST_TMP"]
    e_path, i_path = me.embed_paths(root, "embed", "X/", "cand_pool", "union")
    n = 4 if rank == 0 else 3
    emb = (torch.arange(n * 8, dtype=torch.float32).view(n, 8) + 1000 * rank).half()
    ids = torch.arange(n, dtype=torch.int64) + 100 * rank
    me.save_rank_shards(e_path, i_path, emb, ids)
    dist.barrier()
    if rank == 0:
        assert me._concat_rank_shards(e_path, i_path, world) == 7
        mr.write_index_link(os.path.join(root, "u.index"), e_path, i_path)
    dist.barrier()
    full = np.load(e_path)
    lo, hi = shard_range(7, world, rank)
    e, i = mr.read_index_rows(os.path.join(root, "u.index"), lo, hi)
    return bool(full.shape == (7, 8) and np.array_equal(e, emb.numpy()) and np.array_equal(i, ids.numpy()) and
                mr.index_shape(os.path.join(root, "u.index")) == (7, 8))


def test_rank_shard_handoff_embedder_to_retriever(tmp_path):
    os.environ["UNIIR_TEST_TMP"] = str(tmp_path)
    for rank, ok in _run(_shard_handoff):
        assert ok, rank
