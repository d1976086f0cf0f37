"""Exact cosine top-k retrieval on B200 -- the replacement for the reference's FAISS stage
(src/common/mbeir_retriever.py: create_index :34-129, search_index :188-222,
search_index_with_batch :225-232).

``FlatIPIndex`` plays the role of ``faiss.index_factory(d, "IDMap,Flat", METRIC_INNER_PRODUCT)``
after ``normalize_L2`` + ``add_with_ids``: it keeps the *stored fp16 embeddings* plus an fp32
1/||row|| vector and the int64 id table on the device (index build is O(N) and needs no 17 GB
fp32 copy), and ``search`` runs the fused tcgen05 GEMM + top-k kernel followed by the k-way
merge kernel.  With an initialised process group the index is row-sharded over the ranks
(FAISS ``GpuMultipleClonerOptions.shard=True``, mbeir_retriever.py:98-100): one all-gather of
the query embeddings, a local search of every query against the local shard, one all-to-all
of the per-shard [Q,k] partial lists, and a final merge on the rank that owns the query.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import dist_utils as du
from . import lib as L
from . import ops
from .dist_utils import shard_range  # noqa: F401  (re-exported)


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class FlatIPIndex:
    def __init__(self, dim, device="cuda"):
        L.load()
        self.dim = dim
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.UniirError("FlatIPIndex runs on CUDA only (no CPU fallback)")
        self.emb = None      # fp16 [N, d] stored embeddings (this rank's shard)
        self.inv = None      # fp32 [N]   1/||row||
        self.ids = None      # int64 [N]
        self._ws = None

    @property
    def ntotal(self):
        return 0 if self.emb is None else self.emb.shape[0]

    def add_with_ids(self, emb, ids):
        """emb: fp16 [N,d] (numpy or tensor, host or device), ids: int64 [N]."""
        emb = torch.as_tensor(emb)
        ids = torch.as_tensor(np.asarray(ids) if not torch.is_tensor(ids) else ids).to(torch.int64)
        assert emb.dim() == 2 and emb.shape[1] == self.dim, "The dimension of the index does not match the dimension of the embeddings!"
        assert emb.shape[0] == ids.numel()
        emb = emb.to(self.device, torch.float16, non_blocking=True).contiguous()
        ids = ids.to(self.device).contiguous()
        if self.emb is None:
            self.emb, self.ids = emb, ids
        else:
            self.emb = torch.cat([self.emb, emb])
            self.ids = torch.cat([self.ids, ids])
        self.inv = ops.row_inv_norm(self.emb)

    def _local_partials(self, q, q_inv, k):
        splits, ws_bytes = ops.ip_topk_plan(q.shape[0], self.emb.shape[0], k)
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(ws_bytes, device=self.device, dtype=torch.uint8)
        return ops.ip_topk_partial(q, q_inv, self.emb, self.inv, k, workspace=self._ws)

    def search_device(self, q, k):
        """q: fp16 [Q,d] on device -> (scores fp32 [Q,k] descending, ids int64 [Q,k]) on device.
        Missing results (N < k) are id -1 / score -3.4028235e38 like FAISS."""
        assert q.is_cuda and q.dtype == torch.float16 and q.dim() == 2 and q.shape[1] == self.dim
        q = q.contiguous()
        if q.shape[0] == 0:
            return (torch.empty(0, k, device=self.device), torch.empty(0, k, device=self.device, dtype=torch.int64))
        q_inv = ops.row_inv_norm(q)
        if self.ntotal == 0:
            s = torch.full((q.shape[0], k), -3.4028234663852886e38, device=self.device)
            return s, torch.full((q.shape[0], k), -1, device=self.device, dtype=torch.int64)
        ps, pi = self._local_partials(q, q_inv, k)
        return ops.topk_merge(ps, pi, k, id_table=self.ids)

    def search(self, q, k):
        """FAISS-style host API: numpy/tensor queries in, numpy (D, I) out."""
        q = torch.as_tensor(q).to(self.device, torch.float16)
        s, i = self.search_device(q, k)
        return s.cpu().numpy(), i.cpu().numpy()


class ShardedFlatIPIndex:
    """Row-sharded index across the ranks of the default process group (one process per GPU)."""

    def __init__(self, dim, device="cuda", local=None):
        # `local` is injectable so the exchange logic can be exercised on CPU (gloo) in tests
        self.local = local if local is not None else FlatIPIndex(dim, device)
        self.dim = dim
        self.device = torch.device(device) if local is not None else self.local.device
        self.world = dist.get_world_size() if _dist_on() else 1
        self.rank = dist.get_rank() if _dist_on() else 0

    def add_local_shard(self, emb, ids):
        self.local.add_with_ids(emb, ids)

    def search_sharded(self, q_local, k):
        """q_local: this rank's contiguous slice of the query matrix (fp16 [Q_r, d], device).
        Returns the final (scores, ids) for the SAME slice.  Exchanges: one all-gather of the
        queries, one all-to-all of [Q,k] (score, id) partials."""
        if self.world == 1:
            return self.local.search_device(q_local, k)
        W, dev = self.world, self.device
        counts = torch.zeros(W, dtype=torch.int64, device=dev)
        counts[self.rank] = q_local.shape[0]
        dist.all_reduce(counts)
        counts = counts.tolist()
        qmax = max(counts)
        pad = torch.zeros(qmax, self.dim, device=dev, dtype=torch.float16)
        pad[:q_local.shape[0]] = q_local
        allq = du.all_gather_rows(pad)                               # the one query all-gather
        s, i = self.local.search_device(allq, k)                     # [W*qmax, k] vs local shard
        # all-to-all: rank r receives the W partial lists of its own query slice
        rs = du.all_to_all_rows(s.view(W, qmax, k))
        ri = du.all_to_all_rows(i.view(W, qmax, k))
        n = q_local.shape[0]
        ps = rs[:, :n].permute(1, 0, 2).contiguous()                 # [Q_r, W, k]
        pi = ri[:, :n].permute(1, 0, 2).contiguous()
        return self._merge(ps, pi, k)

    def _merge(self, ps, pi, k):
        return ops.topk_merge(ps, pi, k)


class MultiGPUFlatIPIndex:
    """ONE process driving every visible GPU, the way the reference's FAISS stage does
    (`faiss.index_cpu_to_all_gpus(..., co.shard=True)`, mbeir_retriever.py:96-100, 202-206): the candidate rows are
    split into contiguous blocks, one `FlatIPIndex` per device; a search copies the queries to every device, runs the
    fused kernel on each (the launches are asynchronous, so the devices work concurrently), brings the [Q,k] partial
    lists to device 0 and merges them there.  With one visible GPU this is a plain FlatIPIndex."""

    def __init__(self, dim, devices=None):
        L.load()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise L.UniirError("MultiGPUFlatIPIndex needs at least one CUDA device (no CPU fallback)")
        self.dim = dim
        self.devices = [torch.device("cuda", int(i)) if not isinstance(i, torch.device) else i for i in devices]
        self.shards = [FlatIPIndex(dim, d) for d in self.devices]

    @property
    def ntotal(self):
        return sum(s.ntotal for s in self.shards)

    def add_with_ids(self, emb, ids):
        n = len(ids)
        W = len(self.shards)
        for r, sh in enumerate(self.shards):
            lo, hi = shard_range(n, W, r)
            if hi > lo:
                with torch.cuda.device(sh.device):
                    sh.add_with_ids(emb[lo:hi], ids[lo:hi])

    def search(self, q, k):
        q = torch.as_tensor(q)
        parts = []
        for sh in self.shards:                      # launch everywhere first ...
            if sh.ntotal == 0:
                continue
            with torch.cuda.device(sh.device):
                parts.append(sh.search_device(q.to(sh.device, torch.float16, non_blocking=True), k))
        d0 = self.devices[0]
        if not parts:
            return self.shards[0].search(q, k)
        with torch.cuda.device(d0):                 # ... then gather (the copies wait for each device's stream)
            for sh in self.shards:
                torch.cuda.synchronize(sh.device)
            ps = torch.stack([s.to(d0) for s, _ in parts], dim=1).contiguous()      # [Q, W, k]
            pi = torch.stack([i.to(d0) for _, i in parts], dim=1).contiguous()
            s, i = ops.topk_merge(ps, pi, k) if len(parts) > 1 else parts[0]
            return s.cpu().numpy(), i.cpu().numpy()
