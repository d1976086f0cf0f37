"""Process-group helpers for the hot path's three exchanges:

  * all_gather_rows      -- p-embedding all-gather of the in-batch loss (reference
                            clip_sf.py:103, torch.distributed.nn.all_gather forward) and the
                            query all-gather of sharded retrieval;
  * reduce_scatter_rows  -- its backward (sum of every rank's gradient for my rows);
  * all_to_all_rows      -- exchange of per-shard [Q,k] partial top-k lists.
NCCL uses the native collectives (NVLink/NVSwitch); gloo (CPU tests, world_size 2) has no
reduce_scatter / all_to_all, so those fall back to all_reduce / all_gather + slicing.
Also ``init_distributed_mode`` (env:// init like reference common/dist_utils.py:62-91) and the
contiguous row partition of ``ContiguousDistributedSampler`` (dist_utils.py:94-115).
"""
from __future__ import annotations

import datetime
import math
import os

import torch
import torch.distributed as dist


def is_on():
    return dist.is_available() and dist.is_initialized()


def world():
    return dist.get_world_size() if is_on() else 1


def rank():
    return dist.get_rank() if is_on() else 0


def _native(group=None):
    return is_on() and dist.get_backend(group) == "nccl"


def all_gather_rows(x, out=None):
    """x [n, ...] (same n on every rank) -> [W*n, ...] in rank order."""
    W = world()
    if W == 1:
        return x if out is None else out.copy_(x)
    if out is None:
        out = x.new_empty((W * x.shape[0],) + tuple(x.shape[1:]))
    x = x.contiguous()
    if _native():
        dist.all_gather_into_tensor(out, x)
    else:
        parts = [torch.empty_like(x) for _ in range(W)]
        dist.all_gather(parts, x)
        out.copy_(torch.cat(parts, 0))
    return out


def reduce_scatter_rows(x, out=None):
    """x [W*n, ...] on every rank -> sum over ranks of rows [r*n, (r+1)*n)."""
    W, r = world(), rank()
    n = x.shape[0] // W
    if W == 1:
        return x if out is None else out.copy_(x)
    if out is None:
        out = x.new_empty((n,) + tuple(x.shape[1:]))
    if _native():
        dist.reduce_scatter_tensor(out, x.contiguous(), op=dist.ReduceOp.SUM)
    else:
        tmp = x.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        out.copy_(tmp[r * n:(r + 1) * n])
    return out


def all_to_all_rows(x):
    """x [W, n, ...]: block j goes to rank j; returns [W, n, ...] with block i = what rank i sent me."""
    W, r = world(), rank()
    if W == 1:
        return x
    x = x.contiguous()
    out = torch.empty_like(x)
    if _native():
        dist.all_to_all_single(out.view(-1), x.view(-1))
    else:
        parts = [torch.empty_like(x) for _ in range(W)]
        dist.all_gather(parts, x)
        for i in range(W):
            out[i] = parts[i][r]
    return out


def shard_range(n, world_size, r):
    """Rows of rank r: [r*ceil(n/W), min((r+1)*ceil(n/W), n)) (reference dist_utils.py:100-109)."""
    per = math.ceil(n / world_size) if world_size > 0 else n
    lo = min(n, r * per)
    return lo, min(n, lo + per)


class ContiguousDistributedSampler(torch.utils.data.Sampler):
    """Same contract as the reference sampler (common/dist_utils.py:94-115)."""

    def __init__(self, dataset, num_replicas=None, rank=None):
        self.dataset = dataset
        self.num_replicas = num_replicas if num_replicas is not None else world()
        self.rank = rank if rank is not None else globals()["rank"]()
        self.epoch = 0
        self.num_samples_per_replica = math.ceil(len(self.dataset) / self.num_replicas)
        self.total_size = self.num_samples_per_replica * self.num_replicas

    def __iter__(self):
        lo, hi = shard_range(len(self.dataset), self.num_replicas, self.rank)
        return iter(range(lo, hi))

    def __len__(self):
        return self.num_samples_per_replica

    def set_epoch(self, epoch):
        self.epoch = epoch


def init_distributed_mode(args):
    """env:// initialisation with the reference's field names (dist_utils.py:62-91):
    sets args.rank / world_size / gpu / distributed; 60-minute timeout."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    elif "SLURM_PROCID" in os.environ:
        args.rank = int(os.environ["SLURM_PROCID"])
        args.gpu = args.rank % max(1, torch.cuda.device_count())
        args.world_size = int(os.environ.get("SLURM_NTASKS", 1))
    else:
        args.distributed = False
        args.rank, args.world_size, args.gpu = 0, 1, 0
        return
    args.distributed = True
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(args.gpu)
    dist.init_process_group(backend=backend, init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank,
                            timeout=datetime.timedelta(minutes=60))
    dist.barrier()
