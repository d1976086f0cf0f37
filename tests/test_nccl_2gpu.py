"""Two NCCL ranks on two GPUs (skipped with fewer): correctness of the NCCL branches of the hot path's exchanges
against the single-process CPU oracle --
  * the gathered in-batch loss (clip_sf.py:102-103,133-136: all-gather of p, targets rank*bs + arange) and its
    backward (reduce-scatter of d all_p) + the DDP-style gradient all-reduce: per-rank loss, and the averaged
    gradient of every parameter, equal those of the oracle on the concatenated global batch;
  * sharded retrieval (mbeir_retriever.py:204-206 IndexShards semantics): query all-gather, local fused search,
    all-to-all of the partial lists, merge == un-sharded FAISS-semantics oracle.
Run with `gpurun --gpus 2 -- python -m pytest tests/test_nccl_2gpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BS = 8   # pairs per rank


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_batch(dims, world, seed=31):
    from uniir_b200.preprocess import synthetic_clip_tokens
    g = torch.Generator().manual_seed(seed)
    n = 2 * BS * world
    mt = torch.ones(n, dtype=torch.long); mv = torch.ones(n, dtype=torch.long)
    mt[1::5] = 0; mv[2::5] = 0
    return dict(txt_batched=synthetic_clip_tokens(n, g, dims.context, dims.vocab),
                image_batched=torch.randn(n, 3, dims.image_size, dims.image_size, generator=g),
                txt_mask_batched=mt, image_mask_batched=mv)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from types import SimpleNamespace as NS
        from oracle.clip_oracle import CLIPScoreFusionOracle
        from oracle.search_oracle import flat_ip_search
        from uniir_b200.clip_sf import CLIPScoreFusion
        from uniir_b200.dist_utils import shard_range
        from uniir_b200.engine import allreduce_grads
        from uniir_b200.retrieval import ShardedFlatIPIndex
        assert dist.get_backend() == "nccl"
        # ---- gathered loss + reduce-scatter + gradient all-reduce
        torch.manual_seed(0)
        orc = CLIPScoreFusionOracle("tiny-test")
        cfg = NS(model=NS(gather_embeddings=True), data_config=NS(in_batch_neg_num=0))
        mine = CLIPScoreFusion("tiny-test", device=f"cuda:{rank}", config=cfg, micro_batch=8)
        mine.load_state_dict(orc.state_dict(), strict=True)
        gb = _global_batch(orc.clip_model.dims, world)
        lo, hi = 2 * BS * rank, 2 * BS * (rank + 1)
        local = {k: v[lo:hi].cuda() for k, v in gb.items()}
        local["index_mapping"] = {"query": [[2 * i] for i in range(BS)], "pos_cand": [[2 * i + 1] for i in range(BS)]}
        loss, acc = mine._loss_forward(local)
        mine._loss_backward(local, None)
        allreduce_grads(mine.clip_model)
        torch.cuda.synchronize()
        # oracle: the whole global batch in one process; rank r's loss = mean CE over its rows of the global logits
        full = dict(gb, index_mapping={"query": [[2 * i] for i in range(BS * world)],
                                       "pos_cand": [[2 * i + 1] for i in range(BS * world)]})
        emb = orc.encode_multimodal_input(full["txt_batched"], full["image_batched"], full["txt_mask_batched"],
                                          full["image_mask_batched"])
        q = torch.nn.functional.normalize(emb[0::2], dim=-1); p = torch.nn.functional.normalize(emb[1::2], dim=-1)
        logits = q @ p.t() * orc.clip_model.logit_scale.exp()
        tgt = torch.arange(BS * world)
        per_rank = [torch.nn.functional.cross_entropy(logits[r * BS:(r + 1) * BS], tgt[r * BS:(r + 1) * BS]) for r in range(world)]
        (sum(per_rank) / world).backward()
        res = {"loss": loss.item(), "loss_ref": per_rank[rank].item(), "acc": acc.item(),
               "acc_ref": (logits[rank * BS:(rank + 1) * BS].argmax(1) == tgt[rank * BS:(rank + 1) * BS]).float().mean().item()}
        worst = 1.0
        for n_, p_o in orc.named_parameters():
            a, c = p_o.grad.flatten().double(), mine.clip_model.g(n_[len("clip_model."):]).cpu().flatten().double()
            if a.numel() > 1:
                worst = min(worst, (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item())
            else:
                res["dls"] = (a.item(), c.item())
        res["grad_cos_min"] = worst
        # ---- sharded retrieval over NCCL
        rng = np.random.default_rng(3)
        N, Q, d, k = 5001, 301, 64, 10
        c = rng.standard_normal((N, d)).astype(np.float16)
        qq = rng.standard_normal((Q, d)).astype(np.float16)
        ids = np.arange(N, dtype=np.int64) * 13 + 7
        clo, chi = shard_range(N, world, rank)
        qlo, qhi = shard_range(Q, world, rank)
        idx = ShardedFlatIPIndex(d, f"cuda:{rank}")
        idx.add_local_shard(c[clo:chi], ids[clo:chi])
        D, I = idx.search_sharded(torch.from_numpy(qq[qlo:qhi]).cuda(), k)
        Dr, Ir = flat_ip_search(qq[qlo:qhi], c, ids, k)
        res["search_same"] = float((I.cpu().numpy() == Ir).mean())
        res["search_err"] = float(np.abs(D.cpu().numpy() - Dr).max())
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_gathered_loss_gradients_and_sharded_search(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    for r in range(world):
        res = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert abs(res["loss"] - res["loss_ref"]) < 2e-2 * max(1.0, abs(res["loss_ref"])), res
        assert abs(res["acc"] - res["acc_ref"]) < 1.0 / BS + 1e-6, res      # bf16 towers: one near-tie may flip
        assert res["grad_cos_min"] > 0.995, res
        assert abs(res["dls"][0] - res["dls"][1]) < 2e-2 * max(1.0, abs(res["dls"][0])), res
        assert res["search_same"] == 1.0 and res["search_err"] < 1e-5, res
