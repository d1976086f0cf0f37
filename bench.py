#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native UniIR hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Default workload = BASELINE.json configs[1]: CLIP_SF ViT-L/14 in-batch contrastive training
step, bf16, GLOBAL batch 4096 pairs (strong scaling: 4096/N pairs per GPU), synthetic
224x224 images + 77-token text, random-init weights.  One "step" = gradient-cached forward +
loss + backward + gradient all-reduce + AdamW over the whole global batch.
`--workload search` = configs[4]: 190k x 5.6M x 768 fp16 top-50 (also attached to the default
line under "extra" at N=1 unless --no-extra).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "ViT-L/14"
GLOBAL_BATCH = 4096
FLOPS_PER_PAIR = 1.052e12          # 2 items x 3 x 175.33 GFLOP (algorithmic, fwd+bwd)
SEARCH_Q, SEARCH_N, SEARCH_D, SEARCH_K = 190_000, 5_600_000, 768, 50


def load_peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json (copy bandwidth, cuBLAS bf16 8192^3 burst =
    best of 10 for a kernel timed alone, sustained = back to back for 4 s for a kernel timed inside a long step);
    else the fallback of /opt/skills/guides/B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "B200_PROFILING.md fallback"}


def _fp16_peaks():
    """The kernel computes in fp16, MEASURED_PEAKS.json holds bf16: cuBLAS fp16 8192^3 measured on this pool with
    scripts/peak_fp16.py (profiles/peak_fp16_r02.json) is reported beside the roofline for reference."""
    p = os.path.join(ROOT, "profiles", "peak_fp16_r02.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def gemm_roofline(sink_timed, iso_ms, iso_fl, peaks, step_ms, step_alg_tflops, traffic):
    """Roofline record of the tcgen05 GEMM family of one step.
      achieved / frac            : sum(algorithmic FLOPs) / sum(durations) of every GEMM launch, each timed IN ITS PLACE
                                   in the step (CUDA events around the launch, real predecessors / L2 state) -> against
                                   the SUSTAINED peak (kernel inside a long step);
      isolated_*                 : the same launches timed alone, back to back on warm operands -> against the BURST peak;
      frac_step_algorithmic      : the whole step's algorithmic FLOP/s against the sustained peak (what the step, not
                                   the kernel, achieves: re-forward, attention, LayerNorm, optimizer are all in it)."""
    fl = sum(2.0 * c["M"] * c["N"] * c["K"] for c in sink_timed)
    ms = sum(c["events"][0].elapsed_time(c["events"][1]) for c in sink_timed)
    ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    iso = iso_fl / (iso_ms * 1e-3) / 1e12 if iso_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "uniir::gemm_epi_kernel (tcgen05.mma kind::f16, TMA-fed, TMEM accum)",
            "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
            "peak_src": f"{peaks['src']}: sustained bf16 (launches timed in place inside the step)",
            "timing": "CUDA events around every GEMM launch of one eagerly executed micro-batch pass (embedding pass + "
                      "re-forward + backward; both towers on one stream for this pass), scaled by the micro-batches per step",
            "isolated_achieved": iso, "isolated_peak": peaks["bf16_burst"], "isolated_frac": iso / peaks["bf16_burst"],
            "isolated_note": "same launches timed alone (5 back-to-back launches on warm operands) against the burst peak",
            "gemm_share_of_step": None, "traffic": traffic, "step_algorithmic_tflops": step_alg_tflops,
            "frac_step_algorithmic": step_alg_tflops / peaks["bf16_sustained"]}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch(dims, pairs, seed, device, pinned=False, mask_mix="ones"):
    """SURVEY.md 8d synthetic inputs: randn images (CLIP-normalised images are ~N(0,1)), token ids
    [SOT, U x L, EOT, 0..], all-ones masks, query = even rows / pos_cand = odd rows.
    mask_mix="mbeir": the second SURVEY 8d configuration, 1/3 text-only, 1/3 image-only, 1/3 both."""
    from uniir_b200.preprocess import synthetic_clip_tokens
    n = 2 * pairs
    g = torch.Generator().manual_seed(seed)
    ids = synthetic_clip_tokens(n, g, dims.context, dims.vocab)
    if device.type == "cuda" and not pinned:
        gd = torch.Generator(device=device).manual_seed(seed)
        img = torch.randn(n, 3, dims.image_size, dims.image_size, generator=gd, device=device)
        ids = ids.to(device)
        ones = torch.ones(n, dtype=torch.long, device=device)
    else:
        img = torch.empty(n, 3, dims.image_size, dims.image_size, pin_memory=pinned)
        chunk = 256
        for s in range(0, n, chunk):
            img[s:s + chunk].normal_(generator=g)
        ones = torch.ones(n, dtype=torch.long)
        if pinned:
            ids, ones = ids.pin_memory(), ones.pin_memory()
    mt, mv = ones, ones.clone()
    if mask_mix == "mbeir":
        kind = torch.arange(n, device=ones.device) % 3          # 0: text-only, 1: image-only, 2: both
        mt, mv = (kind != 1).long(), (kind != 0).long()
        if pinned:
            mt, mv = mt.pin_memory(), mv.pin_memory()
    return dict(txt_batched=ids, image_batched=img, txt_mask_batched=mt, image_mask_batched=mv,
                index_mapping={"query": [[2 * i] for i in range(pairs)], "pos_cand": [[2 * i + 1] for i in range(pairs)]})


def timed(fn, steps, world):
    """barrier + sync on both sides, CUDA events on the launch stream, max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = t.item()
    return ms / steps


# ----------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the oracle restatement of the reference, on host cores)
# ----------------------------------------------------------------------------------------
def _cpu_threads():
    """Threads the CPU arm actually uses: all cores up to 32 (beyond that torch's intra-op pool
    slows these batch sizes down -- measured 9x slower with 128 threads than with 8)."""
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    return cores


def cpu_train_reference(steps, warmup, pairs=4):
    from oracle.clip_oracle import CLIPScoreFusionOracle
    cores = _cpu_threads()
    torch.manual_seed(2023)
    model = CLIPScoreFusionOracle(MODEL)
    decay = [p for n, p in model.named_parameters() if p.ndim >= 2 and "ln" not in n and "bias" not in n]
    nodecay = [p for n, p in model.named_parameters() if not (p.ndim >= 2 and "ln" not in n and "bias" not in n)]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.2}, {"params": nodecay, "weight_decay": 0.0}],
                            lr=1e-5, betas=(0.9, 0.98), eps=1e-6)
    batch = synth_batch(model.clip_model.dims, pairs, 2023, torch.device("cpu"))

    def step():
        out = model(batch)
        out["loss"].backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"oracle CLIP_SF {MODEL} fp32, {pairs} pairs/step (of the {GLOBAL_BATCH}-pair batch), "
                      f"fwd+bwd+AdamW, {steps} timed steps"}, dt


def cpu_search_reference(steps, warmup, queries=1024, cands=500_000):
    import numpy as np
    from oracle.search_oracle import build_index, flat_ip_search
    cores = _cpu_threads()
    g = torch.Generator().manual_seed(2023)
    q = torch.randn(queries, SEARCH_D, generator=g).half().numpy()
    c = torch.randn(cands, SEARCH_D, generator=g).half().numpy()
    ids = np.arange(cands, dtype=np.int64)
    index = build_index(c)  # create_index happens once, outside the timed search (as in the reference)
    for _ in range(warmup):
        flat_ip_search(q[:32], None, ids, SEARCH_K, index=index[:65536])
    t0 = time.perf_counter()
    for _ in range(steps):
        flat_ip_search(q, None, ids, SEARCH_K, index=index)
    dt = (time.perf_counter() - t0) / steps
    # exactly linear in N and Q: scale the sample to the 5.6M-candidate pool
    qps = queries / (dt * SEARCH_N / cands)
    return {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"oracle fp32 search (normalise queries, matmul, topk) on a prebuilt index, {queries} queries x {cands} candidates, scaled "
                      f"linearly to N={SEARCH_N}"}, dt


def cpu_embed_reference(steps, warmup, items=8):
    """configs[3] on the host cores: the oracle's embed path (both towers + score fusion, fp32) on `items` items."""
    from oracle.clip_oracle import CLIPScoreFusionOracle
    cores = _cpu_threads()
    torch.manual_seed(2023)
    model = CLIPScoreFusionOracle(MODEL).eval()
    b = synth_batch(model.clip_model.dims, items // 2, 2023, torch.device("cpu"))
    b.pop("index_mapping"); b["did_list"] = list(range(items))
    b["txt_batched"] = b["txt_batched"].long()

    def step():
        with torch.no_grad():
            emb, _ = model(b, encode_mbeir_batch=True)
            return emb.half()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": items / dt, "unit": "items/s", "cores": cores, "kind": "port",
            "sample": f"oracle CLIP_SF {MODEL} fp32 encode_mbeir_batch, {items} items/step (of the 5.6M-candidate pool), {steps} timed steps"}, dt


def run_reference(args, rank):
    if rank != 0:
        return
    if args.workload == "embed":
        cb, dt = cpu_embed_reference(args.steps, args.warmup)
        metric, unit = "embedded_items_per_sec", "items/s"
        cfg = {"workload": f"embedding extraction, CLIP_SF {MODEL}, batch 2048/GPU (BASELINE configs[3] shape)",
               "sampled": "8 items per step on the host cores"}
    elif args.workload == "search":
        cb, dt = cpu_search_reference(args.steps, args.warmup)
        metric, unit = "queries_per_sec_top50", "queries/s"
        cfg = {"workload": f"dense retrieval {SEARCH_Q} queries x {SEARCH_N} candidates, d={SEARCH_D} fp16, top-{SEARCH_K} (BASELINE configs[4])",
               "sampled": "1024 queries x 500000 candidates per step, scaled linearly"}
    elif args.workload == "blip":
        cb, dt = cpu_blip_reference(args.steps, args.warmup, sf=args.blip_model == "sf")
        metric, unit = "contrastive_pairs_per_sec", "pairs/s"
        cfg = {"workload": ("BLIP_SF ViT-L/16 + text-mode BERT-base momentum-distillation training step (SURVEY 8f-3; BASELINE configs[2] shape)"
                            if args.blip_model == "sf" else
                            "BLIP_FF ViT-L/16 + BERT-base momentum-distillation training step (BASELINE configs[2])"),
               "global_batch": BLIP_GLOBAL_BATCH, "sampled": "2 pairs per step on the host cores"}
    else:
        cb, dt = cpu_train_reference(args.steps, args.warmup)
        metric, unit = "contrastive_pairs_per_sec", "pairs/s"
        cfg = {"workload": f"CLIP_SF {MODEL} in-batch contrastive training step (BASELINE configs[1])",
               "model": f"CLIP_SF {MODEL}", "global_batch": GLOBAL_BATCH, "sampled": "4 pairs per step on the host cores"}
    print(json.dumps({"impl": "reference", "metric": metric, "value": cb["value"], "unit": unit,
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------
# our arm: training step
# ----------------------------------------------------------------------------------------
def run_train(args, rank, world, dev, peaks):
    from uniir_b200 import lib, ops
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.engine import FlatAdamW, train_step

    class Cfg:  # the two fields the reference constructor reads (clip_sf.py:28-30)
        class model: gather_embeddings = world > 1
        class data_config: in_batch_neg_num = 0

    GB = GLOBAL_BATCH if args.pairs_per_gpu is None else args.pairs_per_gpu * world   # BASELINE configs[1] unless emulating a rank
    pairs = GB // world
    model = CLIPScoreFusion(MODEL, device=dev, config=Cfg, micro_batch=args.micro_batch)
    opt = FlatAdamW(model.clip_model, lr=1e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    batch = synth_batch(model.dims, pairs, 2023 + rank, dev, mask_mix=args.mask_mix)
    model.skip_masked_towers = not args.no_tower_skip
    if args.retain is not None:
        model.retain_micro_batches = args.retain
    state = {}

    def step():
        state["loss"], state["acc"] = train_step(model, batch, opt)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = lib.launch_count() + ops.graph_launches()
    ms = timed(step, args.steps, world)
    launches = lib.launch_count() + ops.graph_launches() - l0   # direct launches + kernels replayed from CUDA graphs
    clocks = sampler.stop()
    value = GB / (ms / 1e3)

    # ---- end-to-end through the public API with HOST (pinned) buffers -------------------
    hb = {k: (torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) if torch.is_tensor(v) else v)
          for k, v in batch.items()}
    h2d = sum(v.numel() * v.element_size() for v in hb.values() if torch.is_tensor(v))

    def e2e_step():
        # the public call with the pinned HOST batch: the model uploads it (images in micro-batch chunks on a copy
        # stream, overlapped with the towers; everything else up front) -- every byte of `h2d` crosses PCIe each step
        loss, acc = train_step(model, hb, opt)
        state["e2e_loss"] = loss.item()  # device -> host read of the step's result

    e2e_step()
    ms_e2e = timed(e2e_step, max(1, min(2, args.steps)), world)
    e2e = {"value": GB / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e}

    # ---- roofline of the dominant kernel (tcgen05 GEMM), timed live with CUDA events -----
    # The step replays CUDA graphs, which cannot be event-timed per kernel.  So one micro-batch pass (embedding pass
    # + re-forward + backward: every kernel of the step in its real order) runs eagerly with a pair of CUDA events
    # around each GEMM launch; the same launches are additionally timed alone for the burst-peak comparison.
    from collections import Counter
    enc = model._enc[(model._train_mb, True)]
    sink = []
    enc.serial_towers = True          # one stream: a timed GEMM does not share the SMs with the other tower's kernels
    ops.set_gemm_profiler(sink, events=True)
    enc._body_embed()
    enc._body_fwdbwd()
    ops.set_gemm_profiler(None)
    enc.serial_towers = False
    torch.cuda.synchronize()
    for gbuf in model.clip_model.grad.values():
        gbuf.zero_()
    keyf = lambda c: tuple(sorted((k, str(v)) for k, v in c.items() if k != "events"))
    uniq, counts = {}, Counter()
    for c in sink:
        uniq[keyf(c)] = c
        counts[keyf(c)] += 1
    n_mb = (2 * pairs) // model._train_mb
    iso_ms = iso_fl = 0.0
    for k_, c in uniq.items():
        t = ops.time_gemm_config(c, iters=5, device=dev)
        iso_ms += t * counts[k_]
        iso_fl += 2.0 * c["M"] * c["N"] * c["K"] * counts[k_]
    roof = gemm_roofline(sink, iso_ms, iso_fl, peaks, ms, FLOPS_PER_PAIR * GB / world / (ms / 1e3) / 1e12, _ncu_traffic("gemm"))
    in_place_ms = sum(c["events"][0].elapsed_time(c["events"][1]) for c in sink)
    n_fwd_only = getattr(model, "_n_retained", 0)
    # per step: every micro-batch runs the embedding pass; the re-forward + backward pass runs for all of them (the
    # retained ones skip only the re-forward, which this estimate ignores: it is an upper bound of the GEMM share)
    roof.update(launches=len(sink) * n_mb, distinct_configs=len(uniq), gemm_share_of_step=in_place_ms * n_mb / ms if ms else None)
    out = {"metric": "contrastive_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"CLIP_SF {MODEL} in-batch contrastive training step (BASELINE configs[1])"
                                  + ("" if args.pairs_per_gpu is None else f" -- DIAGNOSTIC: {pairs} pairs per GPU, not the 4096-pair batch"),
                      "model": f"CLIP_SF {MODEL}", "global_batch": GB, "pairs_per_gpu": pairs,
                      "micro_batch_items": model._train_mb if hasattr(model, "_train_mb") else args.micro_batch,
                      "seq_len": "257 vision tokens + 77 text tokens per item, 2 items per pair",
                      "parallelism": f"dp{world}",
                      "grad_cache": f"embedding-level: extra forward (not credited) for {n_mb - getattr(model, '_n_retained', 0)} of "
                                    f"{n_mb} micro-batches per GPU; {getattr(model, '_n_retained', 0)} keep their activations in HBM",
                      "l2": "inputs+weights+activations per step >> 126 MB L2 (no flush needed)",
                      "masks": ("all ones (both modalities present)" if args.mask_mix == "ones" else
                                "1/3 text-only, 1/3 image-only, 1/3 both; masked towers "
                                + ("encoded and multiplied by 0 (reference behaviour)" if args.no_tower_skip else "skipped"))},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
           "loss": float(state["loss"].item()), "accuracy": float(state["acc"].item())}
    del model, opt, batch, hb
    torch.cuda.empty_cache()
    return out


def _ncu_traffic(key, rows=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu captures
    (profiles/roofline_traffic.json, written by scripts/make_traffic_json.py from the `ncu` CSVs beside it).  The
    retrieval entry is keyed by the candidate rows of the launch (5.6M / 2.8M / 1.4M / 0.7M = the 1/2/4/8-GPU shards);
    a size that was not captured reports null rather than another size's number."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        v = json.load(open(p)).get(key)
    except Exception:
        return None
    if isinstance(v, dict):
        return v.get(str(rows))
    return v


# ----------------------------------------------------------------------------------------
# our arm: dense retrieval
# ----------------------------------------------------------------------------------------
def run_search(args, rank, world, dev, peaks, steps=None, warmup=None):
    from uniir_b200 import lib, ops
    from uniir_b200.retrieval import ShardedFlatIPIndex, shard_range
    steps = steps or args.steps
    warmup = warmup if warmup is not None else args.warmup
    Q, N, d, k = args.search_q, args.search_n, SEARCH_D, SEARCH_K
    lo, hi = shard_range(N, world, rank)
    qlo, qhi = shard_range(Q, world, rank)
    gd = torch.Generator(device=dev).manual_seed(2023 + rank)
    cand = torch.empty(hi - lo, d, device=dev, dtype=torch.float16)
    for s in range(0, hi - lo, 1 << 20):
        e = min(hi - lo, s + (1 << 20))
        cand[s:e] = torch.randn(e - s, d, generator=gd, device=dev).half()
    q = torch.randn(qhi - qlo, d, generator=gd, device=dev).half()
    index = ShardedFlatIPIndex(d, dev)
    index.add_local_shard(cand, torch.arange(lo, hi, dtype=torch.int64, device=dev))
    state = {}

    def step():
        state["out"] = index.search_sharded(q, k)

    for _ in range(warmup):
        step()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = lib.launch_count()
    ms = timed(step, steps, world)
    launches = lib.launch_count() - l0
    clocks = sampler.stop()
    value = Q / (ms / 1e3)
    # e2e: host fp16 queries in (pinned), host (D, I) out
    qh = q.cpu().pin_memory()

    def e2e_step():
        qd = qh.to(dev, non_blocking=True)
        s_, i_ = index.search_sharded(qd, k)
        state["host"] = (s_.cpu(), i_.cpu())

    e2e_step()
    ms_e2e = timed(e2e_step, max(1, min(2, steps)), world)
    # dominant kernel timed alone (this rank: all Q queries vs the local shard)
    allq = torch.randn(Q, d, generator=gd, device=dev).half() if world > 1 else q
    qi = ops.row_inv_norm(allq)
    splits, wsb = ops.ip_topk_plan(Q, hi - lo, k)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    outb = (torch.empty(Q, splits, k, device=dev), torch.empty(Q, splits, k, device=dev, dtype=torch.int32))
    ops.ip_topk_partial(allq, qi, index.local.emb, index.local.inv, k, workspace=ws, out=outb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.ip_topk_partial(allq, qi, index.local.emb, index.local.inv, k, workspace=ws, out=outb)
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1)
    fl = 2.0 * Q * (hi - lo) * d
    achieved = fl / (kms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "uniir::ip_topk_kernel_ts (fp16 tcgen05 GEMM, queries resident in TMEM + smem, "
                                         "candidates TMA-streamed, fused top-k epilogue)",
            "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"],
            "peak_src": f"{peaks['src']}: burst bf16 (kernel timed alone)", "kernel_ms": kms,
            "algorithmic_flops_per_launch": fl, "traffic": _ncu_traffic("search", hi - lo),
            "min_bytes_per_launch": (hi - lo) * d * 2 + Q * d * 2 + Q * k * 8 * splits,
            "fp16_cublas_peaks_this_pool": _fp16_peaks()}
    out = {"metric": "queries_per_sec_top50", "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f16", "data": "synthetic",
           "config": {"workload": f"dense retrieval {Q} queries x {N} candidates, d={d} fp16, top-{k} (BASELINE configs[4])",
                      "parallelism": f"candidate rows sharded over {world} GPU(s)", "splits_per_gpu": splits,
                      "l2": "candidate shard (>= 1 GB) >> 126 MB L2"},
           "clocks": clocks, "gpu_launches": launches, "roofline": roof,
           "e2e": {"value": Q / (ms_e2e / 1e3), "unit": "queries/s", "h2d_bytes_per_step": Q * d * 2,
                   "d2h_bytes_per_step": Q * k * 12, "ms_per_step": ms_e2e}}
    del index, cand
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------
# our arm: embedding extraction (BASELINE configs[3]: CLIP_SF-L forward over candidate batches)
# ----------------------------------------------------------------------------------------
def run_embed(args, rank, world, dev, peaks):
    from uniir_b200 import lib, ops
    from uniir_b200.clip_sf import CLIPScoreFusion
    B = 2048                                   # reference inference batch (embed.yaml:88)
    model = CLIPScoreFusion(MODEL, device=dev, micro_batch=256)
    model.eval()
    b = synth_batch(model.dims, B // 2, 2023 + rank, dev)
    b.pop("index_mapping"); b["did_list"] = list(range(B))
    state = {}

    def step():
        state["emb"], _ = model(b, encode_mbeir_batch=True)
        state["emb16"] = state["emb"].half()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(dev.index or 0); sampler.start()
    l0 = lib.launch_count() + ops.graph_launches()
    ms = timed(step, args.steps, world)
    launches = lib.launch_count() + ops.graph_launches() - l0
    clocks = sampler.stop()
    hb = {k: (torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) if torch.is_tensor(v) else v) for k, v in b.items()}

    def e2e_step():
        emb, _ = model(hb, encode_mbeir_batch=True)      # pinned host batch in, uploaded in chunks by the model
        state["host"] = emb.half().cpu()

    e2e_step()
    ms_e2e = timed(e2e_step, max(1, min(2, args.steps)), world)
    items = B * world
    alg = 175.33e9 * B / (ms / 1e3) / 1e12          # per GPU
    return {"metric": "embedded_items_per_sec", "value": items / (ms / 1e3), "unit": "items/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"embedding extraction, CLIP_SF {MODEL}, batch {B}/GPU (BASELINE configs[3] shape)",
                       "note": f"5.6M items = {5.6e6 / (items / (ms / 1e3)):.0f} s at this rate", "micro_batch_items": 256},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": alg, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": alg / peaks["bf16_sustained"], "traffic": None,
                         "note": "whole forward (175.33 GFLOP/item algorithmic), per GPU, against measured sustained bf16"},
            "e2e": {"value": items / (ms_e2e / 1e3), "unit": "items/s", "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in hb.values() if torch.is_tensor(v)) * world,
                    "d2h_bytes_per_step": B * 768 * 2 * world, "ms_per_step": ms_e2e}}


def synth_blip_batch(dims, pairs, seed, device, step=0):
    """SURVEY.md 8d BLIP inputs: input_ids [B,100] = [101, U{1000..vocab-3} x L, 102, 0..], prefix attention
    mask, randn images, query = even rows / pos_cand = odd rows, distinct positive-candidate ids."""
    n, S = 2 * pairs, 100
    g = torch.Generator().manual_seed(seed)
    L_ = torch.randint(1, S - 1, (n,), generator=g)
    ids = torch.randint(1000, dims.vocab - 2, (n, S), generator=g)
    pos = torch.arange(S)[None, :]
    ids = torch.where(pos == 0, torch.full_like(ids, 101), ids)
    ids = torch.where(pos == (L_ + 1)[:, None], torch.full_like(ids, 102), ids)
    att = (pos <= (L_ + 1)[:, None]).long()
    ids = ids * att
    gd = torch.Generator(device=device).manual_seed(seed)
    img = torch.randn(n, 3, dims.image_size, dims.image_size, generator=gd, device=device)
    return dict(txt_batched={"input_ids": ids.to(device), "attention_mask": att.to(device)}, image_batched=img,
                txt_mask_batched=None, image_mask_batched=None,
                p_did_list=torch.arange(pairs, dtype=torch.int64) + (seed * 1_000_003 + step * pairs),
                index_mapping={"query": [[2 * i] for i in range(pairs)], "pos_cand": [[2 * i + 1] for i in range(pairs)]})


BLIP_GLOBAL_BATCH = 2048
BLIP_FLOPS_PER_PAIR = 1.212e12     # SURVEY.md 8d (algorithmic fwd+bwd of the online encoders; momentum forward not credited)
# BLIP_SF: ViT-L/16 123.1 G + BERT-base in text mode at 100 tokens (12 x (QKV 0.354 + attention 0.031 + out 0.118 +
# FFN 0.944) = 17.4 G) = 140.5 GFLOP per item forward; x 2 items x (3 online + 1 momentum forward)
BLIP_SF_FLOPS_PER_PAIR = 1.124e12


def run_train_blip(args, rank, world, dev, peaks):
    """BASELINE configs[2]: BLIP_FF ViT-L/16 + BERT-base cross-attention encoder, momentum distillation with
    a 57344-entry queue, global batch 2048 pairs, text length 100."""
    from collections import Counter
    from uniir_b200 import lib, ops
    from uniir_b200.blip_ff import BLIPFeatureFusion
    from uniir_b200.blip_sf import BLIPScoreFusion
    from uniir_b200.engine import FlatAdamW, train_step_blip
    pairs = BLIP_GLOBAL_BATCH // world
    sf = args.blip_model == "sf"      # BLIP_SF (text-mode BERT + projection heads), same step otherwise
    flops_per_pair = BLIP_SF_FLOPS_PER_PAIR if sf else BLIP_FLOPS_PER_PAIR
    model = (BLIPScoreFusion if sf else BLIPFeatureFusion)(vit="large", image_size=224, embed_dim=768, queue_size=57344,
                                                           device=dev, micro_batch=256)
    model.train_micro_batch = args.blip_micro_batch
    model.train()                       # as uniir_blip/engine.py:16: dropout 0.1 / 0.1 in the BERT, DropPath 0.1 in the ViT-L
    if args.blip_no_dropout:
        model.set_dropout(0.0, 0.0, 0.0)
    opt = FlatAdamW(model._online, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    batch = synth_blip_batch(model.dims, pairs, 2023 + rank, dev)
    state = {}

    def step():
        state["loss"], state["acc"] = train_step_blip(model, batch, opt, alpha=0.4)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0); sampler.start()
    l0 = lib.launch_count() + ops.graph_launches()
    ms = timed(step, args.steps, world)
    launches = lib.launch_count() + ops.graph_launches() - l0
    clocks = sampler.stop()
    value = BLIP_GLOBAL_BATCH / (ms / 1e3)

    hb = {"ids": batch["txt_batched"]["input_ids"].cpu().pin_memory(), "att": batch["txt_batched"]["attention_mask"].cpu().pin_memory(),
          "img": torch.empty(batch["image_batched"].shape, pin_memory=True).copy_(batch["image_batched"])}
    h2d = sum(v.numel() * v.element_size() for v in hb.values())

    def e2e_step():
        gb = dict(batch)
        gb["txt_batched"] = {"input_ids": hb["ids"].to(dev, non_blocking=True), "attention_mask": hb["att"].to(dev, non_blocking=True)}
        gb["image_batched"] = hb["img"].to(dev, non_blocking=True)
        loss, _ = train_step_blip(model, gb, opt, alpha=0.4)
        state["e2e_loss"] = loss.item()

    e2e_step()
    ms_e2e = timed(e2e_step, 1, world)

    sink = []
    graphs_on, model.use_cuda_graphs = model.use_cuda_graphs, False
    ops.set_gemm_profiler(sink, events=True)     # one whole step executed eagerly: every GEMM launch, in place
    step()
    ops.set_gemm_profiler(None)
    model.use_cuda_graphs = graphs_on
    torch.cuda.synchronize()
    keyf = lambda c: tuple(sorted((k, str(v)) for k, v in c.items() if k != "events"))
    uniq, counts = {}, Counter()
    for c in sink:
        uniq[keyf(c)] = c
        counts[keyf(c)] += 1
    iso_ms = iso_fl = 0.0
    for k_, c in uniq.items():
        t = ops.time_gemm_config(c, iters=5, device=dev)
        iso_ms += t * counts[k_]
        iso_fl += 2.0 * c["M"] * c["N"] * c["K"] * counts[k_]
    roof = gemm_roofline(sink, iso_ms, iso_fl, peaks, ms, flops_per_pair * BLIP_GLOBAL_BATCH / world / (ms / 1e3) / 1e12, None)
    roof.update(launches=len(sink), distinct_configs=len(uniq),
                gemm_share_of_step=sum(c["events"][0].elapsed_time(c["events"][1]) for c in sink) / ms if ms else None)
    out = {"metric": "contrastive_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": ("BLIP_SF ViT-L/16 + text-mode BERT-base momentum-distillation training step (SURVEY 8f-3; BASELINE configs[2] shape)"
                                   if sf else "BLIP_FF ViT-L/16 + BERT-base momentum-distillation training step (BASELINE configs[2])"),
                      "model": "BLIP_SF large" if sf else "BLIP_FF large", "global_batch": BLIP_GLOBAL_BATCH, "pairs_per_gpu": pairs, "queue_size": 57344,
                      "seq_len": "197 vision tokens + 100 text tokens per item, 2 items per pair", "micro_batch_items": model.train_micro_batch,
                      "parallelism": f"dp{world}", "grad_cache": "embedding-level (1 extra online forward + momentum forward, not credited)",
                      "l2": "inputs+weights+activations per step >> 126 MB L2 (no flush needed)",
                      "dropout": (f"train mode as the reference: hidden {model.hidden_dropout_prob} / attention "
                                  f"{model.attention_probs_dropout_prob} in the BERT stack, DropPath {model.drop_path_rate} in the ViT "
                                  "(online and momentum encoders; counter-based masks regenerated in the re-forward)")},
           "clocks": clocks, "gpu_launches": launches, "roofline": roof,
           "e2e": {"value": BLIP_GLOBAL_BATCH / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d * world,
                   "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e},
           "loss": float(state["loss"].item()), "accuracy": float(state["acc"].item())}
    del model, opt, batch, hb
    torch.cuda.empty_cache()
    return out


def cpu_blip_reference(steps, warmup, pairs=2, sf=False):
    from oracle.blip_oracle import BLIPFeatureFusionOracle, BLIPScoreFusionOracle
    cores = _cpu_threads()
    torch.manual_seed(2023)
    model = (BLIPScoreFusionOracle if sf else BLIPFeatureFusionOracle)("large", embed_dim=768, queue_size=57344)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.05)
    b = synth_blip_batch(model.dims, pairs, 2023, torch.device("cpu"))
    b["p_did_list"] = b["p_did_list"].tolist()
    if sf:      # score fusion multiplies by the modality masks (blip_sf.py:168-169): both present
        b["txt_mask_batched"] = torch.ones(2 * pairs, dtype=torch.long)
        b["image_mask_batched"] = torch.ones(2 * pairs, dtype=torch.long)

    # train mode like the reference's step (uniir_blip/engine.py:16): hidden / attention dropout 0.1 and DropPath 0.1 at the
    # oracle's (reference-pinned) sites, masks from torch's RNG as in the reference
    import oracle.blip_oracle as bo
    dpr = torch.linspace(0, 0.1, model.dims.v_layers).tolist()

    def train_mode_masks(tag, kind, layer, x):
        if kind.startswith("path"):
            if dpr[layer] == 0.0:
                return x
            keep = (torch.rand(x.shape[0], *([1] * (x.ndim - 1))) >= dpr[layer]).to(x.dtype)
            return x * keep / (1 - dpr[layer])
        return torch.nn.functional.dropout(x, 0.1, training=True)

    def step():
        out = model.compute_contrastive_loss(b, alpha=0.4)
        out["loss"].backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    bo.DROPOUT = train_mode_masks
    try:
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    finally:
        bo.DROPOUT = None
    return {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"oracle {'BLIP_SF' if sf else 'BLIP_FF'} large fp32 in train mode (dropout 0.1 / 0.1, DropPath 0.1), {pairs} pairs/step "
                      f"(of the {BLIP_GLOBAL_BATCH}-pair batch), online+momentum forward, loss, backward, AdamW, {steps} timed steps"}, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "search", "embed", "blip"])
    ap.add_argument("--blip-micro-batch", type=int, default=64)
    ap.add_argument("--blip-model", default="ff", choices=["ff", "sf"], help="--workload blip: BLIP_FF (configs[2]) or BLIP_SF")
    ap.add_argument("--blip-no-dropout", action="store_true", help="--workload blip: train step without the reference's dropout / DropPath (A/B only)")
    ap.add_argument("--micro-batch", type=int, default=128, help="encoder items per micro-batch")
    ap.add_argument("--no-extra", action="store_true", help="skip the attached retrieval measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mask-mix", default="ones", choices=["ones", "mbeir"],
                    help="train workload: modality masks (SURVEY 8d); 'mbeir' = 1/3 text-only, 1/3 image-only, 1/3 both")
    ap.add_argument("--no-tower-skip", action="store_true", help="encode masked modalities like the reference does")
    ap.add_argument("--retain", type=lambda s: s if s == "auto" else int(s), default=None,
                    help="micro-batches whose activations stay in HBM between the embedding pass and the backward "
                         "(default: auto = as many as fit)")
    ap.add_argument("--pairs-per-gpu", type=int, default=None,
                    help="diagnostic: run this many pairs on this GPU (emulates one rank of a larger job; the JSON "
                         "line then reports that batch, not BASELINE's)")
    ap.add_argument("--search-q", type=int, default=SEARCH_Q)
    ap.add_argument("--search-n", type=int, default=SEARCH_N)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the uniir_b200 hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    cpu_ok = rank == 0 and not args.no_cpu_baseline and world == 1
    CPU_STEPS, CPU_WARM = 3, 1          # cpu_baseline: 3 timed steps after 1 warm-up, each a bounded sample of the workload
    if args.workload == "embed":
        out = run_embed(args, rank, world, dev, peaks)
        if cpu_ok:
            out["cpu_baseline"], _ = cpu_embed_reference(CPU_STEPS, CPU_WARM)
    elif args.workload == "blip":
        out = run_train_blip(args, rank, world, dev, peaks)
        if cpu_ok:
            out["cpu_baseline"], _ = cpu_blip_reference(CPU_STEPS, CPU_WARM, sf=args.blip_model == "sf")
    elif args.workload == "search":
        out = run_search(args, rank, world, dev, peaks)
        if cpu_ok:
            out["cpu_baseline"], _ = cpu_search_reference(CPU_STEPS, CPU_WARM)
    else:
        out = run_train(args, rank, world, dev, peaks)
        if not args.no_extra:
            # the other GPU configurations of BASELINE.json on the same line, each with its own roofline / e2e /
            # cpu_baseline: configs[4] retrieval, configs[2] BLIP_FF training, configs[3] embedding extraction
            import copy
            ex = out["extra"] = {}
            # (the sharded search and the extraction step get faster with N: more steps keep the timed region long
            # enough for the 200 ms clock sampler)
            ex["retrieval"] = run_search(args, rank, world, dev, peaks, steps=2 * world, warmup=3)
            a2 = copy.copy(args); a2.steps, a2.warmup = 2, 3
            ex["blip"] = run_train_blip(a2, rank, world, dev, peaks)
            a3 = copy.copy(args); a3.steps, a3.warmup = 8, 3
            ex["embed"] = run_embed(a3, rank, world, dev, peaks)
        if cpu_ok:
            out["cpu_baseline"], _ = cpu_train_reference(CPU_STEPS, CPU_WARM)
            if "extra" in out:
                out["extra"]["retrieval"]["cpu_baseline"], _ = cpu_search_reference(CPU_STEPS, CPU_WARM)
                out["extra"]["blip"]["cpu_baseline"], _ = cpu_blip_reference(CPU_STEPS, CPU_WARM)
                out["extra"]["embed"]["cpu_baseline"], _ = cpu_embed_reference(CPU_STEPS, CPU_WARM)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
