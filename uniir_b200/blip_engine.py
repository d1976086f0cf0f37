"""Training / validation loops of the BLIP arms with the reference's signatures (src/models/uniir_blip/engine.py):

  * ``train_one_epoch(model, data_loader, optimizer, epoch, gpu_id, scheduler, global_step, scaler, config)``
    (engine.py:9-66): the alpha ramp of the momentum-distillation loss -- ``config.model.alpha`` after epoch 0,
    ``alpha * min(1, i / len(loader))`` inside it (:29-32) -- gradient accumulation, the fused step
    (``engine.train_step_blip``: gradient-cached forward / backward, flat all-reduce, fused AdamW);
  * ``eval_engine(model_without_ddp, model, data_loader, gpu_id, config)`` (engine.py:69-114): the forward pass of
    this model updates the momentum encoders and the queues even without gradients, so the reference snapshots the
    whole state, resets the queues (random features, ids -100, pointer 0), evaluates with the in-epoch alpha ramp
    and restores the snapshot; same here.
``scaler`` is accepted and ignored: the bf16 / fp32 path needs no loss scaling.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .engine import FlatAdamW, allreduce_grads


def _to_device(batch, gpu_id):
    for key, value in batch.items():
        if isinstance(value, torch.Tensor):
            batch[key] = value.to(gpu_id, non_blocking=True)
        elif hasattr(value, "items") and not isinstance(value, dict) or (isinstance(value, dict) and key == "txt_batched"):
            for k, v in value.items():                      # HF BatchEncoding / dict of input_ids, attention_mask
                if isinstance(v, torch.Tensor):
                    batch[key][k] = v.to(gpu_id)
    return batch


def _avg_over_ranks(stats, gpu_id):
    if dist.is_available() and dist.is_initialized():
        keys = sorted(stats)
        t = torch.tensor([stats[k] for k in keys], device=gpu_id, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        stats = {k: t[i].item() for i, k in enumerate(keys)}
    return stats


def epoch_alpha(config, epoch, i, n_batches):
    """uniir_blip/engine.py:29-32."""
    return config.model.alpha if epoch > 0 else config.model.alpha * min(1, i / n_batches)


def train_one_epoch(model, data_loader, optimizer, epoch, gpu_id, scheduler, global_step, scaler, config):
    model.train()
    inner = model.module if hasattr(model, "module") else model
    assert isinstance(optimizer, FlatAdamW), "the BLIP arms train through the fused path (FlatAdamW over model._online)"
    print_freq = config.trainer_config.print_freq
    accumulation_steps = config.trainer_config.gradient_accumulation_steps
    counter, losses, accs = 0, [], []
    n_batches = len(data_loader)
    for i, batch in enumerate(data_loader):
        batch = _to_device(batch, gpu_id)
        alpha = epoch_alpha(config, epoch, i, n_batches)
        loss, acc = inner._loss_forward(batch, alpha)
        inner.backward_contrastive()                         # gradients accumulate in the flat buffers
        counter += 1
        if counter == accumulation_steps:
            global_step += 1
            allreduce_grads(optimizer.P)
            optimizer.step(grad_scale=1.0 / accumulation_steps)   # also clears the gradients
            scheduler.step()
            counter = 0
        losses.append(loss.detach())
        accs.append(acc.detach())
        if (i + 1) % print_freq == 0:
            print(f"Train Epoch: [{epoch}] [{i + 1}/{n_batches}] loss: {losses[-1].item():.4f} "
                  f"inbatch_accuracy: {accs[-1].item():.4f} lr: {optimizer.param_groups[0]['lr']:.6f} alpha: {alpha:.4f}")
    stats = {"loss": torch.stack(losses).mean().item() if losses else 0.0,
             "inbatch_accuracy": torch.stack(accs).mean().item() if accs else 0.0}
    stats = _avg_over_ranks(stats, gpu_id)
    stats["lr"] = optimizer.param_groups[0]["lr"]
    print("Averaged stats:", stats)
    return stats


@torch.no_grad()
def eval_engine(model_without_ddp, model, data_loader, gpu_id, config):
    model.eval()
    m = model_without_ddp
    print_freq = config.evaluator.print_freq
    saved_state = {k: v.detach().clone() for k, v in m.state_dict().items()}          # engine.py:77-78
    m.query_queue.copy_(torch.randn_like(m.query_queue))                                 # :80-84
    m.cand_queue.copy_(torch.randn_like(m.cand_queue))
    m.idx_queue.copy_(torch.full_like(m.idx_queue, -100))
    m.new_ptr_queue.zero_()
    print("Cleared model queue states.")
    losses, accs = [], []
    n_batches = len(data_loader)
    for i, batch in enumerate(data_loader):
        batch = _to_device(batch, gpu_id)
        alpha = config.model.alpha * min(1, i / n_batches)                                # :94
        out = m(batch=batch, alpha=alpha)
        losses.append(out["loss"].detach())
        accs.append(out["accuracy"].detach())
        if (i + 1) % print_freq == 0:
            print(f"Test: [{i + 1}/{n_batches}] loss: {losses[-1].item():.4f} inbatch_accuracy: {accs[-1].item():.4f}")
    stats = {"loss": torch.stack(losses).mean().item() if losses else 0.0,
             "inbatch_accuracy": torch.stack(accs).mean().item() if accs else 0.0}
    stats = _avg_over_ranks(stats, gpu_id)
    print("Averaged stats:", stats)
    m.load_state_dict(saved_state)                                                        # :110-112
    m._train_state = None
    print("Restored model queue states and model states from the saved variables.")
    return stats
