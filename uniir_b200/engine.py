"""Training engine for the B200-native CLIPScoreFusion.

``train_one_epoch`` keeps the reference signature (src/models/uniir_clip/engine.py:7) so the
reference's train.py loop can call it unchanged.  ``train_step`` is the fused path: gradient-
cached forward/backward on the C-ABI kernels, one flat gradient all-reduce (DDP semantics:
mean over ranks, reference train.py:217-219), one fused AdamW kernel per parameter group
(reference param groups train.py:52-61: wd 0.2 on >=2-D weights, 0 elsewhere; betas
(0.9, 0.98), eps 1e-6).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import ops


class FlatAdamW:
    """AdamW over FlatParamModule's two flat segments.  ``state_dict()`` / ``load_state_dict()`` speak
    ``torch.optim.AdamW``'s layout for the reference's two parameter groups (clip_scorefusion/train.py:52-61:
    group 0 = gains / biases / logit_scale without weight decay, group 1 = the rest with it, each in
    ``named_parameters()`` order), so the ``optimizer`` entry of a reference checkpoint (train.py:64-79) resumes here
    and a checkpoint written here resumes in the reference (train.py:204-211)."""

    def __init__(self, flat_params, lr=1e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2, lr_scale=None):
        """lr_scale: optional callable(parameter name) -> multiplier of the learning rate for that parameter.  The
        reference trains the T5 fusion layers of CLIP_FF at ``t5_learning_rate`` (5e-4) and the CLIP towers at
        ``learning_rate`` (1e-5) (clip_featurefusion/train.py:49-63, inbatch.yaml:48-49); a scheduler scales both, as
        torch's does.  Adjacent parameters with the same multiplier share one kernel launch over their sub-range
        of the flat segment; without lr_scale each segment is one launch, as before."""
        self.P = flat_params
        self.P.ensure_grad()
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = {k: torch.zeros_like(v) for k, v in self.P.flat.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.flat.items()}
        self.step_count = 0
        self.param_groups = [{"lr": lr}]  # MetricLogger reads optimizer.param_groups[0]["lr"]
        self._ranges = None
        if lr_scale is not None:
            A = self.P.ALIGN
            self._ranges = {k: [] for k in self.P.flat}
            for name, *_ in self.P._specs:                      # allocation order within each segment
                seg, off, n, _ = self.P._index[name]
                hi, sc = off + (n + A - 1) // A * A, float(lr_scale(name))
                r = self._ranges[seg]
                if r and r[-1][2] == sc and r[-1][1] == off:
                    r[-1][1] = hi
                else:
                    r.append([off, hi, sc])

    def step(self, grad_scale=1.0):
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        for decay in (True, False):
            wd = self.wd if decay else 0.0
            if self._ranges is None:
                ops.adamw_step(self.P.flat[decay], self.P.grad[decay], self.m[decay], self.v[decay],
                               self.P.flat16[decay], lr, self.betas, self.eps, wd, self.step_count, grad_scale)
                continue
            for lo, hi, sc in self._ranges[decay]:
                ops.adamw_step(self.P.flat[decay][lo:hi], self.P.grad[decay][lo:hi], self.m[decay][lo:hi],
                               self.v[decay][lo:hi], self.P.flat16[decay][lo:hi], lr * sc, self.betas, self.eps, wd,
                               self.step_count, grad_scale)
        self.P.mark_shadow_fresh()

    def zero_grad(self, set_to_none=False):
        for g in self.P.grad.values():
            g.zero_()

    def _group_names(self):
        """Parameter names of the reference's two groups, in the order torch numbers them."""
        trainable = [(n, decay) for n, _, decay, _ in self.P._specs if self.P._root.get_parameter(n).requires_grad]
        return [n for n, decay in trainable if not decay], [n for n, decay in trainable if decay]

    def state_dict(self):
        if self._ranges is not None:     # per-parameter lr multipliers (CLIP_FF): flat layout, not the reference's groups
            return {"uniir_flat": True, "step": self.step_count, "exp_avg": self.m, "exp_avg_sq": self.v,
                    "param_groups": self.param_groups}
        groups = self._group_names()
        lr = self.param_groups[0]["lr"]
        proto = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=lr, betas=self.betas, eps=self.eps).state_dict()
        state, param_groups, idx = {}, [], 0
        for names, wd in zip(groups, (0.0, self.wd)):
            ids = list(range(idx, idx + len(names)))
            for n in names:
                if self.step_count > 0:
                    state[idx] = {"step": torch.tensor(float(self.step_count)),
                                  "exp_avg": self.P._view(self.m, n).clone(), "exp_avg_sq": self.P._view(self.v, n).clone()}
                idx += 1
            g = dict(proto["param_groups"][0])
            g.update(lr=lr, betas=tuple(self.betas), eps=self.eps, weight_decay=wd, params=ids)
            if "initial_lr" in self.param_groups[0]:
                g["initial_lr"] = self.param_groups[0]["initial_lr"]
            param_groups.append(g)
        return {"state": state, "param_groups": param_groups}

    def load_state_dict(self, sd):
        if "state" not in sd:            # flat layout (round 1 / lr-multiplier runs)
            self.step_count = sd["step"]
            for k in self.m:
                self.m[k].copy_(sd["exp_avg"][k])
                self.v[k].copy_(sd["exp_avg_sq"][k])
            self.param_groups[0]["lr"] = sd["param_groups"][0]["lr"]
            return
        groups = self._group_names()
        assert len(sd["param_groups"]) == 2 and [len(g["params"]) for g in sd["param_groups"]] == [len(g) for g in groups], \
            "optimizer state does not have the reference's two parameter groups (train.py:52-61)"
        steps = set()
        for g, names in zip(sd["param_groups"], groups):
            for pid, n in zip(g["params"], names):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                self.P._view(self.m, n).copy_(st["exp_avg"].to(self.m[True].device, torch.float32).view(self.P._index[n][3]))
                self.P._view(self.v, n).copy_(st["exp_avg_sq"].to(self.v[True].device, torch.float32).view(self.P._index[n][3]))
                steps.add(int(float(st["step"])))
        assert len(steps) <= 1, "parameters with different step counts cannot share the fused AdamW step"
        self.step_count = steps.pop() if steps else 0
        self.param_groups[0]["lr"] = sd["param_groups"][0]["lr"]
        if "initial_lr" in sd["param_groups"][0]:
            self.param_groups[0]["initial_lr"] = sd["param_groups"][0]["initial_lr"]


class CosineLR:
    """torch.optim.lr_scheduler.CosineAnnealingLR(T_max, eta_min=0) for FlatAdamW
    (reference clip_scorefusion/train.py:281-284)."""

    def __init__(self, optimizer, T_max, eta_min=0.0):
        self.opt, self.T_max, self.eta_min = optimizer, T_max, eta_min
        self.base = optimizer.param_groups[0]["lr"]
        self.t = 0

    def step(self):
        self.t += 1
        self.opt.param_groups[0]["lr"] = self.eta_min + (self.base - self.eta_min) * \
            (1 + math.cos(math.pi * self.t / self.T_max)) / 2

    def state_dict(self):
        return {"t": self.t, "base": self.base, "T_max": self.T_max}

    def load_state_dict(self, sd):
        self.t, self.base, self.T_max = sd["t"], sd["base"], sd["T_max"]


def allreduce_grads(P):
    """DDP semantics: parameter gradients averaged over ranks; one NCCL call per flat segment
    (in-switch NVLS reduction when NCCL enables it)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for g in P.grad.values():
            dist.all_reduce(g, op=dist.ReduceOp.AVG)


def train_step(model, batch, optimizer: FlatAdamW, scheduler=None, accumulation_steps=1, do_step=True):
    """One optimisation step on one (already device-resident or host) batch.  Returns the loss and
    accuracy as 0-d device tensors (no host sync)."""
    loss, acc = model._loss_forward(batch)
    loss, acc = loss.clone(), acc.clone()
    model._loss_backward(batch, None)
    if do_step:
        allreduce_grads(optimizer.P)          # the flat parameter set being optimised (CLIP_SF: model.clip_model)
        optimizer.step(grad_scale=1.0 / accumulation_steps)
        if scheduler is not None:
            scheduler.step()
    return loss, acc


def train_one_epoch(model, data_loader, optimizer, epoch, gpu_id, scheduler, global_step, scaler, config):
    """Reference-compatible epoch loop (uniir_clip/engine.py:7-55).  ``scaler`` is accepted and
    ignored: the bf16/fp32 path needs no loss scaling.  Logs like the reference but syncs with
    the device only every ``print_freq`` steps instead of three times per step."""
    model.train()
    inner = model.module if hasattr(model, "module") else model
    print_freq = config.trainer_config.print_freq
    accumulation_steps = config.trainer_config.gradient_accumulation_steps
    fast = isinstance(optimizer, FlatAdamW)
    counter = 0
    losses, accs = [], []
    for i, batch in enumerate(data_loader):
        for key in batch:
            if isinstance(batch[key], torch.Tensor):
                if fast and key == "image_batched" and batch[key].is_pinned():
                    continue    # the model uploads pinned images in micro-batch chunks, overlapped with the towers
                batch[key] = batch[key].to(gpu_id, non_blocking=True)
        counter += 1
        stepping = counter == accumulation_steps
        if fast:
            loss, acc = train_step(inner, batch, optimizer, scheduler, accumulation_steps, do_step=stepping)
        else:  # generic torch optimizer through the autograd bridge
            out = model(batch)
            loss, acc = out["loss"], out["accuracy"]
            (loss / accumulation_steps).backward()
            if stepping:
                optimizer.step()
                model.zero_grad()
                scheduler.step()
        if stepping:
            global_step += 1
            counter = 0
        losses.append(loss.detach())
        accs.append(acc.detach())
        if (i + 1) % print_freq == 0:
            print(f"Train Epoch: [{epoch}] [{i + 1}/{len(data_loader)}] loss: {losses[-1].item():.4f} "
                  f"inbatch_accuracy: {accs[-1].item():.4f} lr: {optimizer.param_groups[0]['lr']:.6f}")
    stats = {"loss": torch.stack(losses).mean().item() if losses else 0.0,
             "inbatch_accuracy": torch.stack(accs).mean().item() if accs else 0.0,
             "lr": optimizer.param_groups[0]["lr"]}
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([stats["loss"], stats["inbatch_accuracy"]], device=gpu_id, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        stats["loss"], stats["inbatch_accuracy"] = t[0].item(), t[1].item()
    print("Averaged stats:", stats)
    return stats


@torch.no_grad()
def eval_engine(model, data_loader, gpu_id, config):
    """In-batch validation (uniir_clip/engine.py:59-84): mean loss / in-batch accuracy over the loader, averaged
    over ranks.  No gradients, no activations kept; one host sync at the end instead of two per batch."""
    model.eval()
    inner = model.module if hasattr(model, "module") else model
    print_freq = config.evaluator.print_freq
    losses, accs = [], []
    for i, batch in enumerate(data_loader):
        for key in batch:
            if isinstance(batch[key], torch.Tensor):
                batch[key] = batch[key].to(gpu_id, non_blocking=True)
        out = inner(batch)
        losses.append(out["loss"].detach())
        accs.append(out["accuracy"].detach())
        if (i + 1) % print_freq == 0:
            print(f"Test: [{i + 1}/{len(data_loader)}] loss: {losses[-1].item():.4f} inbatch_accuracy: {accs[-1].item():.4f}")
    stats = {"loss": torch.stack(losses).mean().item() if losses else 0.0,
             "inbatch_accuracy": torch.stack(accs).mean().item() if accs else 0.0}
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([stats["loss"], stats["inbatch_accuracy"]], device=gpu_id, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        stats["loss"], stats["inbatch_accuracy"] = t[0].item(), t[1].item()
    print("Averaged stats:", stats)
    return stats


def train_step_blip(model, batch, optimizer: FlatAdamW, alpha, scheduler=None):
    """One BLIP_FF optimisation step (reference uniir_blip/engine.py:9-66 inner loop): momentum-distillation
    loss with gradient caching, flat gradient all-reduce, fused AdamW (single group, wd 0.05,
    uniir_blip/train.py:193-197)."""
    loss, acc = model._loss_forward(batch, alpha)
    model.backward_contrastive()
    allreduce_grads(model._online)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss, acc
