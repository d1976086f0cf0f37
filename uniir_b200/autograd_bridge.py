"""Autograd bridge for the gradient-cached models: the reference training loops call
``loss = model(batch)["loss"]; scaler.scale(loss).backward(); optimizer.step()`` (uniir_clip/engine.py:26-45,
uniir_blip/engine.py:35-56).  The B200 models compute their gradients with hand-written kernels into flat buffers, not
with autograd; this Function makes the returned loss a proper graph node whose backward runs that hand-written
backward and hands the gradients to autograd, so ``p.grad`` is filled for any stock torch optimizer.  The fused
engines (``engine.train_step*``) bypass it."""
from __future__ import annotations

import torch


class _GradCacheLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, names, forward_fn, backward_fn, *params):
        loss, acc = forward_fn()
        ctx.flat, ctx.names, ctx.backward_fn = flat, names, backward_fn
        loss, acc = loss.clone(), acc.clone()
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, _g_acc):
        flat = ctx.flat
        flat.ensure_grad()
        ctx.backward_fn(g_loss)
        grads = tuple(flat.g(n).clone() for n in ctx.names)
        for buf in flat.grad.values():      # the flat buffers only carried this call's gradients
            buf.zero_()
        return (None, None, None, None) + grads


def bridged_loss(flat, forward_fn, backward_fn):
    """forward_fn() -> (loss, acc); backward_fn(g_loss) accumulates d(loss * g_loss)/d(param) into ``flat.grad``.
    Returns (loss, acc) with ``loss`` attached to every trainable parameter of ``flat``."""
    names = [n for n, *_ in flat._specs if flat._root.get_parameter(n).requires_grad]
    params = [flat._root.get_parameter(n) for n in names]
    return _GradCacheLoss.apply(flat, names, forward_fn, backward_fn, *params)
