"""Run-to-run reproducibility of the gradients of one CLIP_SF training step (tiny model, eager and graphs):
which parameters differ between two identical runs, and by how much (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from uniir_b200 import ops
from uniir_b200.clip_sf import CLIPScoreFusion
from uniir_b200.engine import FlatAdamW, train_step
from test_clip_gpu import _batch, _cuda


def grads(use_graphs, side=True):
    m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=1)
    m.use_cuda_graphs = use_graphs
    opt = FlatAdamW(m.clip_model, lr=1e-3)
    gb = _cuda(_batch(m.dims, 16))
    loss, _ = train_step(m, gb, opt, do_step=False)
    torch.cuda.synchronize()
    P = m.clip_model
    return loss.item(), {name: P.g(name).clone() for name, *_ in P._specs}


mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ops.attn_set_mode(mode)
for use in (False, True):
    ref = grads(use)
    for rep in range(3):
        cur = grads(use)
        bad = []
        for k in ref[1]:
            a, b = ref[1][k], cur[1][k]
            d = (a - b).abs().max().item()
            if d > 0:
                bad.append((d / max(a.abs().max().item(), 1e-30), k, int(((a - b) != 0).sum().item()), a.numel()))
        bad.sort(reverse=True)
        print(f"mode {mode} graphs={use} rep {rep}: loss {ref[0]:.6f} vs {cur[0]:.6f}; {len(bad)} of {len(ref[1])} params differ; worst: "
              + "; ".join(f"{k} rel {r:.1e} ({n}/{t})" for r, k, n, t in bad[:6]), flush=True)
