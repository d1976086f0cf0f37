"""Eager vs CUDA-graph training steps of the tiny CLIP_SF model under each attention mode (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from uniir_b200 import ops
from uniir_b200.clip_sf import CLIPScoreFusion
from uniir_b200.engine import FlatAdamW, train_step
from test_clip_gpu import _batch, _cuda

for mode in (0, 1, 2, 3):
    ops.attn_set_mode(mode)
    for use in (False, True, True):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=1)
        m.use_cuda_graphs = use
        opt = FlatAdamW(m.clip_model, lr=1e-3)
        gb = _cuda(_batch(m.dims, 16))
        losses = [train_step(m, gb, opt)[0].item() for _ in range(5)]
        print(f"mode {mode} graphs={use}: " + " ".join(f"{x:.5f}" for x in losses), flush=True)
