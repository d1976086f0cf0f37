"""One eager CLIP_SF ViT-L/14 training step for profiling under ncu (1 micro-batch pair).
    ncu --metrics gpu__time_duration.sum ... python scripts/prof_step.py [pairs] [mb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from uniir_b200.clip_sf import CLIPScoreFusion
from uniir_b200.engine import FlatAdamW, train_step

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = CLIPScoreFusion("ViT-L/14", device="cuda", micro_batch=mb)
m.use_cuda_graphs = False
m.retain_micro_batches = 0      # profile the common micro-batch: embedding pass, then re-forward + backward
opt = FlatAdamW(m.clip_model, lr=1e-5)
b = bench.synth_batch(m.dims, pairs, 1, torch.device("cuda"))
train_step(m, b, opt)            # warm-up (ncu: skipped with -s)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
train_step(m, b, opt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
