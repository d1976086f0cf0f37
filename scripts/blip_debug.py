import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle.blip_oracle import BLIPFeatureFusionOracle
from uniir_b200.blip_ff import BLIPFeatureFusion
from test_oracle_blip import _inputs
torch.manual_seed(3)
orc = BLIPFeatureFusionOracle("tiny-test", embed_dim=64, queue_size=16).eval()
SC = float(os.environ.get("SC", 1))
with torch.no_grad():
    for n_, p_ in orc.named_parameters():
        if p_.ndim >= 2 and not n_.endswith(("_m",)) and "_m." not in n_:
            p_.mul_(SC)
orc.copy_params()
m = BLIPFeatureFusion(vit="tiny-test", image_size=64, embed_dim=64, queue_size=16, micro_batch=8)
m.train_micro_batch = int(os.environ.get("MB", 8))
m.load_state_dict(orc.state_dict(), strict=False)
d = orc.dims
ids, att, img = _inputs(d, 16, 24, 20)
b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, p_did_list=[5, 6, 5, 7, 8, 9, 10, 11],
         index_mapping={"query": [[2 * i] for i in range(8)], "pos_cand": [[2 * i + 1] for i in range(8)]})
ref = orc.compute_contrastive_loss(b, alpha=0.3); ref["loss"].backward()
out = m.compute_contrastive_loss(b, alpha=0.3); m.backward_contrastive()
print("loss", out["loss"].item(), ref["loss"].item())
rows = []
for n_, p_o in orc.named_parameters():
    if p_o.grad is None or p_o.numel() < 2: continue
    a, c = p_o.grad.flatten().double(), m._online.g(n_).cpu().flatten().double()
    rows.append(((torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item(), (c.norm() / (a.norm() + 1e-30)).item(), a.norm().item(), n_))
for r in sorted(rows)[:30]:
    print(f"cos={r[0]:.4f} ratio={r[1]:.3f} |g|={r[2]:.2e} {r[3]}")
