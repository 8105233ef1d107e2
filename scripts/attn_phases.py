"""Where a decode-attention launch spends its time: per-CTA phase stamps (min / mean / max over CTAs, us since the first
CTA started).  HQ_ATTN_SCALAR=1 selects the scalar kernel.  usage: attn_phases.py [B] [keys...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hqtransformer_b200 as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
keys = [int(v) for v in sys.argv[2:]] or [8, 32, 64]
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=B)
eng = model.stage2.engine("bf16")
names = ["cta_start", "barriers_ready", "q_ready", "first_keys_landed", "scores_done", "softmax_done", "first_item_written", "cta_end"]
for t in keys:
    for rep in range(2):
        ph = eng.attention_phases(B, t, warm=3 + rep)
        t0 = ph[:, 0].min()
        print(f"[phases] keys={t} ctas={len(ph)} (us since first CTA start; min / mean / max over CTAs)")
        for p, nm in enumerate(names):
            col = ph[:, p]
            col = col[col > 0] - t0
            if len(col):
                print(f"[phases]   {nm:20s} {col.min() / 1e3:7.2f} {col.mean() / 1e3:7.2f} {col.max() / 1e3:7.2f}  (n={len(col)})")
torch.cuda.synchronize()
