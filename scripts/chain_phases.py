"""Per-CTA phase timestamps of single ops inside the persistent chain kernel (hq_debug_chain_phases), ImageNet L12, B = 256."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hqtransformer_b200 as H  # noqa: E402
from hqtransformer_b200.engine import SamplingParams  # noqa: E402

B, S = int(os.environ.get("B", 256)), 4
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=B).eval()
eng = model.stage2.engine("bf16")
cond = torch.randint(0, 1000, (B,), device="cuda")
ct = torch.zeros(B, S, dtype=torch.int64, device="cuda")
cb = torch.zeros(B, S, 4, dtype=torch.int64, device="cuda")
sp = SamplingParams(seed=1)
names = ["begin", "barrier_seen", "released", "work_done", "proxy_fence", "all_warps", "arrived"]
# position 2 = launches 28..41: 28 = [LN1, QKV]; 29..39 = [proj, LN2, fc1, fc2, LN1, QKV]; 40 = last spatial + depth pass 0
for launch, op, what in [(33, 1, "LN2 after proj"), (33, 4, "LN1 after fc2"), (33, 2, "fc1"), (33, 3, "fc2"), (33, 0, "proj"),
                         (33, 5, "qkv"), (41, 0, "embed_depth"), (41, 3, "attn4")]:
    for rep in range(2):
        t = eng.chain_phases(batch=B, seq_len=S, pos_begin=0, pos_end=S, sampling=sp, cond=cond, codes_top=ct, codes_bot=cb,
                             launch_idx=launch, op_idx=op)
    if os.environ.get("RAW"):
        print(t.shape)
        base = t[t > 0].min()
        for r in list(range(0, 4)) + list(range(146, 150)) + list(range(len(t) - 2, len(t))):
            if r < len(t):
                print(r, [int(v - base) if v else 0 for v in t[r]])
    t0 = t[:, 0].min()
    print(f"--- launch {launch} op {op} ({what}): {len(t)} CTAs; ns since the first CTA began the op (min / mean / max over CTAs)")
    for p, n in enumerate(names):
        col = t[:, p]
        col = col[col > 0] - t0
        if len(col):
            print(f"    {n:14s} {col.min():7d} {int(col.mean()):7d} {col.max():7d}")
