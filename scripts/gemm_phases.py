"""Per-CTA phases of single GEMM launches INSIDE the replayed sampling loop (GPU box): HQ_DEBUG=1 HQ_TRACE_PDL=1 python scripts/gemm_phases.py
For the first launch of each GEMM family at top position 32: times (us) relative to the END of the previous kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hqtransformer_b200 as H
from hqtransformer_b200.engine import SamplingParams

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2(H.load_config(cfg), device=0, precision="bf16", max_batch=B)
model.stage2.init_weights(0)
s2 = model.stage2
eng = s2.engine("bf16")
cond = torch.randint(0, 1000, (B,), device="cuda")
ct, cb = H.sampling_ihqgpt(s2, B, cond, max_seq_len=64, is_tqdm=False)
torch.cuda.synchronize()
kw = dict(batch=B, seq_len=64, pos_begin=31, pos_end=34, sampling=SamplingParams(), cond=cond, codes_top=ct, codes_bot=cb)
tl = eng.trace_run(**kw)
tl = eng.trace_run(**kw)
per_pos = len(tl) // 3
names = ["start", "prologue_done", "dep_resolved", "first_stage", "last_mma", "accum_done", "epilogue_done", "end", "epi_tmem_read", "epi_parked", "epi_chunk_stored"]
seen = set()
print(f"# B={B}, {len(tl)} launches over 3 positions; PDL in trace: {os.environ.get('HQ_TRACE_PDL')}")
for i in range(per_pos + 20, 2 * per_pos + 20):          # the middle position, past its first layers
    tag = tl[i][0]
    if not tag.startswith("gemm") or tag in seen:
        continue
    seen.add(tag)
    t2, ph = eng.gemm_phases(launch_id=i, **kw)
    prev_end = t2[i - 1][2]
    ph = ph[ph[:, 0] > 0]
    if len(ph) == 0:
        print(f"{tag}: no stamps"); continue
    nxt = t2[i + 1]
    print(f"{tag}  (prev: {t2[i-1][0]}; {len(ph)} CTAs)  kernel span {(t2[i][1]-prev_end)/1e3:+.2f} .. {(t2[i][2]-prev_end)/1e3:+.2f} us; "
          f"next {nxt[0]} starts {(nxt[1]-prev_end)/1e3:+.2f} ends {(nxt[2]-prev_end)/1e3:+.2f}")
    if "--per-cta" in sys.argv and ("gemm_qkv:256x4608" in tag or "gemm_resid:256x1536x1536" in tag or "gemm_fc1_gelu:256" in tag or "gemm_resid:256x1536x6144" in tag):
        # CTA index -> start / dependency resolved / end: is the pair grid dispatched in index order, and how fast?
        print("    cta: start dep_resolved end   " + "  ".join(
            f"{c}:{(ph[c,0]-prev_end)/1e3:+.2f}/{(ph[c,2]-prev_end)/1e3:+.2f}/{(ph[c,7]-prev_end)/1e3:+.2f}" for c in range(0, len(ph), 8)))
    for p, nm in enumerate(names):
        v = ph[:, p]
        v = v[v > 0]
        if len(v):
            r = (v - prev_end) / 1e3
            print(f"    {nm:14s} min {r.min():+7.2f}  mean {r.mean():+7.2f}  max {r.max():+7.2f}   (n={len(v)})")
