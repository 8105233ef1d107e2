"""Stand-alone timing of the spatial decode attention at batch 256 (ImageNet L12) over cache lengths.
HQ_ATTN_SCALAR=1 selects the scalar bulk-staged kernel.  Prints us per launch and GB/s of algorithmic K+V+q+out bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hqtransformer_b200 as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=B)
eng = model.stage2.engine("bf16")
peak = 6549.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
D = 1536
tot_b = tot_us = 0.0
for t in (1, 4, 8, 16, 17, 24, 32, 33, 40, 48, 56, 64):
    us = eng.bench_attention(B, t, iters=24)
    by = B * (2 * t * D + 2 * D) * 2
    tot_b += by; tot_us += us
    print(f"attn keys {t:3d}: {us:7.2f} us  {by / us * 1e-3:7.1f} GB/s  {by / us * 1e-3 / peak:.3f} of {peak:.0f}")
print(f"all listed lengths: {tot_b / tot_us * 1e-3:.1f} GB/s = {tot_b / tot_us * 1e-3 / peak:.3f}")
torch.cuda.synchronize()
