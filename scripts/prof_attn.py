import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hqtransformer_b200 as H
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=256)
eng = model.stage2.engine("bf16")
for t in (64, 32):
    print("attn", t, eng.bench_attention(256, t, iters=6))
torch.cuda.synchronize()
