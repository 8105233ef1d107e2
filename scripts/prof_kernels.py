"""Small driver for ncu: runs each kernel family a few times on an ImageNet-L12-sized engine (random weights)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hqtransformer_b200 as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
model = H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=B)
eng = model.stage2.engine("bf16")
for kind in range(5):
    for M in (B, 4 * B):
        print("gemm", kind, M, eng.bench_gemm(kind, M, iters=2))
for t in (16, 32, 64):
    print("attn", t, eng.bench_attention(B, t, iters=4))
torch.cuda.synchronize()
