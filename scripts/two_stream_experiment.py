"""Experiment (round 2): does running the batch as G independent sub-batches on G streams hide the per-launch fixed
latency of the dependent kernel chain?  G engines (private weights + cache), B/G rows each, one graph replay per engine
per step, all enqueued from one host thread.  Prints images/s for (B, G) combinations."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hqtransformer_b200 as H  # noqa: E402

cfg = os.path.join(os.path.dirname(H.__file__), "configs", "imagenet_l12.yaml")
dev = torch.device("cuda", 0)
S = 64


def rate(B, G, steps=3):
    b = B // G
    models = [H.ImageGPT2.from_config(cfg, device=0, precision="bf16", max_batch=b).eval() for _ in range(G)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    conds = [torch.randint(0, 1000, (b,), device=dev) for _ in range(G)]
    kw = dict(use_fp16=True, max_seq_len=S, is_tqdm=False)

    def step(i):
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                H.sampling_ihqgpt(models[g].stage2, b, conds[g], seed=i, row_offset=g * b, **kw)

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(2 + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del models
    torch.cuda.empty_cache()
    return B * steps / dt, dt / steps / S * 1e3


for B, G in [(256, 1), (256, 2), (256, 4), (512, 1), (512, 2), (512, 4)]:
    r, ms = rate(B, G)
    print(f"B={B} streams={G} rows/stream={B // G}: {r:8.1f} images/s  {ms:.3f} ms per top position", flush=True)
