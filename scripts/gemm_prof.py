"""Producer / MMA-issuer loop timings of the CTA-pair GEMM (GPU box): HQ_DEBUG=1 HQ_GEMM_PROF=1 python scripts/gemm_prof.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hqtransformer_b200.engine import bench_gemm_shape
torch.cuda.init()
for (M, N, K, tile) in [(256, 4608, 1536, 64), (256, 6144, 1536, 96), (256, 4608, 6144, 64), (256, 9216, 1536, 128),
                        (1024, 4608, 1536, 256), (1024, 6144, 1536, 192)]:
    for flush in (2, 0):
        mean, mn = bench_gemm_shape(M, N, K, tile, 10, flush, 1)
        print(f"M={M} N={N} K={K} tile={tile} flush={flush}: mean {mean:.2f} us min {mn:.2f} us", flush=True)
