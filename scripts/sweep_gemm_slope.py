"""Per-k-block slope of the CTA-pair GEMM (GPU box): time vs K at fixed grid (72 pairs), for several tile widths.
slope = (t(K=6144) - t(K=1536)) / 72 k-blocks; bytes per k-block per CTA = 16 KB (A) + tile * 64 B (W half-tile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hqtransformer_b200.engine import bench_gemm_shape
torch.cuda.init()
print("# KS1 =", os.environ.get("HQ_GEMM_KS1"), " M N K tile flush: mean_us min_us")
for M in (256, 1024):
    for tile in (32, 64, 128, 256):
        N = tile * 72 // (M // 256)
        res = {}
        for K in (512, 1536, 3072, 6144):
            mean, mn = bench_gemm_shape(M, N, K, tile, 15, 2, 1)
            res[K] = mn
            print(M, N, K, tile, round(mean, 2), round(mn, 2), flush=True)
        slope = (res[6144] - res[1536]) / 72.0 * 1000.0
        byt = 16384 + tile * 64
        print(f"#   M={M} tile={tile}: {slope:.0f} ns per 64-wide k-block, {byt} B per CTA -> {byt / slope:.1f} GB/s per CTA", flush=True)
