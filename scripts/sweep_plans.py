"""Which (tile, split-K) plan is fastest per GEMM shape with ONE CTA per SM?  Event-timed (includes ~6 us of launch
overhead per launch; compare within a shape)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from hqtransformer_b200.engine import bench_gemm_shape
    torch.cuda.init()
    s = int(os.environ.get("HQ_BENCH_SPLITS", "1"))
    shapes = [(1024, 1536, 1536), (1024, 1536, 6144), (256, 1536, 1536), (256, 1536, 6144)] if s > 1 else \
             [(1024, 1536, 1536), (1024, 1536, 6144), (1024, 6144, 1536), (1024, 8192, 1536), (1024, 4608, 1536),
              (256, 4608, 1536), (256, 6144, 1536), (256, 3072, 1536), (256, 8192, 1536), (256, 1536, 1536), (256, 1536, 6144)]
    for (M, N, K) in shapes:
        for tile in (32, 64, 96, 128, 192, 256):
            if N % tile or (K // 64) % s:
                continue
            pairs = ((M + 255) // 256) * (N // tile) * s
            if pairs > 300:
                continue
            mean, mn = bench_gemm_shape(M, N, K, tile, 12, 2, 1)
            print(f"{M}x{N}x{K} s{s} tile {tile:3d} pairs {pairs:3d}: mean {mean:6.2f} min {mn:6.2f}", flush=True)
else:
    for s in ("1", "2", "3", "4", "6"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, HQ_BENCH_SPLITS=s))
