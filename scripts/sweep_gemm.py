"""GEMM microbenchmark sweep (GPU box): fixed overhead vs per-k-block slope, tile variants, flush modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hqtransformer_b200.engine import bench_gemm_shape
torch.cuda.init()
print("M N K tile flush copies mean_us min_us")
def run(M, N, K, tile, flush, copies=1, iters=15):
    mean, mn = bench_gemm_shape(M, N, K, tile, iters, flush, copies)
    print(M, N, K, tile, flush, copies, round(mean, 2), round(mn, 2), flush=True)
# fixed overhead: K sweep, QKV shape, flush modes
for flush in (0, 2, 1):
    for K in (64, 512, 1536, 6144):
        run(256, 4608, K, 64, flush)
# tile variants at M=256 QKV / proj / fc1 / fc2 (clean read flush)
for tile in (32, 64, 128, -64, -128):
    run(256, 4608, 1536, tile, 2)
for tile in (32, 64, -64):
    run(256, 1536, 1536, tile, 2)
    run(256, 1536, 6144, tile, 2)
for tile in (64, 96, 128, 192, -64, -128):
    run(256, 6144, 1536, tile, 2)
# M = 1024
for tile in (128, 192, 256, -128):
    run(1024, 4608, 1536, tile, 2)
    run(1024, 6144, 1536, tile, 2)
for tile in (64, 96, 128, -128):
    run(1024, 1536, 1536, tile, 2)
    run(1024, 1536, 6144, tile, 2)
# weights cycled without flush (12 copies ~ the 12 layers)
run(256, 4608, 1536, 64, 0, copies=12, iters=36)
run(256, 1536, 6144, 32, 0, copies=12, iters=36)
run(1024, 4608, 1536, 256, 0, copies=12, iters=36)
