import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HQ_GEMM_TRACE"] = "1"
import torch
from hqtransformer_b200.engine import bench_gemm_shape
torch.cuda.init()
for (M, N, K, tile) in [(256, 4608, 64, 64), (256, 4608, 1536, 64), (256, 1536, 6144, 32), (1024, 4608, 1536, 256), (256, 4608, 1536, 128)]:
    mean, mn = bench_gemm_shape(M, N, K, tile, 6, 2, 1)
    print(M, N, K, tile, "mean", round(mean, 2), "min", round(mn, 2), flush=True)
