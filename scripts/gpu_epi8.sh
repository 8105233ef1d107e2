#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gemm -p no:cacheprovider 2>&1 | tail -3
echo "=== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
show='
import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_top_position"], d["gpu_launches"])
for k in d.get("kernels", []): print("   ", k["kernel"], k["launches_per_position"], k["us"])'
echo "=== bench"; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_depth4.json | python -c "$show"
