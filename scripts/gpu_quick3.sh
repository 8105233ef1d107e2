#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k gemm -p no:cacheprovider 2>&1 | tail -3
timeout 150 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 120 python scripts/trace_loop.py 256 --no-pdl 2>&1 | grep -E "gemm|span"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' | head -2
timeout 100 python bench.py --steps 3 --warmup 2 --batch 1024 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
