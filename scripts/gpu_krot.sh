#!/bin/bash
mkdir -p gpurun_out
echo "--- k rotation ON"; timeout 120 python scripts/trace_gemm.py 2>&1 | grep -E "^256|^1024|last_mma|first_stage"
echo "--- k rotation OFF"; HQ_NO_KROT=1 timeout 120 python scripts/trace_gemm.py 2>&1 | grep -E "^256|^1024|last_mma|first_stage"
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gemm -p no:cacheprovider 2>&1 | tail -2
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
