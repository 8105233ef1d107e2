#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "=== tma attention phases"; timeout 120 python scripts/attn_phases.py 256 8 32 64 2>&1 | grep phases | awk '/keys=/{n++} n%2==1'
echo "=== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
show='
import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_top_position"], d["roofline_attention"]["achieved"])
for k in d["kernels"]: print("   ", k["kernel"], k["launches_per_position"], k["us"])'
echo "=== bench tma attention"; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_tma.json | python -c "$show"
echo "=== bench tma attention stages=2"; HQ_ATTM_STAGES=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$show" | head -1
echo "=== bench scalar attention"; HQ_ATTN_SCALAR=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$show" | grep -E "^[0-9]|attention_decode"
