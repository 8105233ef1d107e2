#!/bin/bash
mkdir -p gpurun_out
show='
import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_top_position"], d["gpu_launches"])
for k in d.get("kernels", []):
    if k["kernel"].startswith("ln+"): print("   ", k["kernel"], k["launches_per_position"], k["us"])'
for v in 1 2 3; do
echo "=== ln tail debug=$v"; HQ_LNTAIL_DEBUG=$v timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "$show"
done
