#!/bin/bash
mkdir -p gpurun_out
for bn in 0 96 128 192; do
  HQ_DEBUG=1 HQ_BN_M256=$bn timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/bn_$bn.log 2>&1
  python - $bn <<'P'
import json,sys
for line in open(f"gpurun_out/bn_{sys.argv[1]}.log"):
    if line.startswith('{'):
        d=json.loads(line); print("bn", sys.argv[1], round(d["value"],1), round(d["ms_per_top_position"],4))
        for k in d["kernels"]:
            if k["kernel"].startswith("gemm") and ":256x" in k["kernel"]: print("   ", k["kernel"], k["us"])
P
done
