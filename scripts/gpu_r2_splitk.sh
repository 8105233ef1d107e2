#!/bin/bash
# plan experiment on the final kernels: pinned split-K factor of every residual GEMM (HQ_FORCE_SPLITK) vs the cost model's choice
mkdir -p gpurun_out
for sk in 0 2 3 4 6; do
  if [ $sk = 0 ]; then unset HQ_DEBUG HQ_FORCE_SPLITK; else export HQ_DEBUG=1 HQ_FORCE_SPLITK=$sk; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/splitk_$sk.log 2>&1
  python - $sk <<'P'
import json,sys
for line in open(f"gpurun_out/splitk_{sys.argv[1]}.log"):
    if line.startswith('{'):
        d=json.loads(line); print("splitk", sys.argv[1], round(d["value"],1), round(d["ms_per_top_position"],4), " ".join(f"{k['kernel']}={k['us']}" for k in d["kernels"] if k["kernel"].startswith("gemm_resid") or k["kernel"].startswith("layernorm")))
P
done
