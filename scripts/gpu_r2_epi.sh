#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TMO=900 TAILN=4 run epi_tests python -m pytest tests -m gpu -q -x -p no:cacheprovider
bash scripts/gpu_r2_phases.sh
bash scripts/gpu_r2_ab.sh head current
TMO=400 TAILN=1 run epi_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
python - <<'P'
import json
for line in open("gpurun_out/epi_bench.log"):
    if line.startswith('{'):
        d=json.loads(line); print("bench", d["value"], d["ms_per_top_position"], d["roofline"]["frac"], d["roofline_gemm_all"]["frac"])
        for k in d["kernels"]: print("   ", k["kernel"], k["us"], k.get("frac_tensor"))
P
grep -A12 'gemm_fc1_gelu:256' gpurun_out/r2_gemm_phases.log | cut -c1-160
