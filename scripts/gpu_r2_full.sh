#!/bin/bash
# full GPU suite + smoke + benches after the issue-path rework (pair GEMM, single-CTA GEMM, stage-1 convolution)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-700; }
TMO=900 TAILN=4 run r2g_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=200 TAILN=2 run r2g_smoke python __graft_entry__.py smoke
TMO=500 TAILN=1 run r2g_bench python bench.py --steps 5 --warmup 3
for b in 64 128 512; do TMO=300 TAILN=1 run r2g_bench_b$b python bench.py --steps 3 --warmup 3 --batch $b --no-cpu-baseline --no-ref-gpu --no-kernel-table; done
TMO=300 TAILN=12 run r2g_s1 python scripts/s1_bench.py
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_bench*.log")):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, round(d["value"],1), round(d["ms_per_top_position"],4), d.get("extras",{}).get("stage1_decode",{}).get("value"))
P
