#!/bin/bash
mkdir -p gpurun_out
for B in 64 128 512 1024; do
  timeout 300 python bench.py --batch $B --steps 3 --warmup 2 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_b$B.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_b$B.log') if l.startswith('{')][-1])
print('B=$B', 'img/s', round(d['value'],1), 'ms/pos', round(d['ms_per_top_position'],3), 'e2e', round(d['e2e']['value'],1))
PY
done
timeout 200 python scripts/trace_loop.py 1024 --no-pdl | tail -16
