#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], 'img/s', round(d['value'],1), 'ms/pos', round(d['ms_per_top_position'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'GB', round(d['device_bytes']/1e9,2))
PY
}
timeout 400 python bench.py --model imagenet_l42 --top-k 2048 --top-p 0.95 --temperature 0.95 --steps 2 --warmup 2 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_l42.log 2>&1; echo "exit=$?"; show gpurun_out/bench_l42.log
timeout 400 python bench.py --model cc15m_l12 --top-k 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_txt.log 2>&1; echo "exit=$?"; show gpurun_out/bench_txt.log
timeout 400 python bench.py --model imagenet_l12 --top-k 2048 --temperature 0.95 --top-p 1.0 --steps 2 --warmup 2 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_l12_topk.log 2>&1; echo "exit=$?"; show gpurun_out/bench_l12_topk.log
timeout 400 python bench.py --model imagenet_l24 --steps 2 --warmup 2 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_l24.log 2>&1; echo "exit=$?"; show gpurun_out/bench_l24.log
