#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-300; }
TMO=200 TAILN=3 run t_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider
TMO=300 TAILN=3 run t_loop python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -x -p no:cacheprovider
for P in 30 60; do timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=$P 2>&1 | grep -E "attention_decode|span"; done
timeout 120 python scripts/trace_loop.py 256 --no-pdl 2>&1 | grep -vE "attention_decode" | tail -22
TMO=200 TAILN=1 run bench python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-table
grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' gpurun_out/bench.log | head -3
