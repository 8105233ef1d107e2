#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/try_tests.log 2>&1; echo "exit=$?" >> gpurun_out/try_tests.log; tail -3 gpurun_out/try_tests.log
for r in 1 2; do timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-kernel-table > gpurun_out/try_bench_$r.log 2>&1; grep '^{' gpurun_out/try_bench_$r.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_top_position'],4), d['sharded_equals_single'])"; done
