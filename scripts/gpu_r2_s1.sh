#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=300 TAILN=30 run r2_s1_tests python -m pytest tests/test_gpu_stage1.py -q -p no:cacheprovider
TMO=300 TAILN=6 run r2_s1_bench python scripts/s1_bench.py 1 8 32
