#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=300 TAILN=4 run r2_s1c_tests python -m pytest tests/test_gpu_stage1.py -q -p no:cacheprovider -x
TMO=300 TAILN=4 run r2_s1c_bench python scripts/s1_bench.py 32
TMO=600 TAILN=2 run r2_s1c_ncu ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'s1_|conv_tc2' -s 435 -c 145 --csv --log-file gpurun_out/r2_s1c_launches.csv python scripts/s1_bench.py 32
