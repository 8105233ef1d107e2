#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=400 TAILN=6 run r2_s1f_tests python -m pytest tests/test_gpu_stage1.py -q -p no:cacheprovider
TMO=300 TAILN=4 run r2_s1f_bench python scripts/s1_bench.py 8 32
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'s1_|conv_tc2' -s 145 -c 145 --csv --log-file gpurun_out/r2_s1f_launches.csv python scripts/s1_bench.py 32 > gpurun_out/r2_s1f_ncu.log 2>&1
