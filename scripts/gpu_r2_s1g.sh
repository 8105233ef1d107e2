#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'s1_|conv_tc2' -s 145 -c 145 --csv --log-file gpurun_out/r2_s1f_launches.csv python scripts/s1_bench.py 32 > gpurun_out/r2_s1f_ncu.log 2>&1
tail -2 gpurun_out/r2_s1f_ncu.log | cut -c1-200
