#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:conv_tc2 -s 212 -c 53 --csv --log-file gpurun_out/r2_s1_conv_metrics.csv python scripts/s1_bench.py 32 > gpurun_out/r2_s1d.log 2>&1; echo "exit=$?" >> gpurun_out/r2_s1d.log; tail -2 gpurun_out/r2_s1d.log
