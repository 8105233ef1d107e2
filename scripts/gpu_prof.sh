#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of one short bench run (warm-up launches skipped): per-launch device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9300 -c 450 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/launches_r1.csv
# (2) full capture of the GEMM and attention kernels run alone
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attention" -c 26 -o gpurun_out/prof_r1 -f \
    python scripts/prof_kernels.py 256 > gpurun_out/prof_kernels.log 2>&1
echo "full capture exit=$?"; ls -la gpurun_out/
