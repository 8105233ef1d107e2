#!/bin/bash
# Round evidence run: full bench (N=1), launch list + ncu full capture of the real loop, reference arm.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1500; }
TMO=400 TAILN=2 run bench_full python bench.py --steps 5 --warmup 3
TMO=300 TAILN=2 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
# launch list of the real loop: skip the first (warm-up) replay, list ~3 positions
TMO=600 TAILN=2 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 9400 -c 450 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table
# full capture of one spatial layer + some depth launches inside the real loop
TMO=900 TAILN=2 run ncu_full ncu --set full --clock-control none --import-source on -s 9400 -c 40 -f -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table
ls -la gpurun_out | tail -8
