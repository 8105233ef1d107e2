#!/bin/bash
# Round-2 evidence run: GPU tests, smoke, full bench (N=1) + reference arm, ncu launch list + ncu --set full capture of the loop
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TMO=900 TAILN=3 run r2f_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=200 TAILN=2 run r2f_smoke python __graft_entry__.py smoke
TMO=500 TAILN=2 run r2f_bench python bench.py --steps 5 --warmup 3
TMO=400 TAILN=2 run r2f_bench_ref python bench.py --impl reference --steps 2 --warmup 1
# launch list of the real loop: skip the first (warm-up) replay, list ~3 positions
TMO=600 TAILN=2 run r2f_ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 9400 -c 450 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table --no-ref-gpu
# full capture of one spatial layer + some depth launches inside the real loop
TMO=900 TAILN=2 run r2f_ncu_full ncu --set full --clock-control none --import-source on -s 9400 -c 40 -f -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table --no-ref-gpu
ls -la gpurun_out | tail -8
