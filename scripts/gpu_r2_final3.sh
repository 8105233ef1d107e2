#!/bin/bash
# Round-2 evidence run on the final kernels: GPU tests, smoke, bench exactly as the driver runs it (N=1) + reference arm,
# in-loop GEMM phases (instrumented twin), ncu launch list + ncu --set full capture of the loop, batch sweep, stage-1 decode
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TMO=900 TAILN=3 run r2z_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=200 TAILN=2 run r2z_smoke python __graft_entry__.py smoke
TMO=600 TAILN=1 run r2z_bench python bench.py --gpus 1 --steps 20 --warmup 5
TMO=500 TAILN=1 run r2z_bench_ref python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
bash scripts/gpu_r2_phases.sh
for b in 64 128 512 1024; do TMO=300 TAILN=1 run r2z_bench_b$b python bench.py --steps 3 --warmup 3 --batch $b --no-cpu-baseline --no-ref-gpu --no-kernel-table; done
TMO=300 TAILN=4 run r2z_s1 python scripts/s1_bench.py
TMO=600 TAILN=2 run r2z_ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 9400 -c 450 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table --no-ref-gpu
TMO=900 TAILN=2 run r2z_ncu_full ncu --set full --clock-control none --import-source on -s 9400 -c 40 -f -o gpurun_out/prof_r2z python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table --no-ref-gpu
ls -la gpurun_out | tail -6
