#!/bin/bash
# round 2, call 1: grid-barrier microbenchmark, sanity of the tree on this pool, baseline bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 120 scripts/grid_barrier_bench.bin 2000 > gpurun_out/r2_grid_barrier.log 2>&1; echo "exit=$?" >> gpurun_out/r2_grid_barrier.log
cat gpurun_out/r2_grid_barrier.log
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_tests0.log 2>&1; echo "exit=$?" >> gpurun_out/r2_tests0.log
tail -3 gpurun_out/r2_tests0.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench0.log 2>&1; echo "exit=$?" >> gpurun_out/r2_bench0.log
tail -2 gpurun_out/r2_bench0.log | cut -c1-600
