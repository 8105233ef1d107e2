#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_r2_ab.sh head current
bash scripts/gpu_r2_phases.sh
grep -B1 -A3 "cta: start" gpurun_out/r2_gemm_phases.log | grep -E "^gemm|cta:|start  " | cut -c1-900
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/try_tests.log 2>&1; echo "exit=$?" >> gpurun_out/try_tests.log; tail -3 gpurun_out/try_tests.log
