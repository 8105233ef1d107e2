#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/try_tests.log 2>&1; echo "exit=$?" >> gpurun_out/try_tests.log; tail -3 gpurun_out/try_tests.log
bash scripts/gpu_r2_ab.sh head current
