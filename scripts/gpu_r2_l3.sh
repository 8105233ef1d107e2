#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=400 TAILN=40 run r2_l3_tests python -m pytest tests/test_gpu_level3.py -x -q -p no:cacheprovider
TMO=900 TAILN=6 run r2_l3_suite python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_level3.py
