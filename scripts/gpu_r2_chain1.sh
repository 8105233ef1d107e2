#!/bin/bash
# round 2: first runs of the persistent chain kernel - exact-equality tests, then the suite and the bench line
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1200; }
TMO=400 TAILN=25 run r2_chain_tests python -m pytest tests/test_gpu_chain.py -x -q -p no:cacheprovider
if grep -q "exit=0" gpurun_out/r2_chain_tests.log; then
  TMO=400 TAILN=3 run r2_chain_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
  TMO=900 TAILN=6 run r2_chain_suite python -m pytest tests -m gpu -x -q -p no:cacheprovider
fi
