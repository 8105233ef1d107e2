#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TMO=900 TAILN=25 run r2_fuse_tests python -m pytest tests/test_gpu_fused_sampler.py -q -p no:cacheprovider -x
TMO=600 TAILN=2 CUT=3000 run r2_fuse_bench python bench.py --steps 3 --warmup 3
