#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=600 TAILN=30 run r2_variants_tests python -m pytest tests/test_gpu_variants.py -x -q -p no:cacheprovider
TMO=900 TAILN=8 run r2_variants_suite python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_variants.py
TMO=300 TAILN=2 CUT=200 run r2_variants_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
TMO=300 TAILN=2 CUT=200 run r2_ffhq_bench python bench.py --model ffhq_l24 --top-k 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu
