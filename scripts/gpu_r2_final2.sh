#!/bin/bash
# Round-2 final evidence: whole GPU suite, smoke, default bench (with extras) and the reference arm
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TMO=1200 TAILN=4 run r2g_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=300 TAILN=2 run r2g_smoke python __graft_entry__.py smoke
TMO=600 TAILN=2 run r2g_bench python bench.py
TMO=400 TAILN=2 run r2g_bench_ref python bench.py --impl reference --steps 2 --warmup 1
