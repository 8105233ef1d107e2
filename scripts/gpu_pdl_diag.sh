#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-400; }
TMO=100 run pdl_stream python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-table
TMO=100 run pdl_graph python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table
TMO=100 run pdl_graph_b64 python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-kernel-table
