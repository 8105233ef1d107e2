#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-300; }
TMO=120 TAILN=16 run trace_b8 python scripts/trace_loop.py 8 --no-pdl
TMO=120 TAILN=16 run trace_b256_L1 python scripts/trace_loop.py 256 --no-pdl --layers=1
TMO=120 TAILN=16 run trace_b8_L1 python scripts/trace_loop.py 8 --no-pdl --layers=1
