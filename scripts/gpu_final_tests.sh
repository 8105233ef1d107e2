#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
for P in 30 60; do timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=$P 2>&1 | grep -E "attention_decode"; done
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' | head -2
