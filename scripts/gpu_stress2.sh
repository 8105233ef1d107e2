#!/bin/bash
for i in 1 2 3 4; do
  timeout 120 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1
  echo "run $i exit=$?"
done
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' | head -2
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-table --no-pdl 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' | head -2
timeout 100 python bench.py --steps 3 --warmup 2 --batch 64 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
timeout 100 python bench.py --steps 3 --warmup 2 --batch 300 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
