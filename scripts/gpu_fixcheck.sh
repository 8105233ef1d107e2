#!/bin/bash
timeout 150 python scripts/diag_asym.py 2>&1 | grep "graph=True pdl=True" | awk '{print $0}' | cut -c1-110 | sort | uniq -c
timeout 200 python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-table 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_top_position": [0-9.]*' | head -2
