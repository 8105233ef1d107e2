#!/bin/bash
mkdir -p gpurun_out
echo "=== invariance test"; timeout 300 python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -x -k "batch_is_cut" -p no:cacheprovider 2>&1 | tail -15
echo "=== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
for B in 256 64 128 1024; do
timeout 300 python bench.py --batch $B --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-table 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('B=$B', round(d['value'],1), 'img/s', round(d['ms_per_top_position']*1e3,1), 'us/pos')"
done
