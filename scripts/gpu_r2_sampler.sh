#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/smp_tests.log 2>&1; echo "exit=$?" >> gpurun_out/smp_tests.log; tail -3 gpurun_out/smp_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --top-k 2048 --temperature 0.95 > gpurun_out/smp_topk.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --top-k 2048 --top-p 0.95 --temperature 0.95 > gpurun_out/smp_topkp.log 2>&1
for f in smp_topk smp_topkp; do grep "^{" gpurun_out/$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), round(d['ms_per_top_position'],4), [ (k['kernel'],k['us']) for k in d['kernels'] if k['kernel'].startswith('sample') or k['kernel'].startswith('gemm_head')])
"; done
