#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-300; }
TMO=200 run t_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider
TMO=300 run t_loop python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -x -p no:cacheprovider
TMO=120 TAILN=30 run trace_256 python scripts/trace_loop.py 256
TMO=120 TAILN=30 run trace_256_nopdl python scripts/trace_loop.py 256 --no-pdl
TMO=120 TAILN=8 run trace_gemm python scripts/trace_gemm.py
TMO=200 TAILN=1 run bench python bench.py --steps 3 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print('value',round(d['value'],1),'ms/pos',round(d['ms_per_top_position'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
for k in d.get('kernels',[]): print('  ',k['kernel'],k['us'],k.get('tflops',k.get('gbs')),k.get('frac_tensor'),k.get('frac_hbm'))
PY
