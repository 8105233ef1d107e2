#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TMO=300 run t_gemm_tiles python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "every_tile" -p no:cacheprovider
TMO=300 run t_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not every_tile" -p no:cacheprovider
TMO=900 run t_loop python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -p no:cacheprovider
TMO=900 TAILN=3 run bench python bench.py --steps 3 --warmup 3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print('value',d['value'],'ms/pos',d['ms_per_top_position'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
for k in d.get('kernels',[]): print(k['kernel'],k['us'],k.get('tflops',k.get('gbs')),k.get('frac_tensor'),k.get('frac_hbm'))
print(d.get('roofline')); print(d.get('cpu_baseline'))
PY
