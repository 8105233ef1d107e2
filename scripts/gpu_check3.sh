#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TMO=300 run t_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider
TMO=900 run t_loop python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -p no:cacheprovider
TMO=600 TAILN=1 run bench_nopdl python bench.py --steps 3 --warmup 3 --no-pdl --no-cpu-baseline --no-kernel-table
TMO=600 TAILN=1 run bench python bench.py --steps 3 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
for f in ('bench_nopdl','bench'):
    d=json.loads([l for l in open(f'gpurun_out/{f}.log') if l.startswith('{')][-1])
    print(f,'value',round(d['value'],1),'ms/pos',round(d['ms_per_top_position'],3),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
    for k in d.get('kernels',[]): print('  ',k['kernel'],k['us'],k.get('tflops',k.get('gbs')),k.get('frac_tensor'),k.get('frac_hbm'))
PY
