#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-200}; }
TMO=300 TAILN=2 run r2_kvpf_on python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
HQ_DEBUG=1 HQ_NO_KV_PREFETCH=1 TMO=300 TAILN=2 run r2_kvpf_off python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
TMO=300 TAILN=3 run r2_kvpf_tests python -m pytest tests/test_gpu_sampling_loop.py tests/test_gpu_full_size.py -q -p no:cacheprovider -x
