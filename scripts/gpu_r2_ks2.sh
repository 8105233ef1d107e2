#!/bin/bash
# pair GEMM with two k-blocks per ring stage and warp-uniform issue: kernel parity, loop parity, loop profile, bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TMO=600 TAILN=4 run ks2_tests_gemm python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k gemm
HQ_DEBUG=1 HQ_GEMM_PROF=1 TMO=300 TAILN=30 run ks2_prof python scripts/gemm_prof.py
TMO=600 TAILN=4 run ks2_tests_loop python -m pytest tests/test_gpu_sampling_loop.py tests/test_gpu_full_size.py tests/test_gpu_fused_sampler.py -m gpu -q -x -p no:cacheprovider
TMO=400 TAILN=1 run ks2_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
for f in ks2_bench; do python - "$f" <<'P'
import json,sys
for line in open(f"gpurun_out/{sys.argv[1]}.log"):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d["value"], d["ms_per_top_position"], d["roofline"]["frac"], d["roofline_gemm_all"]["frac"])
        for k in d["kernels"]: print("   ", k["kernel"], k["us"], k.get("frac_tensor"))
P
done
