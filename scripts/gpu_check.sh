#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests in separate processes (a trapped kernel kills only its own
# process), smoke, then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=300 run t_gemm_tc python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_tcgen05" -p no:cacheprovider
TMO=300 run t_gemm_f32 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_fp32" -p no:cacheprovider
TMO=300 run t_sample python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sample" -p no:cacheprovider
TMO=900 TAILN=60 run t_loop python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -s -p no:cacheprovider
TMO=300 run smoke python __graft_entry__.py smoke
TMO=900 run bench python bench.py --steps 3 --warmup 3
