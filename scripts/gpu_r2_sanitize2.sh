#!/bin/bash
# compute-sanitizer memcheck over the reworked GEMM paths (3-D TMA boxes, both k-block instantiations, single-CTA and pair
# kernels, every epilogue incl. EPI_SAMPLE) and the stage-1 convolution
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/$name.log | head -8; }
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600"
TMO=900 run r2_san2_gemm $CS python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -x -k "every_tile_variant and (300 or 513 or 129)"
TMO=900 run r2_san2_loop $CS python -m pytest tests/test_gpu_sampling_loop.py -q -p no:cacheprovider -x -k "wide_batch or (greedy_codes and asym and True-True)"
TMO=900 run r2_san2_fused $CS python -m pytest tests/test_gpu_fused_sampler.py -q -p no:cacheprovider -x -k "shard or kernel_shape or chi"
TMO=900 run r2_san2_stage1 $CS python -m pytest tests/test_gpu_stage1.py -q -p no:cacheprovider -x -k "golden"
