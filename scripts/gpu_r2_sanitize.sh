#!/bin/bash
# compute-sanitizer memcheck over the small-model tests of the kernels added in round 2 (chain, variants, stage 1, 3 levels)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/$name.log | head -8; }
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600"
TMO=900 run r2_san_stage1 $CS python -m pytest tests/test_gpu_stage1.py -q -p no:cacheprovider -x -k "golden or errors or chunks"
TMO=900 run r2_san_level3 $CS python -m pytest tests/test_gpu_level3.py -q -p no:cacheprovider -x -k "greedy and not True"
TMO=900 run r2_san_variants $CS python -m pytest tests/test_gpu_variants.py -q -p no:cacheprovider -x -k "greedy and False"
TMO=900 run r2_san_chain $CS python -m pytest tests/test_gpu_chain.py -q -p no:cacheprovider -x -k "logits_equal and ASYM"
TMO=900 run r2_san_loop $CS python -m pytest tests/test_gpu_sampling_loop.py -q -p no:cacheprovider -x -k "shared_text_prefix or (greedy_codes and tiny and False-False)"
