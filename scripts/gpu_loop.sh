#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampling_loop.py -m gpu -q -s -p no:cacheprovider > gpurun_out/t_loop.log 2>&1; echo "exit=$?" >> gpurun_out/t_loop.log
grep -E "passed|failed|FAILED|max-abs|agreement|AssertionError: [(n]" gpurun_out/t_loop.log | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
