#!/bin/bash
mkdir -p gpurun_out
HQ_DEBUG=1 timeout 600 python scripts/sweep_plans.py > gpurun_out/r2_gemm_plan_sweep.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_plan_sweep.log
tail -5 gpurun_out/r2_gemm_plan_sweep.log
