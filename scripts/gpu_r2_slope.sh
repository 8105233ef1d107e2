#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sweep_gemm_slope.py > gpurun_out/r2_gemm_slope_ks2.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_slope_ks2.log
HQ_DEBUG=1 HQ_GEMM_KS1=1 timeout 300 python scripts/sweep_gemm_slope.py > gpurun_out/r2_gemm_slope_ks1.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_slope_ks1.log
grep '#' gpurun_out/r2_gemm_slope_ks2.log; grep '#' gpurun_out/r2_gemm_slope_ks1.log
