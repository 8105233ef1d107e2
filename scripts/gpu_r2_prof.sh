#!/bin/bash
mkdir -p gpurun_out
HQ_DEBUG=1 HQ_GEMM_PROF=1 timeout 300 python scripts/gemm_prof.py > gpurun_out/r2_gemm_prof_ks2.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_prof_ks2.log
HQ_DEBUG=1 HQ_GEMM_PROF=1 HQ_GEMM_KS1=1 timeout 300 python scripts/gemm_prof.py > gpurun_out/r2_gemm_prof_ks1.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_prof_ks1.log
cat gpurun_out/r2_gemm_prof_ks2.log; echo; cat gpurun_out/r2_gemm_prof_ks1.log
