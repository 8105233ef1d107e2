#!/bin/bash
# per-CTA phases of single GEMM launches inside the replayed loop (instrumented twin of the library: build.py --phase-stamps)
mkdir -p gpurun_out
HQ_DEBUG=1 HQGRAFT_LIB=$PWD/hqtransformer_b200/libhqgraft_phases.so HQ_TRACE_PDL=1 timeout 600 python scripts/gemm_phases.py --per-cta > gpurun_out/r2_gemm_phases.log 2>&1; echo "exit=$?" >> gpurun_out/r2_gemm_phases.log
