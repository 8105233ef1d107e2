#!/bin/bash
# marginal cost of each kernel family in the real (graph + PDL) loop: step time with the family's launches dropped
mkdir -p gpurun_out
run() { timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-kernel-table 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s %8.1f img/s  %7.1f us/position' % ('$1', d['value'], d['ms_per_top_position']*1e3))"; }
run none
for f in layernorm attention_decode attention_fewkeys sample embed gemm_qkv gemm_fc1 gemm_head; do HQ_DEBUG=1 HQ_ABLATE=$f run $f; done
HQ_DEBUG=1 HQ_ABLATE=layernorm,attention,sample,embed run all_non_gemm
HQ_DEBUG=1 HQ_ABLATE=gemm run all_gemm
