#!/bin/bash
mkdir -p gpurun_out
for G in 1 2 4; do
  echo "=== groups=$G, PDL in trace"; HQ_ATTN_GROUPS=$G HQ_TRACE_PDL=1 timeout 120 python scripts/trace_loop.py 256 2>&1 | grep -E "attention_decode|span"
  echo "=== groups=$G, no PDL"; HQ_ATTN_GROUPS=$G timeout 120 python scripts/trace_loop.py 256 --no-pdl 2>&1 | grep -E "attention_decode|span"
done
