#!/bin/bash
for NS in 0 32 100 300; do
  echo "== sleep $NS"; HQ_ATTN_SLEEP=$NS timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=60 2>&1 | grep -E "attention_decode:t6[24]"
  HQ_ATTN_SLEEP=$NS timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=30 2>&1 | grep -E "attention_decode:t3[24]"
done
