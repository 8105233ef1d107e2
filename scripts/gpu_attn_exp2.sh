#!/bin/bash
for P in 4 14 60; do
  echo "=== pos=$P"; timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=$P 2>&1 | grep -E "attention_decode"
done
