#!/bin/bash
echo "== all-lane polling"; timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=30 2>&1 | grep -E "attention_decode|layernorm "
echo "== elected-lane wait"; HQ_ATTN_WAITWARP=1 timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=30 2>&1 | grep -E "attention_decode|layernorm "
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
