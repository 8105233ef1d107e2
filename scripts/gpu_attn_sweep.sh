#!/bin/bash
mkdir -p gpurun_out
show='
import json,sys
d=json.loads(sys.stdin.read()); a=[k for k in d["kernels"] if k["kernel"].startswith("attention_decode")]
print("   %.1f img/s  %.1f us/pos  attention in-loop %.2f us avg (%s)  %.0f GB/s" % (d["value"], d["ms_per_top_position"]*1e3, sum(k["us"] for k in a)/len(a), ",".join("%.1f"%k["us"] for k in a), d["roofline_attention"]["achieved"]))'
for cfg in "4 4" "4 2" "8 4" "8 2" "8 3" "12 4" "6 4" "2 2"; do
  set -- $cfg
  echo "=== groups=$1 stages=$2"
  HQ_DEBUG=1 HQ_ATTN_GROUPS=$1 HQ_ATTM_STAGES=$2 timeout 120 python scripts/attn_phases.py 256 32 2>&1 | grep -E "first_keys|first_item|cta_end" | head -3
  HQ_DEBUG=1 HQ_ATTN_GROUPS=$1 HQ_ATTM_STAGES=$2 timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "$show"
done
