#!/bin/bash
mkdir -p gpurun_out
echo "=== attention kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k attention -p no:cacheprovider 2>&1 | tail -3
for ST in 4 3 2; do
echo "=== mma kernel phases stages=$ST per_sm=4"; HQ_ATTM_STAGES=$ST HQ_ATTM_PER_SM=4 timeout 120 python scripts/attn_phases.py 256 32 64 2>&1 | grep phases | grep -v "keys=.*ctas" | awk 'NR%2==1 || 1' | head -24
echo "=== bench stages=$ST per_sm=4"; HQ_ATTM_STAGES=$ST HQ_ATTM_PER_SM=4 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_top_position'], d['roofline_attention']['achieved'])"
done
echo "=== bench scalar"; HQ_ATTN_SCALAR=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_top_position'], d['roofline_attention']['achieved'])"
