#!/bin/bash
# decode-attention launch parameters re-swept on the final kernels (HQ_DEBUG switches): images/s at B = 256
mkdir -p gpurun_out
for cfg in "none" "HQ_ATTN_GROUPS=6" "HQ_ATTN_GROUPS=8" "HQ_ATTN_GROUPS=12" "HQ_ATTN_GROUPS=24" "HQ_ATTN_GROUPS=8 HQ_ATTM_STAGES=3"; do
  if [ "$cfg" = none ]; then env -u HQ_DEBUG timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/attn2.log 2>&1; else env HQ_DEBUG=1 $cfg timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/attn2.log 2>&1; fi
  grep '^{' gpurun_out/attn2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$cfg', round(d['value'],1), round(d['ms_per_top_position'],4), [(k['kernel'][-8:],k['us']) for k in d['kernels'] if 'attention_decode' in k['kernel']])" || tail -2 gpurun_out/attn2.log
done
