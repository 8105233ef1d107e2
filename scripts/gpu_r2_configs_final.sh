#!/bin/bash
# one B200, final kernels: the other BASELINE configs at batch 256 (short form: no CPU arm, no kernel table)
mkdir -p gpurun_out
i=0
for cfg in "--model imagenet_l24" "--model imagenet_l42 --top-k 2048 --top-p 0.95 --temperature 0.95" "--model cc15m_l12 --top-k 2048" "--top-k 2048 --temperature 0.95" "--model ffhq_l24"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-kernel-table $cfg > gpurun_out/r2y_cfg$i.log 2>&1
  echo "== $cfg"; grep '^{' gpurun_out/r2y_cfg$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_top_position'],4), d['sharded_equals_single'])" || tail -2 gpurun_out/r2y_cfg$i.log
done
