#!/bin/bash
# Round evidence run: GPU tests, smoke, full bench (N=1) + reference arm, other configs, batch sweep, device timeline,
# ncu launch list + ncu --set full capture of the real loop.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1200; }
TMO=900 TAILN=3 run final_tests python -m pytest tests -m gpu -x -q -p no:cacheprovider
TMO=200 TAILN=2 run final_smoke python __graft_entry__.py smoke
TMO=400 TAILN=2 run bench_full python bench.py --steps 5 --warmup 3
TMO=300 TAILN=2 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
echo "=== configs"; bash scripts/gpu_configs.sh 2>&1 | grep -E "img/s|exit=[1-9]"
echo "=== batch sweep"
for B in 64 128 512 1024; do
  KT="--no-kernel-table"; [ $B = 1024 ] && KT=""
  timeout 300 python bench.py --batch $B --steps 3 --warmup 2 --no-cpu-baseline $KT > gpurun_out/bench_b$B.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_b$B.log') if l.startswith('{')][-1])
print('B=$B', 'img/s', round(d['value'],1), 'ms/pos', round(d['ms_per_top_position'],3), 'e2e', round(d['e2e']['value'],1), 'attention', json.dumps(d.get('roofline_attention', {}).get('frac')))
PY
done
echo "=== timeline"; timeout 200 python scripts/trace_loop.py 256 --no-pdl 2>&1 | tee gpurun_out/trace_256.log | tail -26
for P in 8 60; do timeout 120 python scripts/trace_loop.py 256 --no-pdl --pos=$P 2>&1 | grep -E "attention_decode"; done
# launch list of the real loop: skip the first (warm-up) replay, list ~3 positions
TMO=600 TAILN=2 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 9400 -c 450 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table
# full capture of one spatial layer + some depth launches inside the real loop
TMO=900 TAILN=2 run ncu_full ncu --set full --clock-control none --import-source on -s 9400 -c 40 -f -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-table
ls -la gpurun_out | tail -6
