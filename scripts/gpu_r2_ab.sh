#!/bin/bash
# A/B of library builds on one box: $1.. = library files under hqtransformer_b200/ (the head build lacks the newest debug symbol:
# the binding of that one symbol is dropped for it)
mkdir -p gpurun_out
cp hqtransformer_b200/libhqgraft.so /tmp/lib_current.so
for round in 1 2; do
for v in "$@"; do
  if [ "$v" = current ]; then cp /tmp/lib_current.so hqtransformer_b200/libhqgraft.so; else cp hqtransformer_b200/libhqgraft_$v.so hqtransformer_b200/libhqgraft.so; fi
  cp hqtransformer_b200/_lib.py /tmp/_lib.py.bak
  if ! nm -D hqtransformer_b200/libhqgraft.so | grep -q hq_debug_gemm_phases; then python - <<'P'
import re
p='hqtransformer_b200/_lib.py'; s=open(p).read()
s=re.sub(r'    "hq_debug_gemm_phases": \(C\.c_int, \[.*?\]\),\n', '', s, flags=re.S)
open(p,'w').write(s)
P
  fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-kernel-table > gpurun_out/ab_${v}_$round.log 2>&1
  cp /tmp/_lib.py.bak hqtransformer_b200/_lib.py
  grep '^{' gpurun_out/ab_${v}_$round.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $round, round(d['value'],1), round(d['ms_per_top_position'],4))" || tail -3 gpurun_out/ab_${v}_$round.log
done
done
cp /tmp/lib_current.so hqtransformer_b200/libhqgraft.so
