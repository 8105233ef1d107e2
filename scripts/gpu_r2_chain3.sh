#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TMO=400 TAILN=4 run r2_chain3_tests python -m pytest tests/test_gpu_chain.py -x -q -p no:cacheprovider
if grep -q "exit=0" gpurun_out/r2_chain3_tests.log; then
  TMO=300 TAILN=2 run r2_chain3_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
  HQ_DEBUG=1 HQ_CHAIN_NO_L2PF=1 TMO=300 TAILN=2 run r2_chain3_bench_nol2pf python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu
  TMO=300 TAILN=80 CUT=200 run r2_chain3_phases python scripts/chain_phases.py
fi
