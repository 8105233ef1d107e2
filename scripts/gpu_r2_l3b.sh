#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
L3_REFERENCE=1 TMO=600 TAILN=5 run r2_l3_bench python scripts/level3_bench.py 64 256
TMO=300 TAILN=6 run r2_l3_mt python -m hqtransformer_b200.measure_throughput model_path=hqtransformer_b200/configs/imagenet_l12_level3.yaml batch_size=50 code_levels=3 n_loop=2 n_samples=200
TMO=300 TAILN=6 run r2_mt2 python -m hqtransformer_b200.measure_throughput model_path=hqtransformer_b200/configs/imagenet_l12.yaml batch_size=50 code_levels=2 n_loop=3 n_samples=500
