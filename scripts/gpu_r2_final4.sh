#!/bin/bash
# last evidence run of round 2 on the committed tree: GPU tests, smoke, bench exactly as the driver runs it + reference arm
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
TMO=900 TAILN=3 run r2y_tests python -m pytest tests -m gpu -q -p no:cacheprovider -rs
TMO=200 TAILN=2 run r2y_smoke python __graft_entry__.py smoke
TMO=600 TAILN=1 run r2y_bench python bench.py --gpus 1 --steps 20 --warmup 5
TMO=500 TAILN=1 run r2y_bench_ref python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
TMO=300 TAILN=1 run r2y_mt python -m hqtransformer_b200.measure_throughput model_path=hqtransformer_b200/configs/imagenet_l12.yaml batch_size=50 code_levels=2
