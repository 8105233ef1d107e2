#!/bin/bash
# round 2, call 2: the widened parity suite, bench with the honest baselines, measure_throughput entry, sub-batch streams experiment
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1500; }
TMO=1200 TAILN=6 run r2_tests1 python -m pytest tests -m gpu -x -q -p no:cacheprovider -rs --durations=8
TMO=300 TAILN=8 run r2_two_stream python scripts/two_stream_experiment.py
TMO=200 TAILN=2 run r2_smoke1 python __graft_entry__.py smoke
TMO=500 TAILN=1 run r2_bench1 python bench.py --steps 5 --warmup 3
TMO=500 TAILN=1 run r2_bench_ref1 python bench.py --impl reference --steps 2 --warmup 1
TMO=300 TAILN=4 run r2_measure_throughput python -m hqtransformer_b200.measure_throughput model_path=hqtransformer_b200/configs/imagenet_l12.yaml batch_size=50 n_loop=3
