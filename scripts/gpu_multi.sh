#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/multi_gpu_check.py 2>&1 | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-kernel-table > gpurun_out/bench_n2.log 2>&1; echo "exit=$?"; tail -n 2 gpurun_out/bench_n2.log | cut -c1-900
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "exit=$?"; tail -n 1 gpurun_out/bench_n1.log | cut -c1-2500
