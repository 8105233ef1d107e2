#!/bin/bash
# round 2: N-GPU bit-equality of the sharded sampler (post split-K / PDL fixes) + the bench line with sharded_equals_single
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n$N.log 2>&1; echo "exit=$?" >> gpurun_out/r2_multi_gpu_check_n$N.log
grep -E "MULTI_GPU|exit=" gpurun_out/r2_multi_gpu_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "exit=$?" >> gpurun_out/r2_bench_n$N.log
grep '^{' gpurun_out/r2_bench_n$N.log | cut -c1-400; tail -1 gpurun_out/r2_bench_n$N.log
