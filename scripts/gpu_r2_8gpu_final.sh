#!/bin/bash
# round 2, final kernels: 8 x B200 - bit-equality of the sharded sampler and BASELINE configs 2 / 4 / 5 (short form: no CPU arm)
mkdir -p gpurun_out
N=${1:-8}
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f tr); N=$N tr 29541 scripts/multi_gpu_check.py" > gpurun_out/r2_multi_gpu_check_n$N.log 2>&1; echo "exit=$?" >> gpurun_out/r2_multi_gpu_check_n$N.log
grep -E "MULTI_GPU|exit=" gpurun_out/r2_multi_gpu_check_n$N.log
i=0
for cfg in "--batch 256" "--batch 64" "--model imagenet_l42 --top-k 2048 --top-p 0.95 --temperature 0.95" "--model cc15m_l12 --top-k 2048"; do
  i=$((i+1))
  name=r2z_n${N}_cfg$i
  timeout 300 bash -c "$(declare -f tr); N=$N tr $((29550+i)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-kernel-table $cfg" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log
  echo "== $cfg"; grep '^{' gpurun_out/$name.log | cut -c1-150; grep -o '"sharded_equals_single": [a-z]*' gpurun_out/$name.log; tail -1 gpurun_out/$name.log
done
