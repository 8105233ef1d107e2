#!/bin/bash
# round 2: 8 x B200 - bit-equality of the sharded sampler, BASELINE configs 2-5 (weak scaling per GPU batch 64 / 256 / 512,
# the largest model with top-k / top-p, text-to-image with the prefix cache).  One JSON line per run under gpurun_out/.
mkdir -p gpurun_out
N=${1:-8}
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f tr); N=$N tr 29541 scripts/multi_gpu_check.py" > gpurun_out/r2_multi_gpu_check_n$N.log 2>&1; echo "exit=$?" >> gpurun_out/r2_multi_gpu_check_n$N.log
grep -E "MULTI_GPU|exit=" gpurun_out/r2_multi_gpu_check_n$N.log
i=0
for cfg in "--batch 256" "--batch 64" "--batch 512" "--model imagenet_l42 --top-k 2048 --top-p 0.95 --temperature 0.95" "--model cc15m_l12 --top-k 2048"; do
  i=$((i+1))
  name=r2_n${N}_cfg$i
  timeout 400 bash -c "$(declare -f tr); N=$N tr $((29550+i)) bench.py --gpus $N --steps 3 --warmup 3 $cfg" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log
  echo "== $cfg"; grep '^{' gpurun_out/$name.log | cut -c1-160; tail -1 gpurun_out/$name.log
done
