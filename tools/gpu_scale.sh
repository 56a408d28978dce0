#!/bin/bash
# Multi-GPU bench lines on one box: rollout (weak scaling) and full PPO (strong scaling) at N = 2, 4, 8 (as many as the box has).
tag=${1:-r01}
o=gpurun_out
ngpu=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  [ $n -le $ngpu ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 20 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_n$n.json') if l.startswith('{')][-1]); print('rollout N=$n value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --mode ppo --steps 100 --warmup 2 > $o/${tag}_bench_ppo_n$n.json 2> $o/${tag}_bench_ppo_n$n.err
  cut -c1-120 $o/${tag}_bench_ppo_n$n.json; grep -o '"split_ms.*' $o/${tag}_bench_ppo_n$n.json
done
