#!/bin/bash
# Multi-GPU pass on one box (gpurun --gpus N): the driver-style default line (rollout with the all-gather inside the timed region + the
# extra legs: physics only, full PPO sharded fp32 / tf32, rough terrain) and full PPO in replicated mode (ONE all-gather of the rollout +
# identical update on every rank) for comparison with sharded (gradient all-reduce per minibatch, captured into the epoch graph).
# Usage: bash tools/gpu_multi.sh <tag> "<list of N>" [ref|-] [lite]      ("ref": also the CPU reference arm at the largest N; "lite": only the default line and replicated tf32)
tag=${1:-r02k}
NS=${2:-"2"}
o=gpurun_out
mkdir -p $o
ngpu=$(nvidia-smi -L | wc -l)
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
last=1
for n in $NS; do
  [ $n -le $ngpu ] || continue
  last=$n
  run $n --steps 200 --warmup 20 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_n$n.json') if l.startswith('{')][-1]); print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gather', d.get('gather')); print('   ', {k: d[k] for k in d if k.endswith('_per_s')}); print('    ppo', d.get('ppo_leg', {}).get('split_ms_per_training_step'), d.get('ppo_leg', {}).get('update_mode'), d.get('ppo_leg', {}).get('error')); print('    ppo_tf32', d.get('ppo_tf32_leg', {}).get('split_ms_per_training_step'), d.get('ppo_tf32_leg', {}).get('error'))" || tail -5 $o/${tag}_bench_n$n.err
  for M in $( [ "$4" = "lite" ] && echo tf32 || echo fp32 tf32 ); do
    run $n --mode ppo --update-mode replicated --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_replicated_${M}_n$n.json 2> $o/${tag}_bench_ppo_replicated_${M}_n$n.err
    python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_ppo_replicated_${M}_n$n.json') if l.startswith('{')][-1]); print('ppo replicated $M N=$n value', d['value'], d['update_mode'], d['split_ms_per_training_step'])" || tail -5 $o/${tag}_bench_ppo_replicated_${M}_n$n.err
  done
  [ "$4" = "lite" ] && continue
  ODUCK_PPO_GRAPH_NCCL=0 run $n --mode ppo --update-mode sharded --learner-matmul tf32 --steps 100 --warmup 2 > $o/${tag}_bench_ppo_sharded_nograph_n$n.json 2> $o/${tag}_bench_ppo_sharded_nograph_n$n.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_ppo_sharded_nograph_n$n.json') if l.startswith('{')][-1]); print('ppo sharded tf32 (all-reduce outside the graphs) N=$n value', d['value'], d['update_mode'], d['split_ms_per_training_step'])" || tail -5 $o/${tag}_bench_ppo_sharded_nograph_n$n.err
done
if [ "$3" = "ref" ]; then
  run $last --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref_n$last.json 2> $o/${tag}_bench_ref_n$last.err; cut -c1-200 $o/${tag}_bench_ref_n$last.json
fi
ls $o | grep ${tag}
