#!/bin/bash
# Multi-GPU pass on one box (gpurun --gpus N): the headline rollout bench with the all-gather inside the timed region, full PPO in both
# update modes (sharded = gradient all-reduce per minibatch; replicated = ONE all-gather of the rollout + identical update on every rank).
# Usage: bash tools/gpu_multi.sh <tag> "<list of N>" [full]      ("full": also the driver-style default line with the extra legs at the largest N)
tag=${1:-r02e}
NS=${2:-"2"}
o=gpurun_out
mkdir -p $o
ngpu=$(nvidia-smi -L | wc -l)
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
last=1
for n in $NS; do
  [ $n -le $ngpu ] || continue
  last=$n
  run $n --steps 200 --warmup 20 --no-extra > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_n$n.json') if l.startswith('{')][-1]); print('rollout N=$n value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gather', d.get('gather'))" || tail -5 $o/${tag}_bench_n$n.err
  for mode in sharded replicated; do
    run $n --mode ppo --update-mode $mode --steps 100 --warmup 2 > $o/${tag}_bench_ppo_${mode}_n$n.json 2> $o/${tag}_bench_ppo_${mode}_n$n.err
    python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_ppo_${mode}_n$n.json') if l.startswith('{')][-1]); print('ppo $mode N=$n value', d['value'], d['update_mode'], d['split_ms_per_training_step'])" || tail -5 $o/${tag}_bench_ppo_${mode}_n$n.err
  done
done
if [ "$3" = "full" ]; then
  run $last --steps 20 --warmup 5 > $o/${tag}_bench_driver_n$last.json 2> $o/${tag}_bench_driver_n$last.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_driver_n$last.json') if l.startswith('{')][-1]); print('driver-style N=$last', d['value'], {k: d[k] for k in d if k.endswith('_per_s')})" || tail -5 $o/${tag}_bench_driver_n$last.err
  run $last --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref_n$last.json 2> $o/${tag}_bench_ref_n$last.err; cut -c1-200 $o/${tag}_bench_ref_n$last.json
fi
ls $o | grep ${tag}
