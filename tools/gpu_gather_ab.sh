#!/bin/bash
# A/B of the rollout all-gather at N GPUs: once at the unroll boundary vs slice by slice behind the steps, with CTA limits for the slice collectives.
tag=${1:-r02q}; n=${2:-2}
o=gpurun_out
mkdir -p $o
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 200 --warmup 20 --no-extra "$@"; }
IFS=","; for cfg in ${CFGS:-boundary 0,sliced 0,sliced 2,sliced 4,sliced 8}; do
  IFS=" "; set -- $cfg
  run --gather $1 --gather-max-ctas $2 > $o/${tag}_bench_n${n}_$1_$2.json 2> $o/${tag}_bench_n${n}_$1_$2.err
  python -c "
import json; d=json.loads([l for l in open('$o/${tag}_bench_n${n}_$1_$2.json') if l.startswith('{')][-1]); g=d['gather']; print('N=$n gather $1 max_ctas $2: value', d['value'], 'ms', d['ms_per_step'], 'exposed', g.get('exposed_ms_per_unroll'))" || tail -5 $o/${tag}_bench_n${n}_$1_$2.err
done
