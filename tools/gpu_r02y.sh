#!/bin/bash
# Round-2 GPU pass Y: split-K factor of the weight-gradient GEMMs (ODUCK_PPO_SPLITS) against the reduce kernel's partial traffic.
tag=${1:-r02y}
o=gpurun_out
mkdir -p $o
for S in 16 8 12 6; do
  ODUCK_PPO_SPLITS=$S timeout 600 python bench.py --mode ppo --learner-matmul tf32 --steps 100 --warmup 2 > $o/${tag}_bench_ppo_tf32_s$S.json 2> $o/${tag}_bench_ppo_tf32_s$S.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_tf32_s$S.json')); print('ppo tf32 splits $S', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_tf32_s$S.err
done
ODUCK_PPO_SPLITS=8 timeout 600 python -m pytest tests/test_ppo_device.py -m gpu -q 2>&1 | tail -2
