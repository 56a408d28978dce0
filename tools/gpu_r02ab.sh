#!/bin/bash
# Round-2 GPU pass AB: one-pass observation statistics in the PPO update.
tag=${1:-r02ab}
o=gpurun_out
mkdir -p $o
for M in tf32 fp32; do
  timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_$M.json 2> $o/${tag}_bench_ppo_$M.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_$M.json')); print('ppo $M', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_$M.err
done
timeout 600 python -m pytest tests/test_ppo_device.py tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -2
