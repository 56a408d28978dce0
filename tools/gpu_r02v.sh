#!/bin/bash
# Round-2 GPU pass V: whole-update graph + input prefetch: gpu suite, PPO bench fp32 / tf32.
tag=${1:-r02v}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -8
for M in fp32 tf32; do
  timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_$M.json 2> $o/${tag}_bench_ppo_$M.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_$M.json')); print('ppo $M', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_$M.err
done
ls $o | grep ${tag}
