#!/bin/bash
# Round-2 GPU pass Z: GAE merged into the loss kernel, 8-way split-K of the dW GEMMs: learner tests + PPO bench.
tag=${1:-r02z}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 2 $o/${tag}_smoke.log
for M in tf32 fp32; do
  timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_$M.json 2> $o/${tag}_bench_ppo_$M.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_$M.json')); print('ppo $M', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_$M.err
done
ODUCK_PPO_SPLITS=10 timeout 600 python bench.py --mode ppo --learner-matmul tf32 --steps 100 --warmup 2 > $o/${tag}_bench_ppo_tf32_s10.json 2> $o/${tag}_bench_ppo_tf32_s10.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_tf32_s10.json')); print('ppo tf32 splits 10', j['value'], j['split_ms_per_training_step'])"
