#!/bin/bash
# Round-2 GPU pass AA: split-K factor of the dW GEMMs per precision mode, two runs each (run-to-run noise).
tag=${1:-r02aa}
o=gpurun_out
mkdir -p $o
for rep in a b; do for M in tf32 fp32; do for S in 8 16; do
  ODUCK_PPO_SPLITS=$S timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_${M}_s${S}_$rep.json 2> $o/${tag}_bench_ppo_${M}_s${S}_$rep.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_${M}_s${S}_$rep.json')); print('ppo $M splits $S $rep', j['value'], j['split_ms_per_training_step']['update_ms'])"
done; done; done
