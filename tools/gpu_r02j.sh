#!/bin/bash
# Round-2 GPU pass J (WPB = 16 default): gpu suite, sub-batch pipeline sweep, PPO pipeline sweep, rough bench, ncu --set full of k_step (flat, HF).
tag=${1:-r02j}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
for P in 1 2 3 4 6 8; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline --pipeline $P > $o/${tag}_bench_n1_p$P.json 2> $o/${tag}_bench_n1_p$P.err; python -c "import json; j=json.load(open('$o/${tag}_bench_n1_p$P.json')); print('flat pipeline $P', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'])"; tail -2 $o/${tag}_bench_n1_p$P.err
done
for P in 1 2 4; do
  timeout 600 python bench.py --mode ppo --learner-matmul tf32 --ppo-pipeline $P --steps 5 --warmup 2 > $o/${tag}_bench_ppo_tf32_p$P.json 2> $o/${tag}_bench_ppo_tf32_p$P.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_tf32_p$P.json')); print('ppo tf32 pipeline $P', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_tf32_p$P.err
done
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
du -sh $o; ls $o | grep ${tag}
