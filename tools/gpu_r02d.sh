#!/bin/bash
# Round-2 GPU pass D: gpu suite (tf32 learner, gait-clocked reward terms, LDL default), the full default bench line, rough-terrain bench,
# PPO fp32 vs tf32, ncu captures exported to CSV on the box (outputs stay under gpurun's 64 MiB).
tag=${1:-r02d}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-160 $o/${tag}_bench_ref.json
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
done
for M in fp32 tf32; do
  timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_$M.json 2> $o/${tag}_bench_ppo_$M.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_$M.json')); print('ppo $M', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_$M.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file $o/${tag}_launches_ppo_tf32.csv python bench.py --mode ppo --learner-matmul tf32 --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_launches_rollout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step_1024 -f python bench.py --pipeline 1 --envs-per-gpu 1024 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step_1024.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_1024.ncu-rep
du -sh $o; ls $o | grep ${tag}
