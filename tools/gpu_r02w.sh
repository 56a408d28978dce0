#!/bin/bash
# Round-2 GPU pass W (final single-GPU record): gpu suite, smoke, default bench line, reference arm, ncu --set full of k_step<HF>, launch lists.
tag=${1:-r02w}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 3 $o/${tag}_smoke.log
timeout 600 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-160 $o/${tag}_bench_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_launches_rollout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1400 --csv --log-file $o/${tag}_launches_ppo_tf32.csv python bench.py --mode ppo --learner-matmul tf32 --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
python tools/launch_summary.py $o/${tag}_launches_ppo_tf32.csv | tail -4
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"
done
ls $o | grep ${tag}
