#!/bin/bash
# Round-2 GPU pass O: counted named barriers (synccheck), vectorised gradient reduce / Adam kernels.
tag=${1:-r02o}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -8
SAN_ENVS=24 timeout 500 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_run.py flat hf > $o/${tag}_sanitizer_synccheck.log 2>&1; tail -n 3 $o/${tag}_sanitizer_synccheck.log
SAN_ENVS=24 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py flat hf > $o/${tag}_sanitizer_racecheck.log 2>&1; tail -n 3 $o/${tag}_sanitizer_racecheck.log
for M in fp32 tf32; do
  timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 100 --warmup 2 > $o/${tag}_bench_ppo_$M.json 2> $o/${tag}_bench_ppo_$M.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_$M.json')); print('ppo $M', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_$M.err
done
timeout 600 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --mode rough --rough-envs 16384 --steps 40 > $o/${tag}_bench_rough_16384.json 2> $o/${tag}_bench_rough_16384.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_16384.json')); print('rough', 16384, j['value'], j['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1300 --csv --log-file $o/${tag}_launches_ppo_tf32.csv python bench.py --mode ppo --learner-matmul tf32 --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
python tools/launch_summary.py $o/${tag}_launches_ppo_tf32.csv | grep -E "adam|reduce|pack|gae|loss|total"
ls $o | grep ${tag}
