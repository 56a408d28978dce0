#!/bin/bash
# Round-2 GPU pass N: compute-sanitizer on the round-2b kernels (group-parallel HF collider, 16-warp CTAs), batch-size sweep (configs[4]),
# launch lists of the rollout and of full PPO (TF32), sass summary inputs.
tag=${1:-r02n}
o=gpurun_out
mkdir -p $o
SAN_ENVS=24 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py flat hf policy > $o/${tag}_sanitizer_memcheck.log 2>&1; tail -n 3 $o/${tag}_sanitizer_memcheck.log
SAN_ENVS=24 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py flat hf policy > $o/${tag}_sanitizer_racecheck.log 2>&1; tail -n 3 $o/${tag}_sanitizer_racecheck.log
SAN_ENVS=24 timeout 400 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_run.py flat hf > $o/${tag}_sanitizer_synccheck.log 2>&1; tail -n 3 $o/${tag}_sanitizer_synccheck.log
SAN_ENVS=24 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py ppo > $o/${tag}_sanitizer_memcheck_ppo.log 2>&1; tail -n 3 $o/${tag}_sanitizer_memcheck_ppo.log
timeout 600 python tools/sweep.py > $o/${tag}_sweep_n1.jsonl 2> $o/${tag}_sweep_n1.err; cat $o/${tag}_sweep_n1.jsonl | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_launches_rollout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1300 --csv --log-file $o/${tag}_launches_ppo_tf32.csv python bench.py --mode ppo --learner-matmul tf32 --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
python tools/launch_summary.py $o/${tag}_launches_rollout.csv | tail -12
ls $o | grep ${tag}
