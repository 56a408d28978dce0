#!/bin/bash
# Full GPU-box pass: all gpu tests, rollout bench, reference arm, full-PPO bench, launch lists.  Usage: tools/gpu_all.sh <tag> [ncu]
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
grep -E "passed|failed|pytest exit|one-call|^E  " gpurun_out/${tag}_pytest_gpu.log | tail -20
timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cat gpurun_out/${tag}_bench_n1.json; tail -3 gpurun_out/${tag}_bench_n1.err
timeout 300 python bench.py --mode ppo --steps 100 --warmup 2 > gpurun_out/${tag}_bench_ppo.json 2> gpurun_out/${tag}_bench_ppo.err; cat gpurun_out/${tag}_bench_ppo.json; tail -3 gpurun_out/${tag}_bench_ppo.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/${tag}_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > gpurun_out/${tag}_launches_ppo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ls gpurun_out | grep ${tag}
