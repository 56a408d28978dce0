#!/bin/bash
# GPU-box pass for the device PPO learner: parity tests (CUDA-core twin and tcgen05), full-PPO bench, launch list.  Usage: tools/gpu_ppo.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_device.py -m gpu -q -s -x > gpurun_out/${tag}_pytest_ppo.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_ppo.log
tail -40 gpurun_out/${tag}_pytest_ppo.log
timeout 600 python -m pytest tests/test_ppo_device.py -m gpu -q -s > gpurun_out/${tag}_pytest_ppo_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_ppo_all.log
grep -E "passed|failed|rel err|forward:|loss:|adam step" gpurun_out/${tag}_pytest_ppo_all.log | tail -80
timeout 300 python bench.py --mode ppo --steps 100 --warmup 2 > gpurun_out/${tag}_bench_ppo.json 2> gpurun_out/${tag}_bench_ppo.err; cat gpurun_out/${tag}_bench_ppo.json; tail -5 gpurun_out/${tag}_bench_ppo.err
timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cat gpurun_out/${tag}_bench_n1.json; tail -5 gpurun_out/${tag}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1500 --csv --log-file gpurun_out/${tag}_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > gpurun_out/${tag}_launches_ppo.log 2>&1
ls -la gpurun_out | tail -8
