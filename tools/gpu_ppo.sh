#!/bin/bash
# GPU-box pass for the device PPO learner: parity tests (CUDA-core twin and tcgen05), full-PPO bench, launch list, ncu of the GEMM.  Usage: tools/gpu_ppo.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_device.py -m gpu -q -s > gpurun_out/${tag}_pytest_ppo.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_ppo.log
grep -E "passed|failed|rel err|forward:|loss:|adam step|Error|pytest exit" gpurun_out/${tag}_pytest_ppo.log | grep -v "print(" | tail -90
timeout 300 python bench.py --mode ppo --steps 100 --warmup 2 > gpurun_out/${tag}_bench_ppo.json 2> gpurun_out/${tag}_bench_ppo.err; cat gpurun_out/${tag}_bench_ppo.json; tail -5 gpurun_out/${tag}_bench_ppo.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/${tag}_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > gpurun_out/${tag}_launches_ppo.log 2>&1
if [ -n "$NCU_FULL" ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 66 -c 22 -o gpurun_out/${tag}_k_gemm -f python bench.py --mode ppo --steps 20 --warmup 1 > gpurun_out/${tag}_ncu_gemm.log 2>&1; fi
ls -la gpurun_out | tail -6
