#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, full ncu capture of k_step.  Usage: tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 300 --warmup 30 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cat gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 300 --warmup 30 > gpurun_out/${tag}_bench_ref.json 2>&1; cat gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --mode ppo --steps 100 --warmup 2 > gpurun_out/${tag}_bench_ppo.json 2> gpurun_out/${tag}_bench_ppo.err; cat gpurun_out/${tag}_bench_ppo.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 2 -o gpurun_out/${tag}_k_step -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
