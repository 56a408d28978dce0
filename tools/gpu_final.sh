#!/bin/bash
# Round-end GPU pass on one B200: smoke, all gpu tests, reference arm, headline bench (with cpu_baseline), full-PPO bench, envs sweep,
# launch lists and --set full captures of the dominant kernels.  Usage: tools/gpu_final.sh <tag>
tag=${1:-r01}
o=gpurun_out
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -s > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit" $o/${tag}_pytest_gpu.log | tail -3
timeout 300 python bench.py --impl reference --steps 200 --warmup 20 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-200 $o/${tag}_bench_ref.json
timeout 400 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-400 $o/${tag}_bench_n1.json; tail -2 $o/${tag}_bench_n1.err
timeout 300 python bench.py --mode ppo --steps 100 --warmup 2 > $o/${tag}_bench_ppo_n1.json 2> $o/${tag}_bench_ppo_n1.err; cat $o/${tag}_bench_ppo_n1.json
timeout 600 python tools/sweep.py > $o/${tag}_sweep_n1.jsonl 2> $o/${tag}_sweep_n1.err; cat $o/${tag}_sweep_n1.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $o/${tag}_launches_rollout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file $o/${tag}_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o $o/${tag}_k_step -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_k_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 66 -c 22 -o $o/${tag}_k_gemm -f python bench.py --mode ppo --steps 20 --warmup 1 > $o/${tag}_ncu_k_gemm.log 2>&1
ls $o | grep ${tag}
