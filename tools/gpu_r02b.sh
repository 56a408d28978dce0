#!/bin/bash
# Round-2 GPU pass B: smoke (all kernel families), full gpu suite, the new bench line (pipeline 4 + extra legs + full-size CPU arm), pipeline / LDL A/B,
# launch lists, --set full + source of flat k_step, compute-sanitizer logs.  Usage on the box: bash tools/gpu_r02b.sh r02b
tag=${1:-r02b}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -4 $o/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-160 $o/${tag}_bench_ref.json; tail -2 $o/${tag}_bench_ref.err
timeout 600 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_n1_driver.json 2> $o/${tag}_bench_n1_driver.err; cut -c1-200 $o/${tag}_bench_n1_driver.json
for P in 1 2 8; do
  timeout 300 python bench.py --pipeline $P --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_pipe$P.json 2> $o/${tag}_bench_n1_pipe$P.err; cut -c1-200 $o/${tag}_bench_n1_pipe$P.json; tail -2 $o/${tag}_bench_n1_pipe$P.err
done
[ -f $V/liboduck_cuda_ldl.so ] && ODUCK_CUDA_LIB=$V/liboduck_cuda_ldl.so timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_ldl.json 2> $o/${tag}_bench_n1_ldl.err; cut -c1-200 $o/${tag}_bench_n1_ldl.json
timeout 600 python bench.py --mode ppo --steps 100 --warmup 2 > $o/${tag}_bench_ppo_n1.json 2> $o/${tag}_bench_ppo_n1.err; cat $o/${tag}_bench_ppo_n1.json | cut -c1-700; tail -2 $o/${tag}_bench_ppo_n1.err
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --pipeline 1 --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_launches_rollout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file $o/${tag}_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > $o/${tag}_launches_ppo.log 2>&1
# the dominant kernel: --set full with source, one full-batch launch of flat k_step (4096 envs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; tail -2 $o/${tag}_ncu_k_step.log | cut -c1-200
# compute-sanitizer: memcheck on every kernel family, racecheck + synccheck on the per-warp kernels
SAN_ENVS=16 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py flat hf policy > $o/${tag}_sanitizer_memcheck.log 2>&1; tail -3 $o/${tag}_sanitizer_memcheck.log
SAN_ENVS=16 timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py flat policy > $o/${tag}_sanitizer_racecheck.log 2>&1; tail -3 $o/${tag}_sanitizer_racecheck.log
SAN_ENVS=16 timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_run.py flat > $o/${tag}_sanitizer_synccheck.log 2>&1; tail -3 $o/${tag}_sanitizer_synccheck.log
SAN_ENVS=16 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py ppo > $o/${tag}_sanitizer_memcheck_ppo.log 2>&1; tail -3 $o/${tag}_sanitizer_memcheck_ppo.log
ls $o | grep ${tag}
