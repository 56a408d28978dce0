#!/bin/bash
# Round-2 GPU pass C: gpu suite with the tightened tolerances + new height-field collider, k_step variant A/B (chain-specialised sweeps are the
# default now; LDL, 9-warp CTAs, no CTA barrier), rough-terrain bench + ncu of k_step<HF>, env-count sweep, k_step at the sub-batch size.
tag=${1:-r02c}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
for v in ldl wpb9 nobar wpb9ldl; do
  [ -f $V/liboduck_cuda_$v.so ] || continue
  ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_$v.json 2> $o/${tag}_bench_n1_$v.err; cut -c1-200 $o/${tag}_bench_n1_$v.json; tail -2 $o/${tag}_bench_n1_$v.err
done
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; cut -c1-300 $o/${tag}_bench_rough_$E.json; tail -2 $o/${tag}_bench_rough_$E.err
done
timeout 900 python tools/sweep.py > $o/${tag}_sweep_n1.jsonl 2> $o/${tag}_sweep_n1.err; cat $o/${tag}_sweep_n1.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; tail -1 $o/${tag}_ncu_k_step_hf.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; tail -1 $o/${tag}_ncu_k_step.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step_1024 -f python bench.py --pipeline 1 --envs-per-gpu 1024 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step_1024.log 2>&1; tail -1 $o/${tag}_ncu_k_step_1024.log | cut -c1-200
SAN_ENVS=16 timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py flat hf policy > $o/${tag}_sanitizer_racecheck.log 2>&1; tail -3 $o/${tag}_sanitizer_racecheck.log
SAN_ENVS=16 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py flat hf policy > $o/${tag}_sanitizer_memcheck.log 2>&1; tail -3 $o/${tag}_sanitizer_memcheck.log
ls $o | grep ${tag}
