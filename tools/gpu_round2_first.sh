#!/bin/bash
# First GPU call of the next round: verify the shipped state, then A/B the variants prepared at the end of round 1.
# Before calling (CPU, builds travel with the snapshot):
#   python tools/variants.py build hfcull=-DODUCK_HF_CULL hfpairs=-DODUCK_HF_PAIRS symvilp=-DODUCK_SYMV_ILP \
#     cholldl=-DODUCK_CHOL_LDL symvunroll=-DODUCK_SYMV_UNROLL ancpipe=-DODUCK_ANC_PIPE \
#     "fast1=-DODUCK_CHOL_LDL -DODUCK_SYMV_UNROLL -DODUCK_ANC_PIPE" "fast1hf=-DODUCK_CHOL_LDL -DODUCK_SYMV_UNROLL -DODUCK_ANC_PIPE -DODUCK_HF_PAIRS"
# Usage on the box: bash tools/gpu_round2_first.sh r02a
tag=${1:-r02a}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 600 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit" $o/${tag}_pytest_gpu.log | tail -3
# (tests/test_zz_first_gpu_runs.py holds the first GPU run of k_step<HF, RL = true>: -rxX prints its outcome)
timeout 600 python -m pytest tests/test_zz_first_gpu_runs.py -m gpu -q -rxX > $o/${tag}_pytest_first_runs.log 2>&1; tail -4 $o/${tag}_pytest_first_runs.log
# parity of each variant library through the same tests (ODUCK_CUDA_LIB selects the build, capi.py)
for v in hfcull hfpairs symvilp cholldl symvunroll ancpipe fast1 fast1hf; do
  [ -f $V/liboduck_cuda_$v.so ] || continue
  ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 600 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu_$v.log 2>&1; echo "$v pytest exit $?" >> $o/${tag}_pytest_gpu_$v.log; tail -2 $o/${tag}_pytest_gpu_$v.log
done
# flat rollout: shipped vs the k_step variants; rough terrain: shipped vs the collider variants
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-220 $o/${tag}_bench_n1.json
for v in symvilp cholldl symvunroll ancpipe fast1; do
  [ -f $V/liboduck_cuda_$v.so ] && ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $o/${tag}_bench_n1_$v.json 2> $o/${tag}_bench_n1_$v.err; cut -c1-220 $o/${tag}_bench_n1_$v.json
done
# sub-batch pipelining of the rollout step (bench.py --pipeline P: P handles / graphs / streams; fills the 0.73-wave tail of k_step at 4096 envs)
for P in 2 4; do
  timeout 300 python bench.py --pipeline $P --steps 200 --warmup 20 > $o/${tag}_bench_n1_pipe$P.json 2> $o/${tag}_bench_n1_pipe$P.err; cut -c1-220 $o/${tag}_bench_n1_pipe$P.json
done
for P in 1 2; do
  timeout 600 python bench.py --mode ppo --pipeline $P --steps 100 --warmup 2 > $o/${tag}_bench_ppo_n1_pipe$P.json 2> $o/${tag}_bench_ppo_n1_pipe$P.err; cut -c1-160 $o/${tag}_bench_ppo_n1_pipe$P.json; grep -o '"split_ms.*' $o/${tag}_bench_ppo_n1_pipe$P.json
done
[ -f $V/liboduck_cuda_fast1.so ] && ODUCK_CUDA_LIB=$V/liboduck_cuda_fast1.so timeout 300 python bench.py --pipeline 2 --steps 200 --warmup 20 > $o/${tag}_bench_n1_pipe2_fast1.json 2> $o/${tag}_bench_n1_pipe2_fast1.err
timeout 300 python bench.py --task rough_terrain_backlash --steps 100 --warmup 10 --no-cpu-baseline > $o/${tag}_bench_rough.json 2> $o/${tag}_bench_rough.err; cut -c1-220 $o/${tag}_bench_rough.json
for v in hfcull hfpairs fast1hf; do
  [ -f $V/liboduck_cuda_$v.so ] && ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 300 python bench.py --task rough_terrain_backlash --steps 100 --warmup 10 --no-cpu-baseline > $o/${tag}_bench_rough_$v.json 2> $o/${tag}_bench_rough_$v.err; cut -c1-220 $o/${tag}_bench_rough_$v.json
done
# the capture round 1 had no budget left for
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --task rough_terrain_backlash --steps 6 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_k_step_hf.log 2>&1; tail -1 $o/${tag}_ncu_k_step_hf.log | cut -c1-200
ls $o | grep ${tag}
