#!/bin/bash
# Round-2 GPU pass E: gpu suite (blocked learner input, HF collider with the plane-side cull), rough-terrain bench with / without the CTA barrier
# in the HF instantiations, opaque-thread-id variant of k_step, ncu of k_step<HF>.
tag=${1:-r02e}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
  ODUCK_CUDA_LIB=$V/liboduck_cuda_hfnobar.so timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_${E}_hfnobar.json 2> $o/${tag}_bench_rough_${E}_hfnobar.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_${E}_hfnobar.json')); print('rough hfnobar', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_${E}_hfnobar.err
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json
[ -f $V/liboduck_cuda_opaque.so ] && ODUCK_CUDA_LIB=$V/liboduck_cuda_opaque.so timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_opaque.json 2> $o/${tag}_bench_n1_opaque.err; cut -c1-200 $o/${tag}_bench_n1_opaque.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
du -sh $o; ls $o | grep ${tag}
