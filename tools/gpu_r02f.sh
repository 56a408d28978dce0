#!/bin/bash
# Round-2 GPU pass F: group-parallel height-field clipping (hf_clip_pass).  HF parity tests, rough-terrain bench (default / extra CTA barrier
# before the Newton phase), ncu --set full of k_step<HF>.
tag=${1:-r02f}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
timeout 900 python -m pytest tests/test_hfield.py tests/test_zz_first_gpu_runs.py -m gpu -q -s -rxX > $o/${tag}_pytest_hf.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_hf.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL|out of tolerance" $o/${tag}_pytest_hf.log | tail -12
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
  ODUCK_CUDA_LIB=$V/liboduck_cuda_hfbar9.so timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_${E}_hfbar9.json 2> $o/${tag}_bench_rough_${E}_hfbar9.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_${E}_hfbar9.json')); print('rough hfbar9', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_${E}_hfbar9.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
du -sh $o; ls $o | grep ${tag}
