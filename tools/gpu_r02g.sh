#!/bin/bash
# Round-2 GPU pass G: full gpu suite, rough-terrain bench on settled states with the barrier-mask variants of the HF instantiations, ncu of k_step<HF>.
tag=${1:-r02g}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
  for v in hfbar9 hfbar0b hfbar19; do
    ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_${E}_$v.json 2> $o/${tag}_bench_rough_${E}_$v.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_${E}_$v.json')); print('rough $v', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_${E}_$v.err
  done
done
for M in fp32 tf32; do for P in 1 0; do
  ODUCK_PPO_PDL=$P timeout 600 python bench.py --mode ppo --learner-matmul $M --steps 5 --warmup 2 > $o/${tag}_bench_ppo_${M}_pdl$P.json 2> $o/${tag}_bench_ppo_${M}_pdl$P.err; python -c "import json; j=json.load(open('$o/${tag}_bench_ppo_${M}_pdl$P.json')); print('ppo $M pdl=$P', j['value'], j['split_ms_per_training_step'])"; tail -2 $o/${tag}_bench_ppo_${M}_pdl$P.err
done; done
ODUCK_CUDA_LIB=$V/liboduck_cuda_hfbar9.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 130 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
du -sh $o; ls $o | grep ${tag}
