#!/bin/bash
# Round-2 GPU pass I: one 16-warp CTA per SM (-DWPB=16: a single instruction stream per SM) against two 8-warp CTAs, flat and rough.
tag=${1:-r02i}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
for v in default wpb16 wpb16bar1; do
  L=$V/liboduck_cuda_$v.so; [ $v = default ] && L=open_duck_playground_b200/csrc/liboduck_cuda.so
  ODUCK_CUDA_LIB=$L timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_$v.json 2> $o/${tag}_bench_n1_$v.err; python -c "import json; j=json.load(open('$o/${tag}_bench_n1_$v.json')); print('flat $v', j['value'], j['ms_per_step'], j['roofline'].get('kernel_ms_full_batch'))"; tail -2 $o/${tag}_bench_n1_$v.err
  ODUCK_CUDA_LIB=$L timeout 300 python bench.py --steps 200 --warmup 20 --no-extra --no-cpu-baseline --pipeline 1 > $o/${tag}_bench_n1_p1_$v.json 2> $o/${tag}_bench_n1_p1_$v.err; python -c "import json; j=json.load(open('$o/${tag}_bench_n1_p1_$v.json')); print('flat pipeline 1 $v', j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_n1_p1_$v.err
  for E in 4096 16384; do
    ODUCK_CUDA_LIB=$L timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_${E}_$v.json 2> $o/${tag}_bench_rough_${E}_$v.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_${E}_$v.json')); print('rough $v', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_${E}_$v.err
  done
done
ODUCK_CUDA_LIB=$V/liboduck_cuda_wpb16.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_hfield.py -m gpu -q -x 2>&1 | tail -3
ls $o | grep ${tag}
