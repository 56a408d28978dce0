#!/bin/bash
# Round-2 GPU pass AG: --ftz=true / --use_fast_math builds of the library against the default (flat bench + env parity tests).
tag=${1:-r02ag}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
for v in default ftz fast; do
  L=$V/liboduck_cuda_$v.so; [ $v = default ] && L=open_duck_playground_b200/csrc/liboduck_cuda.so
  ODUCK_CUDA_LIB=$L timeout 300 python bench.py --steps 300 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_$v.json 2> $o/${tag}_bench_n1_$v.err; python -c "import json; j=json.load(open('$o/${tag}_bench_n1_$v.json')); print('flat $v', j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_n1_$v.err
done
for v in ftz fast; do
  ODUCK_CUDA_LIB=$V/liboduck_cuda_$v.so timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_standing.py -m gpu -q --deselect tests/test_parity_gpu.py::test_library_is_the_cuda_one 2>&1 | tail -4 | sed "s/^/$v: /"
done
