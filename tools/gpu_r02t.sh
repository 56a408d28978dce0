#!/bin/bash
# Round-2 GPU pass T: unrolled M pair pass (-DODUCK_MPAIR_UNROLL=2 / 4) against the default build.
tag=${1:-r02t}
o=gpurun_out
mkdir -p $o
V=open_duck_playground_b200/csrc/variants
for v in default mp2 mp4 default; do
  L=$V/liboduck_cuda_$v.so; [ $v = default ] && L=open_duck_playground_b200/csrc/liboduck_cuda.so
  ODUCK_CUDA_LIB=$L timeout 300 python bench.py --steps 300 --warmup 20 --no-extra --no-cpu-baseline > $o/${tag}_bench_n1_$v.json 2> $o/${tag}_bench_n1_$v.err; python -c "import json; j=json.load(open('$o/${tag}_bench_n1_$v.json')); print('flat $v', j['value'], j['ms_per_step'], j['roofline'].get('kernel_ms_full_batch'))"; tail -2 $o/${tag}_bench_n1_$v.err
done
ls $o | grep ${tag}
