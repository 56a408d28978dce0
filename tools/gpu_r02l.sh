#!/bin/bash
# Round-2 GPU pass L: gpu suite on the committed default build, rough bench, per-launch DRAM traffic of k_step at 2048 envs (the sub-batch size of the default pipeline).
tag=${1:-r02l}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED|XPASS|XFAIL" $o/${tag}_pytest_gpu.log | tail -12
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"; tail -2 $o/${tag}_bench_rough_$E.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_step -s 30 -c 4 --csv --log-file $o/${tag}_traffic_2048.csv python bench.py --pipeline 1 --envs-per-gpu 2048 --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_traffic_2048.log 2>&1; tail -5 $o/${tag}_traffic_2048.csv | cut -c1-200
timeout 600 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-160 $o/${tag}_bench_ref.json
ls $o | grep ${tag}
